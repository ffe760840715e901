import contextlib, os, sys, tempfile, time, cProfile, pstats, io
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from discrete_mean_field_game_b200.ac_irl import AC_IRL
rng = np.random.RandomState(0)
mat = rng.dirichlet(np.ones(15), size=21)
root = os.getcwd()
os.chdir(tempfile.mkdtemp())
with contextlib.redirect_stdout(sys.stderr):
    ac = AC_IRL(theta=8.64, shift=0, alpha_scale=1e4, d=15, reg="dropout_l1l2", n_fc3=8, n_fc4=4, mat_pi0=mat, demonstrations=[], seed=1, net_seed=2)
    ac.theta = 8.06
    ac.list_demonstrations = ac.generate_trajectories(40)
    ac.list_demonstrations_test = ac.generate_trajectories(10)
    ac.list_eval_demo_transitions = [pair for traj in ac.list_demonstrations for pair in traj]
    ac.theta = 8.64
    ac.outerloop(num_iterations=1, max_reward_iterations=10, max_forward_episodes=10, final_episodes=10, verbose=False)
    torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    ac.outerloop(verbose=False)
    torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(30); print(s.getvalue()[:6000])
