"""AC_IRL.outerloop at the reference's own defaults (ac_irl.py:900-954: 20 iterations x (100 reward updates on 5 + 5 trajectories,
200 forward episodes), 2000 final episodes) on synthetic demonstrations: wall time and, under ncu, the launch list."""
import contextlib, os, random, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from discrete_mean_field_game_b200.ac_irl import AC_IRL
rng = np.random.RandomState(0)
mat = rng.dirichlet(np.ones(15), size=21)
small = len(sys.argv) > 1 and sys.argv[1] == "small"
os.chdir(tempfile.mkdtemp())


def run(small):
    with contextlib.redirect_stdout(sys.stderr):
        ac = AC_IRL(theta=8.64, shift=0, alpha_scale=1e4, d=15, reg="dropout_l1l2", n_fc3=8, n_fc4=4, mat_pi0=mat,
                    demonstrations=[], seed=1, net_seed=2)
        ac.theta = 8.06
        ac.list_demonstrations = ac.generate_trajectories(40)       # "expert" trajectories from another policy
        ac.list_demonstrations_test = ac.generate_trajectories(10)
        ac.list_eval_demo_transitions = [pair for traj in ac.list_demonstrations for pair in traj]
        ac.theta = 8.64
        random.seed(0)                                              # update_reward draws its minibatches with random.sample
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if small:
            ac.outerloop(num_iterations=2, max_reward_iterations=20, max_forward_episodes=20, final_episodes=50, verbose=False)
        else:
            ac.outerloop(verbose=False)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("outerloop (%s): %.2f s, theta = %.5f, %d reward updates, %d forward episodes" % (
        "small" if small else "reference defaults", dt, ac.theta, ac.reward_params.step, ac._episodes))


if small:
    run(True)
else:
    # the loop is one-CTA kernels and 5-CTA launches: on a box that has been idle the first pass also pays the clock ramp
    # and the first-use costs (module load, allocator growth) -- both passes are printed
    for k in range(int(os.environ.get("PASSES", "3"))):
        run(False)
