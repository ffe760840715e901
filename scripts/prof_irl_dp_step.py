import contextlib, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from discrete_mean_field_game_b200.ac_irl import AC_IRL
dev = torch.device("cuda:0"); D = 15
rng = np.random.RandomState(5)
g = rng.standard_gamma(1.0, size=(64, D)); mat = g / g.sum(1, keepdims=True)
with contextlib.redirect_stdout(sys.stderr):
    irl = AC_IRL(theta=8.64, shift=0, alpha_scale=1e4, d=D, reg="none", n_fc3=8, n_fc4=4, mat_pi0=mat, demonstrations=[], device=dev, seed=1, net_seed=2)
    Bi = 1 << 20
    g2 = rng.standard_gamma(1.0, size=(1 << 14, D)).astype(np.float32)
    pii = torch.as_tensor(g2 / g2.sum(1, keepdims=True), device=dev).repeat(Bi >> 14, 1).contiguous()
    Md = 4096
    dsd, dad = irl.generate_batch(Md, theta=8.06)
    dsd, dad = dsd[:15].reshape(-1, D).contiguous(), dad.reshape(-1, D, D)
    irl.irl_step_batch(pii[:1 << 12], dsd, dad, Md, episode=1)
    torch.cuda.synchronize()
    irl.irl_step_batch(pii, dsd, dad, Md, episode=2)
    torch.cuda.synchronize()
print("done")
