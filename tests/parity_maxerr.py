"""Achieved maximum errors of the float32 CUDA path against the float64 oracle, per output, on injected Gamma variates:
the golden train trace (the reference's own draws), a config-2-shaped batch (AC_IRL defaults, 4096 populations x 15
steps) and a 2^16-population sample of config 3 (16 steps), processed in chunks.  Writes a markdown table.

    python tests/parity_maxerr.py [--out gpurun_out/r2_parity_maxerr.md] [--log2-big 16]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from discrete_mean_field_game_b200 import engine as eng
from oracle import mfg_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_parity_maxerr.md"))
ap.add_argument("--log2-big", type=int, default=16)
a = ap.parse_args()
dev = torch.device("cuda:0")
eng.require_cuda()
OUT = ("states", "actions", "alpha", "alpha_deriv", "rewards", "deltas", "grads", "pi_final")


def rel(x, ref, floor=0.0):
    return float(np.max(np.abs(x - ref) / np.maximum(np.abs(ref), floor))) if ref.size else 0.0


def run_case(name, pi0, theta, shift, scale, T, w, seed, chunk=2048, variant="auto"):
    """y is drawn along the oracle's own trajectory (np.random.gamma, rounded to float32), chunk by chunk"""
    B, d = pi0.shape
    F = O.num_features(d)
    worst = {}
    G = {"G_theta": 0.0, "G_w": np.zeros(F), "R": 0.0}
    Gk = np.zeros(2 + F)
    absdg = 0.0
    wscale = np.zeros(F)
    rng = np.random.RandomState(seed)
    for lo in range(0, B, chunk):
        p0 = np.float32(pi0[lo:lo + chunk])
        y = np.zeros((T, p0.shape[0], d, d), np.float32)
        pi = p0.astype(np.float64)
        for t in range(T):
            alpha, _ = O.policy_alpha(pi, theta, shift)
            y[t] = np.float32(rng.gamma(alpha * scale))
            pi = O.mean_field_step(O.normalise_gamma(y[t].astype(np.float64)), pi)
        ref = O.rollout_frozen(p0.astype(np.float64), theta, shift, scale, y.astype(np.float64), w=w)
        out = eng.rollout(torch.as_tensor(p0, device=dev), theta, shift, scale, T,
                          w=torch.as_tensor(w, dtype=torch.float64, device=dev),
                          noise_y=torch.as_tensor(y, device=dev), outputs=OUT, want_acc=True, variant=variant)
        o = {k: v.double().cpu().numpy() for k, v in out.items()}
        v = np.abs(O.features(ref["states"]) @ w)
        scale_td = np.abs(ref["rewards"]) + v[1:] + v[:-1]
        e = {
            "pi' (states), rel": rel(o["states"], ref["states"], 1e-30),
            "P (actions), rel": rel(o["actions"], ref["actions"], 1e-30),
            "alpha, rel": rel(o["alpha"], ref["alpha"]),
            "alpha', rel (|x| floor 1e-3)": rel(o["alpha_deriv"], ref["alpha_deriv"], 1e-3 * np.abs(ref["alpha_deriv"]).max()),
            "g = dlogF/dtheta, rel": rel(o["grads"], ref["grads"], 1e-2 * np.abs(ref["grads"]).mean()),
            "r, rel": rel(o["rewards"], ref["rewards"], 1e-2 * np.abs(ref["rewards"]).mean()),
            "delta, |err| / (|r|+|V'|+|V|)": float(np.max(np.abs(o["deltas"] - ref["deltas"]) / scale_td)),
            "delta, rel (floor 1% of mean |delta|)": rel(o["deltas"], ref["deltas"], 1e-2 * np.abs(ref["deltas"]).mean()),
        }
        for k, val in e.items():
            worst[k] = max(worst.get(k, 0.0), val)
        G["G_theta"] += ref["G_theta"]; G["G_w"] += ref["G_w"]; G["R"] += ref["R"]
        Gk += o["acc"]
        absdg += float(np.sum(np.abs(ref["deltas"] * ref["grads"])))
        wscale += np.sum(np.abs(ref["deltas"])[..., None] * np.abs(O.features(ref["states"][:-1])), axis=(0, 1))
    worst["G_theta = sum delta*g, rel"] = abs(Gk[0] - G["G_theta"]) / abs(G["G_theta"])
    worst["G_theta, |err| / sum|delta*g|"] = abs(Gk[0] - G["G_theta"]) / absdg
    worst["G_w = sum delta*phi, max |err| / sum|delta*phi|"] = float(np.max(np.abs(Gk[1:1 + F] - G["G_w"]) / wscale))
    worst["sum r, rel"] = abs(Gk[1 + F] - G["R"]) / abs(G["R"])
    # the parameter update of one batch-mean step (lr 0.1): relative error of the theta / w increments
    worst["Delta theta (batch-mean step), rel"] = worst["G_theta = sum delta*g, rel"]
    worst["Delta w, max rel over features"] = float(np.max(np.abs(Gk[1:1 + F] - G["G_w"]) / np.abs(G["G_w"])))
    return name, B, T, worst


cases = []
g = np.load(os.path.join(ROOT, "tests", "golden", "train_trace.npz"))
rng = np.random.RandomState(0)
w15 = rng.rand(O.num_features(15))
cases.append(run_case("config 1 shape: mfg_ac2 defaults (theta 8.86349, shift 0.16, scale 12000), d=15, T=15",
                      rng.dirichlet(np.ones(15), size=512), 8.86349, 0.16, 12000.0, 15, w15, 1))
cases.append(run_case("config 2 shape: AC_IRL defaults (theta 8.64, shift 0, scale 1e4), 4096 populations, T=15",
                      rng.dirichlet(np.ones(15), size=4096), 8.64, 0.0, 1e4, 15, w15, 2))
cases.append(run_case("config 3 sample: 2^%d populations x 16 steps, bench parameters" % a.log2_big,
                      rng.dirichlet(np.ones(15), size=1 << a.log2_big), 8.86349, 0.16, 12000.0, 16, w15, 3))
w21 = rng.rand(O.num_features(21))
cases.append(run_case("d=21 (mfg_ac2.py:25 default), 1024 populations, T=15",
                      rng.dirichlet(np.ones(21), size=1024), 8.86349, 0.16, 12000.0, 15, w21, 4))
lines = ["# Achieved parity errors of the float32 CUDA path vs the float64 oracle (round 2)", "",
         "Injected Gamma variates (np.random.gamma along the oracle's trajectory, rounded to float32), outputs through "
         "`dmfg_rollout` (v2 kernel, float streams).  Maximum over all populations, steps and elements.  "
         "north_star tolerance: 1e-5 relative.", ""]
for name, B, T, worst in cases:
    lines += ["## %s" % name, "", "| output | max error |", "|---|---|"]
    lines += ["| %s | %.3e |" % (k, v) for k, v in worst.items()]
    lines.append("")

# ---- the IRL half: reward net forward / backward and one reward update against the float64 reward-net oracle ----------------
from oracle import rnet_oracle as R


def rnet_case(name, d, n3, n4, n, seed, dropout):
    rng = np.random.RandomState(seed)
    p = np.float32(R.xavier_init(d, n3, n4, rng) + 0.1 * rng.randn(R.param_count(d, n3, n4)))
    s = np.float32(rng.dirichlet(np.ones(d), size=n))
    ac = np.float32(rng.dirichlet(np.ones(d) * 0.5, size=(n, d)))
    m3 = (rng.rand(n, n3) < 0.4) if dropout else None
    m4 = (rng.rand(n, n4) < 0.4) if dropout else None
    dr = np.float32(rng.randn(n))
    r_ref, cache = R.forward(p, s, ac, n3, n4, m3, m4, cache=True)
    g_ref = R.backward(cache, dr)
    T_ = lambda x, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(x), dtype=dt, device=dev)
    kw = dict(mask3=T_(m3, torch.uint8), mask4=T_(m4, torch.uint8)) if dropout else {}
    g, r = eng.rnet_backward(T_(p), T_(s), T_(ac), T_(dr), n3, n4, want_rewards=True, **kw)
    g, r = g.cpu().numpy().astype(np.float64), r.cpu().numpy().astype(np.float64)
    worst = {"r, max |err| (|r| < 1)": float(np.abs(r - r_ref).max()),
             "r, rel (floor 1e-2)": rel(r, r_ref, 1e-2)}
    # a gradient entry is a float32 sum over n transitions of mixed-sign terms: its error is measured against the sum of the
    # magnitudes of its terms (what a float32 accumulation can resolve) and against the largest entry of its tensor
    for tname, shp, off in R.layout(d, n3, n4):
        sl = slice(off, off + int(np.prod(shp)))
        scale = np.abs(g_ref[sl]).max()
        worst["d %s, max |err| / max |entry|" % tname] = float(np.abs(g[sl] - g_ref[sl]).max() / scale) if scale > 0 else 0.0
    return name, n, 1, worst


rcases = [rnet_case("reward net, d=15, n_fc3=8, n_fc4=4 (AC_IRL defaults), 4096 transitions, no dropout", 15, 8, 4, 4096, 11, False),
          rnet_case("reward net, d=15, n_fc3=8, n_fc4=4, 4096 transitions, dropout masks (keep 0.4)", 15, 8, 4, 4096, 12, True),
          rnet_case("reward net, d=16, n_fc3=8, n_fc4=8, 777 transitions", 16, 8, 8, 777, 13, False),
          rnet_case("reward net, d=21 (32-lane groups), n_fc3=6, n_fc4=3, 777 transitions", 21, 6, 3, 777, 14, False)]
lines += ["# Reward net (float32 kernels, fc3 weight gradient as 3xTF32 on the tensor cores) vs the float64 oracle", "",
          "`dmfg_rnet_backward` with random dL/dr ~ N(0,1): rewards handed back by the backward launch and the flat gradient, per "
          "parameter tensor.", ""]
for name, B, T, worst in rcases:
    lines += ["## %s" % name, "", "| output | max error |", "|---|---|"]
    lines += ["| %s | %.3e |" % (k, v) for k, v in worst.items()]
    lines.append("")
os.makedirs(os.path.dirname(a.out), exist_ok=True)
with open(a.out, "w") as f:
    f.write("\n".join(lines))
print("\n".join(lines))
