"""Explores which (major-ness, LBO, SBO) conventions of the tcgen05 shared-memory matrix descriptor reproduce A.B."""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from discrete_mean_field_game_b200 import engine as eng
np.set_printoptions(linewidth=250, precision=2, suppress=True)
dev = torch.device("cuda:0")
rng = np.random.RandomState(0)
A = np.float32(rng.randint(-4, 5, size=(128, 8)))          # exactly representable in tf32
B = np.float32(rng.randint(-4, 5, size=(8, 16)))
ref = A @ B
At, Bt = torch.as_tensor(A, device=dev), torch.as_tensor(B, device=dev)
out = eng.umma_probe(At, Bt, (-1, 128, 256), (0, 128, 256)).cpu().numpy()
print("TMEM st/ld round trip (expect 1000 + lane + col/100):"); print(out[:3, :6]); print(out[64:66, 26:32])
out = eng.umma_probe(At, Bt, (0, 128, 256), (0, 128, 256)).cpu().numpy()
print("after K/K MMA: rows 0..2"); print(out[:3]); print("ref rows 0..2"); print(ref[:3])
print("rows 64..65"); print(out[64:66])
cfgs = {
    "K": [(0, 128, 256), (0, 256, 128)],
    "MN": [(1, 4608, 144), (1, 4096, 128), (1, 512, 128), (1, 128, 512)],
}
for (an, al), (bn, bl) in itertools.product(cfgs.items(), cfgs.items()):
    for a in al:
        for b in bl:
            for swap_a, swap_b in itertools.product((False, True), (False, True)):
                ad = (a[2], a[1]) if swap_a else (a[1], a[2])
                bd = (b[2], b[1]) if swap_b else (b[1], b[2])
                out = eng.umma_probe(At, Bt, a, b, ad, bd).cpu().numpy()[:, :16]
                ok = np.array_equal(out, ref)
                print("A %s place%s desc(lbo,sbo)=%s  B %s place%s desc=%s -> %s (nonzeros %d, max|err| %.1f)" % (
                    an, a[1:], ad, bn, b[1:], bd, "MATCH" if ok else "no", int(np.count_nonzero(out)), float(np.abs(out - ref).max())))
