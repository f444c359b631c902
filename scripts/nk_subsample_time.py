"""Nuth-Kaab with the reference's DEFAULT subsample (5e5 points, affine.py:2405) on the 16384^2 bench pair."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from xdem_b200 import coreg

S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda")
yy = torch.arange(S, device=dev, dtype=torch.float32)[:, None]
xx = torch.arange(S, device=dev, dtype=torch.float32)[None, :]
def surf(dx, dy):
    z = torch.full((S, S), 1500.0, device=dev)
    rng = np.random.default_rng(45)
    for _ in range(12):
        kx, ky = rng.uniform(0.01, 0.12, 2) * rng.choice([-1, 1], 2)
        amp, ph = rng.uniform(5, 40), rng.uniform(0, 2 * np.pi)
        z += float(amp) * torch.sin(float(kx) * (xx + dx) + float(ky) * (yy + dy) + float(ph))
    return z
g = torch.Generator(device=dev).manual_seed(46)
ref = surf(0.0, 0.0)
tba = surf(0.37, -0.61) + 1.5 + 0.01 * torch.randn((S, S), generator=g, device=dev)
for sub in (5e5, 1.0):
    for rep in range(7):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = coreg.nuth_kaab(ref, tba, transform=(5.0, 0, 0, 0, -5.0, 0), tolerance=0.0, max_iterations=10,
                            params_random={"subsample": sub, "random_state": 42})
        torch.cuda.synchronize()
        print(f"subsample={sub:g} rep {rep}: {1e3 * (time.perf_counter() - t0):8.2f} ms  {r}", flush=True)
