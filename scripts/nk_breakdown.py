"""Wall-clock breakdown of one dense Nuth-Kaab fit at 16384^2 (bench c5 inputs): setup vs iterations vs host fit."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from xdem_b200 import coreg

S = 16384
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

def sync():
    torch.cuda.synchronize(); return time.perf_counter()

for rep in range(3):
    t0 = sync()
    st = coreg._NKState(ref, tba, None)
    t1 = sync()
    nv = st.n_valid()
    t2 = sync()
    ok = st.fast_eligible(72)
    t3 = sync()
    its, fits = [], []
    off = (0.0, 0.0)
    for i in range(10):
        a = sync()
        res = st.iteration_fast(off[0], off[1], 72)
        b = sync()
        e, n, _ = coreg._fit_from_bins(res["median"], res["moments"], res["lo"], res["hi"], 72, coreg.scipy.optimize.curve_fit)
        c = time.perf_counter()
        off = (off[0] + e, off[1] - n) if False else (off[0] + e / 1.0, off[1] + n / -1.0)
        its.append((b - a) * 1e3); fits.append((c - b) * 1e3)
    t4 = sync()
    print(f"rep {rep}: state {1e3*(t1-t0):.2f} ms, n_valid {1e3*(t2-t1):.2f}, eligible/setup {1e3*(t3-t2):.2f}, "
          f"10 iterations {1e3*(t4-t3):.2f} (gpu+sync {sum(its):.2f}: {' '.join('%.2f' % v for v in its)}; fit {sum(fits):.2f}) "
          f"total {1e3*(t4-t0):.2f}", flush=True)
    del st
t0 = sync()
r = coreg.nuth_kaab(ref, tba, transform=(5.0, 0, 0, 0, -5.0, 0), tolerance=0.0, max_iterations=10, params_random={"subsample": 1.0})
print("whole nuth_kaab", 1e3 * (sync() - t0), r)
