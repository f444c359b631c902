"""A/B of the variogram tile sweeps (xb_set_option("variogram_full_tiles", mask)) at two sample counts."""
import sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from xdem_b200 import spatialstats as xs, _lib

dev = torch.device("cuda")
S = 32768
for N in (200_000, 1_000_000):
    g = torch.Generator(device=dev).manual_seed(44)
    lin = torch.randint(0, S * S, (int(N * 1.01),), generator=g, device=dev, dtype=torch.int64).unique()[:N]
    lin = lin[torch.randperm(lin.numel(), generator=g, device=dev)]
    x, y = lin % S, lin // S
    v = torch.randn(lin.numel(), generator=g, device=dev)
    maxlag = float(np.hypot(S - 1, S - 1) * 5.0)
    base = None
    for n_lags in (50, 10):
        for mask in (0, 1, 2, 4, 3, 7, 0, 7):
            _lib.set_option("variogram_full_tiles", mask)
            ts = []
            for rep in range(3):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                e, cnt, ssq = xs.pairwise_lag_binning(x, y, v, None, 5.0, n_lags=n_lags, maxlag=maxlag)
                torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
            if mask == 0 and base is None:
                base = (cnt.copy(), ssq.copy())
            same = np.array_equal(cnt, base[0]) if n_lags == 50 else True
            rel = float(np.max(np.abs(ssq - base[1]) / np.maximum(base[1], 1e-30))) if n_lags == 50 else 0.0
            print(f"N={N} bins={n_lags} mask={mask}: {min(ts)*1e3:8.1f} ms (reps {[round(t*1e3,1) for t in ts]}) counts_equal={same} max_rel_sum_diff={rel:.2e}", flush=True)
        base = None
_lib.set_option("variogram_full_tiles", 7)
