"""Fast (bracketed) vs exhaustive Nuth-Kaab iteration on a raster large enough for real sampling (stride > 1):
exact equality of every statistic, fallback count, and iteration timing."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from xdem_b200 import coreg
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
yy = torch.arange(S, device="cuda", dtype=torch.float32)[:, None]
xx = torch.arange(S, device="cuda", dtype=torch.float32)[None, :]
def surf(x, y):
    return 300 * torch.sin(x * 0.0037) * torch.cos(y * 0.0027) + 120 * torch.sin((x + 0.6 * y) * 0.0153) + 0.02 * x
ref = surf(xx, yy)
tba = surf(xx + 0.37, yy - 0.61) + 1.5 + 0.02 * torch.randn((S, S), device="cuda")
st = coreg._NKState(ref, tba, None)
print("eligible", st.fast_eligible(72))
bad = 0
for dx, dy in ((0, 0), (0.37, -0.61), (0.3, -0.5), (0.371, -0.612), (-1.2, 0.8), (0.37, -0.61)):
    res = st.iteration_fast(dx, dy, 72)
    print("stride", st.stride, "ns", st.ns, "gcap", st.gcap, "bcap", st.bcap, "flags", st.fast_last_flags,
          "compact", int(st.f_cnt[2]), int(st.f_cnt[3]))
    if res is None:
        bad += 1
        continue
    lo, hi, n_fin = st.compute_dh(dx, dy)
    med, cnt, _ = st.select_medians(0, 0.0, 0.0, 1.0, 1)
    mb, cb, mom = st.select_medians(1, float(med[0]), lo, hi, 72, want_moments=True)
    ok = (res["vshift"] == float(med[0]) and np.array_equal(res["median"], mb, equal_nan=True)
          and np.array_equal(res["counts"], cb) and res["n_fin"] == n_fin and res["lo"] == lo and res["hi"] == hi)
    print("equal:", ok)
    bad += not ok
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    st.iteration_fast(0.37, -0.61, 72)
torch.cuda.synchronize()
print("fast iteration ms", (time.perf_counter() - t0) * 100, "bad", bad)
