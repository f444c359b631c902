import sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
from xdem_b200 import spatialstats as xs, coreg, _lib
dev = torch.device('cuda')
def vario(N, S=32768, n_lags=50):
    g = torch.Generator(device=dev).manual_seed(44)
    lin = torch.randint(0, S * S, (int(N * 1.01),), generator=g, device=dev, dtype=torch.int64).unique()[:N]
    lin = lin[torch.randperm(lin.numel(), generator=g, device=dev)]
    x, y = lin % S, lin // S
    v = torch.randn(lin.numel(), generator=g, device=dev)
    maxlag = float(np.hypot(S - 1, S - 1) * 5.0)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e, cnt, ssq = xs.pairwise_lag_binning(x, y, v, None, 5.0, n_lags=n_lags, maxlag=maxlag)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    n = lin.numel(); pairs = n * (n - 1) / 2
    print(f"variogram N={n} even {n_lags} bins: {dt*1e3:.1f} ms  {pairs/dt/1e9:.1f} Gpairs/s  (sum cnt {int(cnt.sum())} of {int(pairs)})", flush=True)
def nk(size):
    from oracle import synth
    g = torch.Generator(device=dev).manual_seed(45)
    yy = torch.arange(size, device=dev, dtype=torch.float32)[:, None]; xx = torch.arange(size, device=dev, dtype=torch.float32)[None, :]
    def surf(dx, dy):
        z = torch.full((size, size), 1500.0, device=dev)
        rng = np.random.default_rng(45)
        for _ in range(12):
            kx, ky = rng.uniform(0.01, 0.12, 2) * rng.choice([-1, 1], 2); amp = rng.uniform(5, 40); ph = rng.uniform(0, 2*np.pi)
            z += amp * torch.sin(kx * (xx + dx) + ky * (yy + dy) + ph)
        return z
    ref = surf(0, 0); tba = surf(0.37, -0.61) + 1.5 + 0.01 * torch.randn((size, size), generator=g, device=dev)
    for rep in range(2):
        l0 = _lib.launch_count(); torch.cuda.synchronize(); t0 = time.perf_counter()
        (e, n, v), used = coreg.nuth_kaab(ref, tba, transform=(5.0, 0, 0, 0, -5.0, 0), tolerance=0.0, max_iterations=10, params_random={"subsample": 1.0})
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"NuthKaab {size}^2 dense 10 iterations: {dt*1e3:.1f} ms  {size*size*10/dt/1e6:.1f} Mpix*iter/s  shifts {e/5:.4f} {n/5:.4f} {v:.4f}  launches {_lib.launch_count()-l0}", flush=True)
vario(200_000); vario(1_000_000)
nk(4096); nk(16384)

def nk_pieces(size=16384):
    g = torch.Generator(device=dev).manual_seed(45)
    ref = 1500 + 30 * torch.sin(torch.arange(size, device=dev)[None, :] * 0.05) + 20 * torch.cos(torch.arange(size, device=dev)[:, None] * 0.031) + 0.05 * torch.randn((size, size), generator=g, device=dev)
    tba = torch.roll(ref, shifts=(0, 0), dims=(0, 1)) + 1.5 + 0.01 * torch.randn((size, size), generator=g, device=dev)
    st = coreg._NKState(ref.float(), tba.float(), None)
    def t(f, n=3):
        f(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): r = f()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3, r
    ms, (lo, hi, nf) = t(lambda: st.compute_dh(0.37, -0.61)); print(f"  dh pass            {ms:8.2f} ms")
    ms, (med, cnt, _) = t(lambda: st.select_medians(0, 0.0, 0.0, 1.0, 1)); print(f"  global median      {ms:8.2f} ms  ({med[0]:.5f})")
    ms, r = t(lambda: st.select_medians(1, float(med[0]), lo, hi, 72, want_moments=True)); print(f"  72-bin medians     {ms:8.2f} ms")
nk_pieces()
