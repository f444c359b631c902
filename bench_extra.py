#!/usr/bin/env python
"""bench_extra.py -- the other two BASELINE.json workloads of the hot path, one JSON line each (not the driver's headline;
the lines are committed under profiles/):

    python bench_extra.py variogram [--n 1000000] [--cpu-n 40000]     # configs[2]: 1e6 samples, 50 lag bins
    python bench_extra.py nuthkaab  [--size 16384] [--cpu-size 1024]  # configs[4]: 10 dense iterations

Each line carries the GPU figure (CUDA events / wall clock around the public API), the bound it is compared against and a
`cpu_baseline` from the oracle (C/OpenMP all-pairs binning; NumPy/SciPy restatement of the reference's iteration)."""

from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _dist_setup():
    """(rank, world, device); initialises NCCL when launched by torchrun."""
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(
        os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    return rank, world, dev


def _max_over_ranks(dt: float, dev) -> float:
    import torch
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return dt


def bench_variogram(args) -> dict:
    import torch

    from oracle import c_oracle
    from xdem_b200 import _lib, spatialstats as xs

    rank, world, dev = _dist_setup()
    S, gsd, n_lags = 32768, 5.0, 50
    g = torch.Generator(device=dev).manual_seed(44)
    lin = torch.randint(0, S * S, (int(args.n * 1.01),), generator=g, device=dev, dtype=torch.int64).unique()
    lin = lin[torch.randperm(lin.numel(), generator=g, device=dev)][: args.n]
    x, y = lin % S, lin // S
    v = torch.randn(lin.numel(), generator=g, device=dev)
    maxlag = float(np.hypot(S - 1, S - 1) * gsd)
    n = lin.numel()
    pairs = n * (n - 1) // 2
    times = []
    t_gpu0 = time.perf_counter()
    for rep in range(args.steps + 1):
        l0 = _lib.launch_count()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        edges, cnt, ssq = xs.pairwise_lag_binning(x, y, v, None, gsd, n_lags=n_lags, maxlag=maxlag,
                                                  distributed=world > 1)
        torch.cuda.synchronize()
        times.append(_max_over_ranks(time.perf_counter() - t0, dev))
        launches = _lib.launch_count() - l0
    t_gpu1 = time.perf_counter()
    dt = min(times[1:])
    assert int(cnt.sum()) == pairs - 1  # every pair binned once; the farthest pair sits on the last (open) edge
    if rank != 0:
        return {}
    # CPU baseline: plain-C all-pairs binning (oracle), all threads, bounded N
    m = args.cpu_n
    xc = np.stack([(x[:m] * gsd).cpu().numpy().astype(np.float64), (y[:m] * gsd).cpu().numpy().astype(np.float64)], 1)
    vc = v[:m].cpu().numpy().astype(np.float64)
    c_oracle.variogram_pairs(xc[:2000], vc[:2000], edges)
    t0 = time.perf_counter()
    c_oracle.variogram_pairs(xc, vc, edges)
    tc = time.perf_counter() - t0
    sm_clock, sms = 1.965e9, 148
    issue_peak = sms * 4 * 32 * sm_clock  # thread-instructions / s
    return {
        "timed_window": [t_gpu0, t_gpu1],
        "metric": "Gpairs/s all-pairs empirical variogram (Matheron, 50 even lag bins)", "value": pairs / dt / 1e9,
        "unit": "Gpairs/s", "n_gpus": world, "steps": args.steps, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "dtype": "u32 distances / f32 diffs, u64 counts, f64 sums", "data": "synthetic",
        "config": {"workload": f"sample_empirical_variogram core: {n} random samples of a 32768^2 grid, gsd 5, 50 even "
                               f"bins, includes Morton sort + max-distance pass + binning ({launches} kernel launches)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "issue (no HBM traffic: 16 MB of samples stay in L2)",
                     "achieved": pairs / dt * 13 / 1e12, "peak": world * issue_peak / 1e12, "unit": "T thread-instr/s",
                     "frac": pairs / dt * 13 / (world * issue_peak),
                     "note": "13 issue slots per pair (DESIGN.md K2) x pairs/s vs 148 SM x 4 x 32 lanes x 1.965 GHz"},
        "cpu_baseline": {"value": m * (m - 1) / 2 / tc / 1e9, "unit": "Gpairs/s", "cores": c_oracle.num_threads(),
                         "kind": "port", "sample": f"first {m} of the same samples ({m*(m-1)//2:.3e} pairs, {tc:.1f} s)",
                         "what": "oracle/variogram_oracle.c (restated scikit-gstat pairwise binning, C/OpenMP)"},
    }


def bench_nuthkaab(args) -> dict:
    import torch

    from oracle import nk_oracle
    from xdem_b200 import _lib, coreg

    rank, world, dev = _dist_setup()
    size = args.size
    rows_local = size // world
    row0 = rank * rows_local

    def surf(n, dx, dy, device):
        yy = (row0 + torch.arange(rows_local if n == size else n, device=device, dtype=torch.float32))[:, None]
        xx = torch.arange(n, device=device, dtype=torch.float32)[None, :]
        z = torch.full((yy.shape[0], n), 1500.0, device=device)
        rng = np.random.default_rng(45)
        for _ in range(12):
            kx, ky = rng.uniform(0.01, 0.12, 2) * rng.choice([-1, 1], 2)
            amp, ph = rng.uniform(5, 40), rng.uniform(0, 2 * np.pi)
            z += float(amp) * torch.sin(float(kx) * (xx + dx) + float(ky) * (yy + dy) + float(ph))
        return z

    g = torch.Generator(device=dev).manual_seed(46)
    ref = surf(size, 0.0, 0.0, dev)
    tba = surf(size, 0.37, -0.61, dev) + 1.5 + 0.01 * torch.randn(tuple(ref.shape), generator=g, device=dev)
    times = []
    t_gpu0 = time.perf_counter()
    for rep in range(args.steps + 1):
        l0 = _lib.launch_count()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if world > 1:
            from xdem_b200 import distributed as xbd

            (e, n, vz), used = xbd.sharded_nuth_kaab(ref, tba, transform=(5.0, 0, 0, 0, -5.0, 0), tolerance=0.0,
                                                     max_iterations=10)
        else:
            (e, n, vz), used = coreg.nuth_kaab(ref, tba, transform=(5.0, 0, 0, 0, -5.0, 0), tolerance=0.0,
                                               max_iterations=10, params_random={"subsample": 1.0})
        torch.cuda.synchronize()
        times.append(_max_over_ranks(time.perf_counter() - t0, dev))
        launches = _lib.launch_count() - l0
    t_gpu1 = time.perf_counter()
    dt = min(times[1:])
    assert abs(e / 5 + 0.37) < 2e-3 and abs(n / 5 + 0.61) < 2e-3 and abs(vz + 1.5) < 2e-3, (e, n, vz)
    # the reference's DEFAULT configuration, NuthKaab(subsample=5e5) (affine.py:2405): a point-list fit
    default_sub = None
    if world == 1:
        ts = []
        for rep in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            (e2, n2, vz2), used2 = coreg.nuth_kaab(ref, tba, transform=(5.0, 0, 0, 0, -5.0, 0), tolerance=0.0,
                                                   max_iterations=10,
                                                   params_random={"subsample": 5e5, "random_state": 42})
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        default_sub = {"subsample": 500000, "ms_per_fit": min(ts[1:]) * 1e3, "points_used": int(used2),
                       "recovered_shift_px": [-e2 / 5, -n2 / 5], "dz": vz2,
                       "what": "same pair, the reference's default 5e5-point subsample, 10 iterations: preparation "
                               "pass over the rasters once, then every iteration touches the points only"}
    if rank != 0:
        return {}
    # CPU baseline: NumPy/SciPy restatement of the reference's iteration (same code path as xdem on a CPU)
    cs = min(args.cpu_size, rows_local)
    rc, tc_ = ref[:cs, :cs].cpu().numpy(), tba[:cs, :cs].cpu().numpy()
    t0 = time.perf_counter()
    nk_oracle.nuth_kaab(rc, tc_, None, (5.0, -5.0), 0.0, 10)
    tcpu = time.perf_counter() - t0
    peak = 6481.1
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    algo = (12 + 10 * 16) * size * size  # aux once + 16 B/px/iteration
    return {
        "timed_window": [t_gpu0, t_gpu1],
        "metric": "Mpixel*iteration/s Nuth-Kaab (dense, 10 iterations)", "value": size * size * 10 / dt / 1e6,
        "unit": "Mpixel*iter/s", "n_gpus": world, "steps": args.steps, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong",
        "dtype": "f32 rasters, f64 interpolation/moments, exact medians", "data": "synthetic",
        "config": {"workload": f"NuthKaab {size}^2 ref/tba pair, subsample=1, 10 iterations, 72 aspect bins, host "
                               f"curve_fit ({launches} kernel launches); recovered shift px "
                               f"({-e/5:.4f}, {-n/5:.4f}), dz {vz:.4f}"},
        "gpu_launches": int(launches),
        "default_subsample": default_sub,
        "roofline": {"bound": "hbm", "achieved": algo / dt / 1e9, "peak": peak * world, "unit": "GB/s",
                     "frac": algo / dt / 1e9 / (peak * world),
                     "note": "algorithmic minimum 12 B/px (aux) + 16 B/px/iteration; the bracketed exact selection "
                             "streams two full passes (dh 13 B/px, y 9 B/px) + a 4 M-pixel sample per iteration, "
                             "host curve_fit included in the time (DESIGN.md K3)"},
        "cpu_baseline": {"value": cs * cs * 10 / tcpu / 1e6, "unit": "Mpixel*iter/s", "cores": 1, "kind": "port",
                         "sample": f"{cs}^2 crop of the same pair, 10 iterations ({tcpu:.1f} s)",
                         "what": "oracle/nk_oracle.py (NumPy/SciPy restatement of affine.py:477-609, pinned to reference "
                                 "fixtures; the reference itself is single-threaded NumPy here)"},
    }


def bench_texture(args) -> dict:
    """Texture shading: prepare -> rfft2 -> filter -> irfft2 -> finish on one GPU (the FFT is global:
    replicas only, no row sharding)."""
    import torch

    from oracle import terrain_oracle as to
    from xdem_b200 import _lib, freq

    dev = torch.device("cuda", 0)
    size = args.size
    g = torch.Generator(device=dev).manual_seed(47)
    z = torch.randn((size, size), generator=g, device=dev)
    z = torch.cumsum(z, 0)
    z = (1000.0 + 0.05 * torch.cumsum(z, 1)).float()
    z[100:110, 200:230] = float("nan")
    times = []
    for rep in range(args.steps + 1):
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        out = freq.texture_shading_device(z, 0.8)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e-3)
        launches = _lib.launch_count() - l0
        del out
    dt = min(times[1:])
    cs = min(args.cpu_size * 4, size)
    zc = z[:cs, :cs].cpu().numpy()
    to.texture_shading(zc[:256, :256], 0.8)
    t0 = time.perf_counter()
    to.texture_shading(zc, 0.8)
    tcpu = time.perf_counter() - t0
    peak = 6481.1
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    # algorithmic traffic of the five stages, float32: stats 4, pad 4+4, rfft2 >= 4+8*(1/2)*2.., counted as one read and
    # one write of every array each stage touches: 4 + 8 + (4+4) + (4+4) + (4+4) + (8+4) = 48 B/px
    algo = 48.0 * size * size
    return {
        "metric": "Mpixel/s texture shading (alpha 0.8)", "value": size * size / dt / 1e6, "unit": "Mpixel/s",
        "n_gpus": 1, "steps": args.steps, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "replicas only",
        "dtype": "f32 raster, complex64 spectrum, f64 filter", "data": "synthetic",
        "config": {"workload": f"texture_shading {size}^2 float32 fractal DEM with a NaN hole, alpha 0.8 "
                               f"({launches} library kernels of ours + 2 cuFFT transforms per step)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": algo / dt / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": algo / dt / 1e9 / peak,
                     "note": "48 B/px = one read + one write per array per stage if each FFT were a single pass; cuFFT "
                             "makes several passes over the 2-D data, so frac is a lower bound on its efficiency"},
        "cpu_baseline": {"value": cs * cs / tcpu / 1e6, "unit": "Mpixel/s", "cores": 1, "kind": "port",
                         "sample": f"{cs}^2 crop of the same DEM ({tcpu:.1f} s)",
                         "what": "oracle/terrain_oracle.py:texture_shading (NumPy + scipy.fft restatement of "
                                 "freq.py:62-148, pinned to reference fixtures)"},
    }


def bench_binning(args) -> dict:
    """nd_binning core: count + exact median + NMAD of a dh-like raster in 100 bins of one terrain
    variable (TerrainBias' default binning, biascorr.py:467-473) -- bin numbers + 2 x 4 radix-select passes."""
    import torch

    from oracle import binning_oracle as bo
    from xdem_b200 import _lib, binning as xb

    dev = torch.device("cuda", 0)
    n = args.size * args.size
    g = torch.Generator(device=dev).manual_seed(48)
    var = torch.randn(n, generator=g, device=dev) * 1.5
    vals = 0.3 * var + torch.randn(n, generator=g, device=dev) * 0.5
    edges = xb.bin_edges(float(var.min()), float(var.max()), 100, np.float32)
    times = []
    for rep in range(args.steps + 1):
        l0 = _lib.launch_count()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = xb.binned_robust_stats(vals, [var], [edges], want_nmad=True)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        launches = _lib.launch_count() - l0
    dt = min(times[1:])
    assert int(st["count"].sum()) == n
    m = min(n, args.cpu_size * args.cpu_size * 4)
    vc, xc = vals[:m].cpu().numpy(), var[:m].cpu().numpy()
    t0 = time.perf_counter()
    bo.nd_binning(vc, [xc], [100])
    tcpu = time.perf_counter() - t0
    peak = 6481.1
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    algo = (8 + 6) * n + 2 * 4 * 6.0 * n  # keys pass (8 B read, 6 B written) + 8 select passes over 6-byte pairs
    return {
        "metric": "Msamples/s binned count + median + NMAD (100 bins, 1 variable)", "value": n / dt / 1e6,
        "unit": "Msample/s", "n_gpus": 1, "steps": args.steps, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "replicas only", "dtype": "f32 values, u32 keys, u16 bin numbers, exact selects", "data": "synthetic",
        "config": {"workload": f"{n} samples ({args.size}^2 raster), 100 linspace bins, statistics count / nanmedian / "
                               f"nmad ({launches} kernel launches, 10 host round trips for the radix digits)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": algo / dt / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": algo / dt / 1e9 / peak,
                     "note": "14 B/sample for the key pass + 8 radix passes x 6 B/sample (+ absdev re-key 10 B)"},
        "cpu_baseline": {"value": m / tcpu / 1e6, "unit": "Msample/s", "cores": 1, "kind": "port",
                         "sample": f"first {m} samples ({tcpu:.1f} s)",
                         "what": "oracle/binning_oracle.py (scipy.stats.binned_statistic_dd with np.nanmedian and nmad, the "
                                 "calls the reference makes)"},
    }


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", choices=["variogram", "nuthkaab", "texture", "binning"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--cpu-n", type=int, default=40_000)
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--cpu-size", type=int, default=1024)
    args = ap.parse_args()
    line = {"variogram": bench_variogram, "nuthkaab": bench_nuthkaab, "texture": bench_texture,
            "binning": bench_binning}[args.workload](args)
    if line:
        print(json.dumps(line), flush=True)
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
