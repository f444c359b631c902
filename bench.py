#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 terrain hot path (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA path)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port, all host threads)

Workload (config.workload): fused slope + aspect + hillshade + curvature on a synthetic float32 DEM of
SIZE x SIZE pixels per GPU (default 32768 x 32768, resolution 5 m), i.e. `get_terrain_attribute(dem, ["slope",
"aspect","hillshade","curvature"], resolution=5, surface_fit=FIT)`.  At N>1 the raster is (N*SIZE) x SIZE, row-sharded,
and every step starts with the NCCL halo-row exchange between neighbouring shards (weak scaling).

One JSON line is printed by rank 0 (see the task contract): `value` is whole-job Mpixel/s with the DEM resident in
HBM; `e2e` is the same metric through the host-buffer C-ABI call (pinned host DEM in, pinned host planes out, copies
inside the timed region); `roofline` relates the kernel to the measured HBM copy bandwidth; `cpu_baseline` times the
oracle's C/OpenMP restatement of the reference's Numba engine on a bounded sample of the same workload.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ATTRS = ["slope", "aspect", "hillshade", "curvature"]
RESOLUTION = 5.0
METRIC = "Mpixel/s fused terrain attrs (slope+aspect+hillshade+curvature)"
ALGO_BYTES_PER_PIXEL = 4 + 4 * len(ATTRS)  # float32 in once + one float32 per attribute plane out


def _env_int(name: str, default: int) -> int:
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """Samples SM clocks / throttle reasons through NVML (nvidia_ml_py) in a thread during the timed region -- the same
    fields as the `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*` line of B200_PROFILING.md."""

    def __init__(self, gpu_index: int) -> None:
        self.gpu_index = gpu_index
        self.samples: list[tuple[float, int, int]] = []  # (time, sm_mhz, reasons bitmask)
        self.max_mhz: float | None = None
        self._stop = threading.Event()
        self.thread: threading.Thread | None = None
        self.t0 = self.t1 = 0.0

    def start(self) -> None:
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            return

        def loop() -> None:
            while not self._stop.is_set():
                try:
                    mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                    try:
                        rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:
                        rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    self.samples.append((time.perf_counter(), int(mhz), int(rs)))
                except Exception:
                    pass
                time.sleep(0.002)

        self.thread = threading.Thread(target=loop, daemon=True)
        self.thread.start()

    def mark_begin(self) -> None:
        self.t0 = time.perf_counter()

    def mark_end(self) -> None:
        self.t1 = time.perf_counter()

    def stop(self) -> dict:
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        inside = [s for s in self.samples if self.t0 <= s[0] <= self.t1] or self.samples[-3:]
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        reasons = sorted({n for _, _, rs in inside for bit, n in names.items() if rs & bit})
        mhz = sorted(m for _, m, _ in inside)
        return {"sm_mhz": float(mhz[len(mhz) // 2]) if mhz else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(inside), "source": "NVML (nvidia_ml_py), 2 ms period, timed region only"}


def measured_peak_gbs() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_per_launch(fit: str, size: int) -> float | None:
    """dram bytes per launch from the committed ncu capture of this same workload (profiles/terrain_traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "terrain_traffic.json")) as f:
            d = json.load(f)
        e = d.get(f"{fit.lower()}_{size}")
        return float(e["dram_bytes_per_launch"]) if e else None
    except (OSError, ValueError, KeyError):
        return None


def cpu_sample_dem(n: int):
    from oracle import synth

    return synth.fractal_dem((n, n), seed=42)


_BEST_THREADS: dict[str, int] = {}


def best_cpu_threads(fit: str, attrs: list[str]) -> int:
    """The C/OpenMP port scales with physical cores, not hyper-threads: time a 1024^2 pass with all logical CPUs and
    with half of them and keep the faster setting ("all the host threads it can use")."""
    if fit in _BEST_THREADS:
        return _BEST_THREADS[fit]
    from oracle import c_oracle

    dem = cpu_sample_dem(1024)
    n_all = os.cpu_count() or c_oracle.num_threads()
    best_t, best_n = float("inf"), n_all
    for n in sorted({n_all, max(1, n_all // 2)}, reverse=True):
        c_oracle.surface_attributes(dem, RESOLUTION, attrs, fit, degrees=True, clip_hillshade=True, nthreads=n)
        t0 = time.perf_counter()
        c_oracle.surface_attributes(dem, RESOLUTION, attrs, fit, degrees=True, clip_hillshade=True, nthreads=n)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best_t, best_n = dt, n
    _BEST_THREADS[fit] = best_n
    return best_n


def run_cpu_arm(sample: int, fit: str, repeats: int, attrs: list[str] | None = None) -> tuple[float, int, float]:
    """Times the oracle's C/OpenMP restatement of the reference's Numba engine on a sample^2 DEM (output buffers
    reused across passes, like a warmed-up NumPy/Numba caller).  Returns (Mpixel/s, threads, seconds per pass)."""
    from oracle import c_oracle

    attrs = attrs or ATTRS
    nthreads = best_cpu_threads(fit, attrs)
    dem = cpu_sample_dem(sample)
    best = float("inf")
    for _ in range(repeats + 1):  # first pass = warm-up (page faults of the output planes)
        t0 = time.perf_counter()
        c_oracle.surface_attributes(dem, RESOLUTION, attrs, fit, degrees=True, clip_hillshade=True, nthreads=nthreads)
        dt = time.perf_counter() - t0
        if _ > 0:
            best = min(best, dt)
    return sample * sample / best / 1e6, nthreads, best


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=_env_int("XB_BENCH_SIZE", 32768), help="rows and cols per GPU")
    ap.add_argument("--fit", default=os.environ.get("XB_BENCH_FIT", "Florinsky"),
                    choices=["Horn", "ZevenbergThorne", "Florinsky"])
    ap.add_argument("--cpu-sample", type=int, default=4096, help="edge of the CPU-baseline sample DEM")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    attrs = ATTRS if args.fit != "Horn" else ATTRS[:3]

    rank = _env_int("RANK", 0)
    world = _env_int("WORLD_SIZE", 1)
    local_rank = _env_int("LOCAL_RANK", 0)

    config = {
        "workload": f"{args.size}x{args.size} float32 synthetic DEM per GPU, fused {'+'.join(attrs)}, "
                    f"surface_fit={args.fit}, resolution={RESOLUTION}, degrees, hillshade clip",
        "rows_per_gpu": args.size, "cols": args.size, "attributes": attrs, "surface_fit": args.fit,
        "parallelism": f"row-shard x{world} + NCCL halo rows" if world > 1 else "single GPU",
        "l2": "inputs (4.3 GB/GPU) exceed L2 (126 MB); no flush needed",
    }

    # ------------------------------------------------------------------ reference arm (CPU) ----------------------
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import c_oracle

        threads = best_cpu_threads(args.fit, attrs)
        dem = cpu_sample_dem(args.cpu_sample)
        steps = max(1, args.steps)
        for _ in range(max(1, args.warmup)):
            c_oracle.surface_attributes(dem, RESOLUTION, attrs, args.fit, degrees=True, clip_hillshade=True,
                                        nthreads=threads)
        t0 = time.perf_counter()
        for _ in range(steps):
            c_oracle.surface_attributes(dem, RESOLUTION, attrs, args.fit, degrees=True, clip_hillshade=True,
                                        nthreads=threads)
        dt = time.perf_counter() - t0
        val = args.cpu_sample * args.cpu_sample * steps / dt / 1e6
        sample = f"{args.cpu_sample}x{args.cpu_sample} crop-sized DEM of the same generator per step"
        line = {
            "impl": "reference", "metric": METRIC, "value": val, "unit": "Mpixel/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 accumulate, f32 in/out", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": val, "unit": "Mpixel/s", "cores": threads, "kind": "port", "sample": sample,
                             "what": "oracle/terrain_oracle.c: C/OpenMP restatement of the reference's Numba engine "
                                     "(bit-exact vs the reference fixtures), all host threads"},
            "e2e": {"value": val, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------ CUDA arm ----------------------------------
    import torch
    import torch.distributed as dist

    from xdem_b200 import _engine, _lib
    from xdem_b200 import distributed as xbd

    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    depth = 2 if args.fit == "Florinsky" else 1
    rows, cols = args.size, args.size

    # synthetic DEM shard (+ halo rows filled by the exchange); same generator family as SURVEY 8d
    g = torch.Generator(device=dev).manual_seed(42 + rank)
    buf = torch.empty((rows + 2 * depth, cols), dtype=torch.float32, device=dev)
    core = buf[depth:depth + rows]
    chunk = 4096
    carry = torch.zeros((1, cols), dtype=torch.float32, device=dev)
    for r0 in range(0, rows, chunk):
        r1 = min(rows, r0 + chunk)
        n = torch.randn((r1 - r0, cols), generator=g, device=dev)
        blk = torch.cumsum(n, dim=0) + carry
        carry = blk[-1:].clone()
        core[r0:r1] = 1000.0 + 0.05 * torch.cumsum(blk, dim=1)
    del n, blk
    out = torch.empty((len(attrs), rows, cols), dtype=torch.float32, device=dev)
    shard = xbd.RowShard(rank, world, depth)
    kwargs = dict(surface_attributes=attrs, surface_fit=args.fit, degrees=True, clip_hillshade=True)

    def step() -> None:
        r_begin, r_end, view = shard.prepare(buf, rows)  # NCCL halo exchange (no-op at world=1)
        _engine.terrain_fused(view, RESOLUTION, row_begin=r_begin, row_end=r_end, out=out, **kwargs)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    sampler.mark_begin()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    sampler.mark_end()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - launches0
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else {}

    # kernel-only timing (no halo exchange) for the roofline, same stream, CUDA events
    kev0, kev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r_begin, r_end, view = shard.prepare(buf, rows)
    torch.cuda.synchronize()
    kev0.record()
    for _ in range(args.steps):
        _engine.terrain_fused(view, RESOLUTION, row_begin=r_begin, row_end=r_end, out=out, **kwargs)
    kev1.record()
    torch.cuda.synchronize()
    kern_ms = kev0.elapsed_time(kev1) / args.steps

    pixels_total = rows * cols * world
    ms_per_step = elapsed_ms / args.steps
    value = pixels_total / (ms_per_step * 1e-3) / 1e6
    peak, peak_src = measured_peak_gbs()
    bytes_per_px = 4 + 4 * len(attrs)
    achieved = rows * cols * bytes_per_px / (kern_ms * 1e-3) / 1e9
    traffic = ncu_traffic_per_launch(args.fit, args.size)

    # what the same traffic mix (one plane read, len(attrs) planes written, no arithmetic) reaches on this box: the
    # write-heavy mix does not attain the COPY bandwidth the contract's `peak` is (profiles/rw_mix_microbench_*.txt)
    stream_ceiling = None
    if len(attrs) <= 4 and (rows * cols) % 4 == 0:
        src = buf.reshape(-1)[: rows * cols]
        L = _lib.lib()
        sev0, sev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(L.xb_probe_stream(src.data_ptr(), out.data_ptr(), rows * cols, len(attrs), cur))
        torch.cuda.synchronize()
        sev0.record()
        for _ in range(3):
            _lib.check(L.xb_probe_stream(src.data_ptr(), out.data_ptr(), rows * cols, len(attrs), cur))
        sev1.record()
        torch.cuda.synchronize()
        probe_gbs = rows * cols * bytes_per_px / (sev0.elapsed_time(sev1) / 3 * 1e-3) / 1e9
        stream_ceiling = {"gbs": probe_gbs, "frac_of_ceiling": achieved / probe_gbs,
                          "what": "xb_probe_stream: trivial kernel, same bytes read/written per pixel, streaming stores"}

    # ------------------------------ e2e: host buffers through the C ABI (copies inside the timed region) ----------
    e2e = None
    if not args.no_e2e:
        del out
        torch.cuda.empty_cache()
        host_in = torch.empty((rows, cols), dtype=torch.float32, pin_memory=True)
        host_in.copy_(core)
        host_out = torch.empty((len(attrs), rows, cols), dtype=torch.float32, pin_memory=True)
        h_in, h_out = host_in.numpy(), host_out.numpy()
        _engine.terrain_fused_host(h_in, RESOLUTION, out=h_out, **kwargs)  # warm-up (scratch alloc)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            _engine.terrain_fused_host(h_in, RESOLUTION, out=h_out, **kwargs)
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {
            "value": pixels_total * args.e2e_steps / dt / 1e6, "unit": "Mpixel/s",
            "h2d_bytes_per_step": rows * cols * 4 * world, "d2h_bytes_per_step": rows * cols * 4 * len(attrs) * world,
            "steps": args.e2e_steps, "ms_per_step": dt / args.e2e_steps * 1e3,
            "call": "xb_terrain_fused_host (pinned host DEM -> pinned host planes, row-block streaming; at N>1 each "
                    "rank streams its own shard, no halo exchange needed because halo rows come from the host raster)",
        }
        del host_in, host_out

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        mpix, threads, sec = run_cpu_arm(args.cpu_sample, args.fit, 2, attrs)
        cpu = {"value": mpix, "unit": "Mpixel/s", "cores": threads, "kind": "port",
               "sample": f"{args.cpu_sample}x{args.cpu_sample} DEM of the same generator, best of 2 "
                         f"({sec:.2f} s per pass)",
               "what": "oracle/terrain_oracle.c (C/OpenMP restatement of the reference's Numba engine)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (exact-difference fp32 stencils; fp64 curvature algebra)",
            "data": "synthetic", "config": config, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": ("xbt::florinsky_sliding_kernel" if args.fit == "Florinsky" and len(attrs) > 3
                                    else "xbt::terrain_fused_kernel"),
                         "kernel_ms": kern_ms, "algorithmic_bytes_per_pixel": bytes_per_px,
                         "stream_ceiling": stream_ceiling},
            "e2e": e2e, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
