#!/usr/bin/env python
"""bench.py -- benchmark of the B200 hot path against BASELINE.json.

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA path)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port, all host threads)

Headline (the one JSON line's top-level keys) = BASELINE.json configs[1]: fused slope + aspect + hillshade + curvature on
a synthetic float32 DEM of SIZE x SIZE pixels per GPU (default 32768 x 32768, resolution 5 m), i.e.
`get_terrain_attribute(dem, ["slope","aspect","hillshade","curvature"], resolution=5, surface_fit=FIT)`.  At N>1 the
raster is (N*SIZE) x SIZE, row-sharded, and every step contains the NCCL halo-row exchange (weak scaling).

  value      whole-job Mpixel/s with the DEM resident in HBM (CUDA events, barrier + sync on both sides, max over ranks)
  roofline   the dominant kernel against the measured HBM copy bandwidth; kernel_ms is timed with CUDA events around the
             kernel launches INSIDE the timed steps
  e2e        the same metric through the public API a reference caller uses -- `xdem_b200.terrain.get_terrain_attribute`
             on the caller's (pageable) NumPy raster, host<->device copies inside the timed region; `e2e.pinned` is the
             same request through the C ABI with page-locked buffers on both sides
  cpu_baseline  oracle/terrain_oracle.c (C/OpenMP port of the reference's Numba engine) on a bounded sample

`extra` carries the other BASELINE configs so that the driver sees them in the same line (each with its own roofline
and cpu_baseline; a failure in one of them is reported as {"error": ...} and never hides the headline):
  c1  configs[0]: Horn slope, 4096^2 (device-resident, through the API, and the CPU arm at the full size)
  c3  configs[2]: all-pairs variogram, 1e6 random samples of a 32768^2 grid, 50 lag bins
  c4  configs[3]: ONE 65536^2 raster, row-sharded over the N GPUs (strong scaling), all 13 stencil attributes of
      terrain.py:41-57 (9 surface-fit + TPI, TRI, roughness, rugosity), NCCL halo exchange overlapped inside the step
  c5  configs[4]: Nuth-Kaab, 16384^2 pair, 10 dense iterations, row-sharded at N>1
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
import traceback

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ATTRS = ["slope", "aspect", "hillshade", "curvature"]
SURF9 = ["slope", "aspect", "hillshade", "profile_curvature", "tangential_curvature", "planform_curvature",
         "flowline_curvature", "max_curvature", "min_curvature"]
WIN4 = ["topographic_position_index", "terrain_ruggedness_index", "roughness", "rugosity"]
RESOLUTION = 5.0
METRIC = "Mpixel/s fused terrain attrs (slope+aspect+hillshade+curvature)"


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
        self.active = False

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
                if not self.active:  # NVML queries take driver locks: only poll inside the marked timed regions
                    time.sleep(0.005)
                    continue
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
        self.active = True
        self.t0 = time.perf_counter()

    def mark_end(self) -> None:
        self.t1 = time.perf_counter()
        self.active = False

    def window(self, t0: float, t1: float) -> dict:
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-3:]
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        reasons = sorted({n for _, _, rs in inside for bit, n in names.items() if rs & bit})
        mhz = sorted(m for _, m, _ in inside)
        return {"sm_mhz": float(mhz[len(mhz) // 2]) if mhz else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(inside), "source": "NVML (nvidia_ml_py), 2 ms period, timed region only"}

    def stop(self) -> dict:
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        return self.window(self.t0, self.t1)


def measured_peak_gbs() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_per_launch(key: str) -> tuple[float | None, str | None]:
    """dram bytes per launch from the committed ncu capture of this same workload (profiles/terrain_traffic.json):
    ncu cannot run inside a timed benchmark, so this figure is READ from profiles/, not measured in this run."""
    try:
        with open(os.path.join(ROOT, "profiles", "terrain_traffic.json")) as f:
            d = json.load(f)
        e = d.get(key)
        if e:
            return float(e["dram_bytes_per_launch"]), f"read from profiles/terrain_traffic.json ({e.get('source', '?')})"
    except (OSError, ValueError, KeyError):
        pass
    return None, None


# ---------------------------------------------------------------------------------------------------------------
# CPU arms (oracle/: C/OpenMP port of the reference's Numba engine; the only place bench.py executes oracle code)
# ---------------------------------------------------------------------------------------------------------------


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


def cpu_terrain_pass(dem, fit: str, surf: list[str], win: list[str], nthreads: int) -> None:
    from oracle import c_oracle

    if surf:
        c_oracle.surface_attributes(dem, RESOLUTION, surf, fit, degrees=True, clip_hillshade=True, nthreads=nthreads)
    w3 = [a for a in win if a != "rugosity"]
    if w3:
        c_oracle.windowed_indexes(dem, 3, w3, nthreads=nthreads)
    if "rugosity" in win:
        c_oracle.rugosity(dem, RESOLUTION, nthreads=nthreads)


def run_cpu_terrain(sample: int, fit: str, surf: list[str], win: list[str], repeats: int = 2) -> dict:
    """Best-of-`repeats` wall time of the C/OpenMP port on a sample^2 DEM of the bench generator family (first pass =
    warm-up of the output pages)."""
    nthreads = best_cpu_threads(fit, surf or ["slope"])
    dem = cpu_sample_dem(sample)
    best = float("inf")
    for i in range(repeats + 1):
        t0 = time.perf_counter()
        cpu_terrain_pass(dem, fit, surf, win, nthreads)
        dt = time.perf_counter() - t0
        if i > 0:
            best = min(best, dt)
    return {"value": sample * sample / best / 1e6, "unit": "Mpixel/s", "cores": nthreads, "kind": "port",
            "sample": f"{sample}x{sample} DEM of the same generator family, best of {repeats} ({best:.2f} s per pass)",
            "what": "oracle/terrain_oracle.c (C/OpenMP restatement of the reference's Numba engine, bit-exact vs the "
                    "reference fixtures for the surface attributes and TPI/TRI/roughness)"}


# ---------------------------------------------------------------------------------------------------------------
# helpers of the CUDA arm
# ---------------------------------------------------------------------------------------------------------------


class Ctx:
    def __init__(self) -> None:
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = _env_int("RANK", 0)
        self.world = _env_int("WORLD_SIZE", 1)
        self.local_rank = _env_int("LOCAL_RANK", 0)
        assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device"
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self) -> None:
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, step, steps: int, warmup: int, kernel_events: list | None = None) -> tuple[float, float, float]:
        """W untimed + exactly K timed steps bracketed by barrier + synchronize on both sides; CUDA events on the
        current stream; max over ranks.  Returns (elapsed ms, wall t0, wall t1)."""
        torch = self.torch
        for _ in range(warmup):
            step()
        if kernel_events is not None:
            kernel_events.clear()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ms = self.max_over_ranks(e0.elapsed_time(e1))
        self.barrier()
        return ms, t0, t1


def bind_to_gpu_numa(local_rank: int) -> str:
    """Pin this process to the CPUs NVML reports as local to its GPU before any pinned allocation (first touch places
    the pages on that NUMA node): the e2e path moves ~4x the DEM size through host memory per step."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {i * 64 + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} CPUs local to GPU {local_rank} (NVML affinity)"
    except Exception as e:  # noqa: BLE001
        return f"not bound ({type(e).__name__})"
    return "not bound"


# ---------------------------------------------------------------------------------------------------------------
# extras: the other BASELINE configs
# ---------------------------------------------------------------------------------------------------------------


def extra_c1(ctx: Ctx, args, peak: float) -> dict:
    """configs[0]: Horn slope on a 4096^2 float32 DEM."""
    import numpy as np

    import bench_data
    import xdem_b200
    from xdem_b200 import _engine

    torch = ctx.torch
    S = 4096
    z = bench_data.device_fractal_dem(S, S, 42, ctx.dev)
    out = torch.empty((1, S, S), dtype=torch.float32, device=ctx.dev)
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=ctx.dev)  # > 126 MB L2
    n = max(ctx_steps(args), 10)
    evs = []
    for i in range(n + 3):
        flush.fill_(i & 255)  # the 64 MB raster would otherwise stay L2-resident between steps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _engine.terrain_fused(z, RESOLUTION, ["slope"], [], surface_fit="Horn", degrees=True, out=out)
        e1.record()
        if i >= 3:
            evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
    res: dict = {"config": {"workload": "4096x4096 float32 DEM, xdem.terrain.slope(surface_fit='Horn'), resolution 5; L2 "
                                       "flushed (160 MB fill) between device-resident steps"},
                 "metric": "Mpixel/s Horn slope", "unit": "Mpixel/s", "value": S * S / ms / 1e3, "ms_per_step": ms,
                 "steps": len(evs), "n_gpus": 1,
                 "roofline": {"bound": "hbm", "achieved": S * S * 8 / ms / 1e6, "peak": peak, "unit": "GB/s",
                              "frac": S * S * 8 / ms / 1e6 / peak, "kernel": "xbt::terrain_fused_kernel<float,1,0,...,1>",
                              "algorithmic_bytes_per_pixel": 8,
                              "note": "a 0.03 ms kernel: launch latency and the tail of the last wave are visible"}}
    if ctx.rank == 0:
        dem_np = z.cpu().numpy()
        xdem_b200.terrain.slope(dem_np, surface_fit="Horn", resolution=RESOLUTION)
        t0 = time.perf_counter()
        for _ in range(5):
            xdem_b200.terrain.slope(dem_np, surface_fit="Horn", resolution=RESOLUTION)
        dt = (time.perf_counter() - t0) / 5
        res["e2e"] = {"value": S * S / dt / 1e6, "unit": "Mpixel/s", "ms_per_step": dt * 1e3,
                      "h2d_bytes_per_step": S * S * 4, "d2h_bytes_per_step": S * S * 4,
                      "call": "xdem_b200.terrain.slope(ndarray, surface_fit='Horn')"}
        if not args.no_cpu:
            cpu = run_cpu_terrain(S, "Horn", ["slope"], [], repeats=3)
            cpu["sample"] = "the full 4096x4096 config, " + cpu["sample"].split(", ", 1)[1]
            res["cpu_baseline"] = cpu
    del z, out, flush
    return res


def ctx_steps(args) -> int:
    return max(1, min(args.steps, 20))


def extra_c4(ctx: Ctx, args, peak: float, sampler: ClockSampler | None) -> dict:
    """configs[3]: one 65536^2 raster, all 13 stencil attributes, row-sharded over the ranks (strong scaling)."""
    import bench_data
    from xdem_b200 import _engine, _lib
    from xdem_b200 import distributed as xbd

    torch = ctx.torch
    S = args.c4_size
    world, rank = ctx.world, ctx.rank
    rows = S // world
    depth = 2  # Florinsky 5x5 surface fit (terrain.py:417-432)
    buf = torch.empty((rows + 2 * depth, S), dtype=torch.float32, device=ctx.dev)
    bench_data.device_fractal_dem(rows, S, 142 + rank, ctx.dev, out=buf[depth:depth + rows])
    free, _ = torch.cuda.mem_get_info()
    plane_bytes = 13 * rows * S * 4
    n_pass = 1
    while plane_bytes / n_pass > free - (6 << 30):
        n_pass *= 2
    rows_pass = rows // n_pass
    out = torch.empty((13, rows_pass, S), dtype=torch.float32, device=ctx.dev)
    shard = xbd.RowShard(rank, world, depth)
    kw = dict(surface_fit="Florinsky", degrees=True, clip_hillshade=True)

    def launch(view, rb: int, re: int, out_row0: int) -> None:
        # N = 1: the 223 GB of planes do not fit next to the raster; the shard is processed in `n_pass` row passes into
        # the same plane buffer (every pass computes and writes its planes; earlier passes are overwritten)
        r = rb
        while r < re:
            p = (out_row0 + (r - rb)) // rows_pass
            stop = min(re, rb - out_row0 + (p + 1) * rows_pass)
            o0 = out_row0 + (r - rb) - p * rows_pass
            _engine.terrain_fused(view, RESOLUTION, SURF9, WIN4, row_begin=r, row_end=stop,
                                  out=out[:, o0:o0 + (stop - r)], **kw)
            r = stop

    def step() -> None:
        shard.run_overlapped(buf, rows, launch)

    l0 = _lib.launch_count()
    steps, warm = max(1, min(args.steps, 5)), 3
    if sampler is not None:
        sampler.mark_begin()
    ms, t0, t1 = ctx.timed(step, steps, warm)
    if sampler is not None:
        sampler.mark_end()
    launches = (_lib.launch_count() - l0) // (steps + warm) * steps
    ms_step = ms / steps
    px = S * S
    res: dict = {
        "config": {"workload": f"ONE {S}x{S} float32 synthetic DEM, row-sharded over {world} GPU(s) "
                               f"({rows} rows each), all 13 stencil attributes of terrain.py:41-57 "
                               "(slope, aspect, hillshade, 6 curvatures; TPI, TRI, roughness, rugosity), "
                               "surface_fit=Florinsky, 3x3 windows, degrees, hillshade clip",
                   "parallelism": (f"row-shard x{world}, NCCL halo rows (depth 2) exchanged inside every step, "
                                   "overlapped with the interior rows") if world > 1 else "single GPU",
                   "passes_per_step": n_pass,
                   "planes_resident": n_pass == 1,
                   "note": ("the 13 planes of a shard (%.0f GB) exceed free HBM: computed in %d row passes into one "
                            "plane buffer" % (plane_bytes / 1e9, n_pass)) if n_pass > 1 else "planes of the whole "
                           "shard resident"},
        "metric": "Mpixel/s all terrain attributes (13 planes)", "unit": "Mpixel/s", "value": px / ms_step / 1e3,
        "ms_per_step": ms_step, "steps": steps, "warmup": warm, "n_gpus": world, "scaling": "strong",
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": px * 56 / ms_step / 1e6, "peak": peak * world, "unit": "GB/s",
                     "frac": px * 56 / ms_step / 1e6 / (peak * world), "algorithmic_bytes_per_pixel": 56,
                     "kernel": "xbt::florinsky_sliding_kernel<1,1015,1> + xbt::window3_sliding_kernel<1,15>",
                     "note": "56 B/px = 4 read + 13 x 4 written; the request runs as two launches (surface + 3x3 "
                             "windowed), so the DEM is read twice (60 B/px of DRAM traffic); whole step incl. halo "
                             "exchange"},
    }
    if sampler is not None and ctx.rank == 0:
        res["clocks"] = sampler.window(t0, t1)
    del out, buf
    torch.cuda.empty_cache()
    if ctx.rank == 0 and not args.no_cpu:
        res["cpu_baseline"] = run_cpu_terrain(args.c4_cpu_sample, "Florinsky", SURF9, WIN4, repeats=1)
    return res


def extra_c3(ctx: Ctx, args) -> dict:
    import bench_extra

    a = argparse.Namespace(n=args.c3_n, cpu_n=args.c3_cpu_n, steps=3)
    return bench_extra.bench_variogram(a)  # reuses the process group bench.py initialised


def extra_c5(ctx: Ctx, args) -> dict:
    import bench_extra

    a = argparse.Namespace(size=args.c5_size, cpu_size=args.c5_cpu_size, steps=2)
    return bench_extra.bench_nuthkaab(a)


# ---------------------------------------------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------------------------------------------


def reference_arm(args, attrs: list[str], config: dict) -> None:
    """The reference's own CPU implementation of the path on the host cores: oracle/terrain_oracle.c, the C/OpenMP
    restatement of the reference's Numba engine (bit-exact against it on the committed fixtures), all host threads,
    each step one pass over a bounded sample of the workload."""
    if _env_int("RANK", 0) != 0:
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
    sample = (f"each step = one pass over a {args.cpu_sample}x{args.cpu_sample} DEM of the same generator family "
              f"(1/{(args.size // args.cpu_sample) ** 2} of the pixels of the {args.size}^2 config; throughput does not "
              "depend on the raster size beyond the last-level cache)")
    config = dict(config)
    config["workload"] = config["workload"] + f" -- CPU arm: bounded sample, {args.cpu_sample}x{args.cpu_sample} per step"
    config["cpu_sample"] = [args.cpu_sample, args.cpu_sample]
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
    if not args.no_extra:
        extra: dict = {}
        try:
            extra["c1"] = {"cpu_baseline": run_cpu_terrain(4096, "Horn", ["slope"], [], repeats=3),
                           "config": {"workload": "4096x4096 float32 DEM, xdem.terrain.slope (Horn) -- the full config"}}
            extra["c4"] = {"cpu_baseline": run_cpu_terrain(args.c4_cpu_sample, "Florinsky", SURF9, WIN4, repeats=1),
                           "config": {"workload": f"all 13 stencil attributes, {args.c4_cpu_sample}^2 sample"}}
        except Exception as e:  # noqa: BLE001
            extra["error"] = f"{type(e).__name__}: {e}"
        line["extra"] = extra
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# main
# ---------------------------------------------------------------------------------------------------------------


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
    ap.add_argument("--no-extra", action="store_true", help="skip the extra BASELINE configs (c1, c3, c4, c5)")
    ap.add_argument("--extra", default=os.environ.get("XB_BENCH_EXTRA", "c1,c3,c4,c5"))
    ap.add_argument("--c4-size", type=int, default=_env_int("XB_BENCH_C4_SIZE", 65536))
    ap.add_argument("--c4-cpu-sample", type=int, default=2048)
    ap.add_argument("--c3-n", type=int, default=1_000_000)
    ap.add_argument("--c3-cpu-n", type=int, default=40_000)
    ap.add_argument("--c5-size", type=int, default=16384)
    ap.add_argument("--c5-cpu-size", type=int, default=1024)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    attrs = ATTRS if args.fit != "Horn" else ATTRS[:3]

    world = _env_int("WORLD_SIZE", 1)
    config = {
        "workload": f"{args.size}x{args.size} float32 synthetic DEM per GPU, fused {'+'.join(attrs)}, "
                    f"surface_fit={args.fit}, resolution={RESOLUTION}, degrees, hillshade clip",
        "rows_per_gpu": args.size, "cols": args.size, "attributes": attrs, "surface_fit": args.fit,
        "parallelism": f"row-shard x{world} + NCCL halo rows" if world > 1 else "single GPU",
        "l2": "inputs (4.3 GB/GPU) exceed L2 (126 MB); no flush needed",
    }
    if args.impl == "reference":
        reference_arm(args, attrs, config)
        return

    import numpy as np  # noqa: F401

    import bench_data
    import xdem_b200
    from xdem_b200 import _engine, _lib
    from xdem_b200 import distributed as xbd

    ctx = Ctx()
    torch, dist, rank, dev = ctx.torch, ctx.dist, ctx.rank, ctx.dev
    numa = bind_to_gpu_numa(ctx.local_rank) if world > 1 else "single process, not bound"
    depth = 2 if args.fit == "Florinsky" else 1
    rows, cols = args.size, args.size

    # synthetic DEM shard (+ halo rows filled by the exchange); same generator as the full-size parity test
    buf = torch.empty((rows + 2 * depth, cols), dtype=torch.float32, device=dev)
    core = buf[depth:depth + rows]
    bench_data.device_fractal_dem(rows, cols, 42 + rank, dev, out=core)
    out = torch.empty((len(attrs), rows, cols), dtype=torch.float32, device=dev)
    shard = xbd.RowShard(rank, world, depth)
    kwargs = dict(surface_attributes=attrs, surface_fit=args.fit, degrees=True, clip_hillshade=True)
    kernel_events: list = []

    def launch(view, rb: int, re: int, out_row0: int) -> None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _engine.terrain_fused(view, RESOLUTION, row_begin=rb, row_end=re, out=out[:, out_row0:out_row0 + (re - rb)],
                              **kwargs)
        e1.record()
        kernel_events.append((e0, e1))

    def step() -> None:
        shard.run_overlapped(buf, rows, launch)  # NCCL halo rows in flight while the interior rows are computed

    sampler = ClockSampler(ctx.local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    sampler.mark_begin()
    elapsed_ms, t0, t1 = ctx.timed(step, args.steps, args.warmup, kernel_events)
    sampler.mark_end()
    launches = (_lib.launch_count() - launches0) * args.steps // (args.steps + args.warmup)
    kern_ms = sum(a.elapsed_time(b) for a, b in kernel_events) / args.steps  # all kernel launches of a step
    kern_ms = ctx.max_over_ranks(kern_ms)
    clocks = sampler.window(t0, t1) if rank == 0 else {}

    pixels_total = rows * cols * world
    ms_per_step = elapsed_ms / args.steps
    value = pixels_total / (ms_per_step * 1e-3) / 1e6
    peak, peak_src = measured_peak_gbs()
    bytes_per_px = 4 + 4 * len(attrs)
    achieved = rows * cols * bytes_per_px / (kern_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic_per_launch(f"{args.fit.lower()}_{args.size}")

    # what the same traffic mix (one plane read, len(attrs) planes written, no arithmetic) reaches on this box
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

    # ------------------------------ e2e: the public API on host rasters (copies inside the timed region) -----------
    e2e = None
    if not args.no_e2e:
        del out
        torch.cuda.empty_cache()
        # every rank streams ITS shard of the one (world*rows) x cols raster, with the neighbours' real halo rows
        r_begin, r_end, view = shard.prepare(buf, rows)
        torch.cuda.synchronize()
        dem_np = view.cpu().numpy()  # pageable ndarray, as a reference caller holds it
        api_kw = dict(resolution=RESOLUTION, surface_fit=args.fit)

        def api_call():
            if world == 1:
                return xdem_b200.terrain.get_terrain_attribute(dem_np, attrs, **api_kw)
            return _engine.terrain_fused_host(dem_np, RESOLUTION, row_begin=r_begin, row_end=r_end, **kwargs)

        res = api_call()  # warm-up: scratch + first page-locked allocation of the planes
        del res
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            res = api_call()
            del res  # the planes go back to the caching host allocator, as when a caller drops the result
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        h2d = int(dem_np.nbytes) * world
        d2h = rows * cols * 4 * len(attrs) * world
        e2e = {
            "value": pixels_total * args.e2e_steps / dt / 1e6, "unit": "Mpixel/s",
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "steps": args.e2e_steps, "ms_per_step": dt / args.e2e_steps * 1e3,
            "d2h_gbs_per_gpu": d2h / world / (dt / args.e2e_steps) / 1e9,
            "host_binding": numa,
            "call": ("xdem_b200.terrain.get_terrain_attribute(ndarray, [...]) -- pageable input staged block-wise "
                     "through pinned scratch, planes returned in page-locked memory from torch's caching host "
                     "allocator" if world == 1 else
                     "xdem_b200._engine.terrain_fused_host(row_begin, row_end): every rank streams its shard of the "
                     "one raster (with the neighbours' halo rows) -- the multi-GPU form of the API's host path"),
        }
        # the same request through the C ABI with page-locked buffers on both sides
        host_in = torch.empty(tuple(view.shape), dtype=torch.float32, pin_memory=True)
        host_in.copy_(view)
        h_in = host_in.numpy()
        h_out = _engine.host_planes(len(attrs), rows, cols, np.float32)
        _engine.terrain_fused_host(h_in, RESOLUTION, out=h_out, row_begin=r_begin, row_end=r_end, **kwargs)
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            _engine.terrain_fused_host(h_in, RESOLUTION, out=h_out, row_begin=r_begin, row_end=r_end, **kwargs)
        dtp = ctx.max_over_ranks(time.perf_counter() - t0)
        e2e["pinned"] = {"value": pixels_total * args.e2e_steps / dtp / 1e6, "unit": "Mpixel/s",
                         "ms_per_step": dtp / args.e2e_steps * 1e3,
                         "call": "xb_terrain_fused_host_rows (C ABI), page-locked DEM in, page-locked planes out"}
        # what the platform gives: plain cudaMemcpyAsync device -> page-locked host, every rank at once (no kernel, no
        # staging, no API): the aggregate is a property of the host (root complexes / memory), not of this code
        try:
            p_src = view if view.is_contiguous() else view.contiguous()
            p_dst = torch.from_numpy(h_out[0])
            n_el = min(p_src.numel(), p_dst.numel())
            p_src, p_dst = p_src.reshape(-1)[:n_el], p_dst.reshape(-1)[:n_el]
            p_dst.copy_(p_src, non_blocking=True)
            torch.cuda.synchronize()
            ctx.barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                p_dst.copy_(p_src, non_blocking=True)
            torch.cuda.synchronize()
            dtl = ctx.max_over_ranks(time.perf_counter() - t0)
            per_gpu = 3 * n_el * 4 / dtl / 1e9
            e2e["d2h_link_probe"] = {
                "gbs_per_gpu": per_gpu, "gbs_total": per_gpu * world,
                "what": "plain cudaMemcpyAsync device -> page-locked host, all ranks at once, 3 x %.1f GB per rank"
                        % (n_el * 4 / 1e9),
                "e2e_pinned_d2h_gbs_total": d2h / (dtp / args.e2e_steps) / 1e9,
            }
        except Exception as ex:  # noqa: BLE001 -- the probe must never cost the bench line
            e2e["d2h_link_probe"] = {"error": f"{type(ex).__name__}: {ex}"}
        del host_in, h_in, h_out, dem_np
    del buf, core
    torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = run_cpu_terrain(args.cpu_sample, args.fit, attrs, [], repeats=2)

    extra: dict = {}
    if not args.no_extra:
        wanted = [w.strip() for w in args.extra.split(",") if w.strip()]
        for name in wanted:
            try:
                if name == "c1":
                    r = extra_c1(ctx, args, peak) if world == 1 else None
                elif name == "c3":
                    t0c = time.perf_counter()
                    if rank == 0:
                        sampler.mark_begin()
                    r = extra_c3(ctx, args)
                    if rank == 0:
                        sampler.mark_end()
                        if r:
                            w = r.pop("timed_window", None) or [t0c, time.perf_counter()]
                            r["clocks"] = sampler.window(w[0], w[1])
                elif name == "c4":
                    r = extra_c4(ctx, args, peak, sampler if rank == 0 else None)
                elif name == "c5":
                    # no NVML polling here: the queries take driver locks and this step is 370 small launches + 10 syncs
                    r = extra_c5(ctx, args)
                    if r:
                        r.pop("timed_window", None)
                else:
                    r = {"error": f"unknown extra '{name}'"}
            except Exception as e:  # noqa: BLE001
                r = {"error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc(limit=3)}
                torch.cuda.empty_cache()
            if r:
                extra[name] = r
            ctx.barrier()
    if rank == 0:
        sampler.stop()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (exact-difference fp32 stencils; fp64 curvature numerators)",
            "data": "synthetic", "config": config, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "kernel": ("xbt::florinsky_sliding_kernel" if args.fit == "Florinsky"
                                    else "xbt::terrain_fused_kernel"),
                         "kernel_ms": kern_ms, "kernel_ms_source": "CUDA events around the launches inside the timed steps",
                         "algorithmic_bytes_per_pixel": bytes_per_px, "stream_ceiling": stream_ceiling},
            "e2e": e2e, "cpu_baseline": cpu, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
