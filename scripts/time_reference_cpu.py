"""Second CPU baseline: the UNMODIFIED reference (GlacioHack/xdem, loaded where it lies through oracle/refload.py) timed
in the build container, both engines, next to the C/OpenMP port bench.py uses on the GPU box.  The reference cannot
travel to the GPU box (no geo stack there either; /root/reference is absent), so these numbers are recorded here:
    NUMBA_CACHE_DIR=/tmp/numba_cache python scripts/time_reference_cpu.py > profiles/reference_cpu_r02.txt"""
import os, sys, time
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import refload, synth, c_oracle

ref = refload.load_reference()
import numba
print(f"host: {os.cpu_count()} logical CPUs, numba threads {numba.get_num_threads()}; reference loaded from {refload.REFERENCE_ROOT}")
C2 = ["slope", "aspect", "hillshade", "curvature"]
cases = [("c1 Horn slope", 4096, ["slope"], "Horn"), ("c2 Florinsky 4 attrs", 4096, C2, "Florinsky"),
         ("c2 ZevenbergThorne 4 attrs", 4096, C2, "ZevenbergThorne")]
for name, n, attrs, fit in cases:
    dem = synth.fractal_dem((n, n), seed=42)
    for engine in ("numba", "scipy"):
        if engine == "scipy" and n > 2048:
            d = dem[:2048, :2048]
        else:
            d = dem
        best = None
        for rep in range(3 if engine == "numba" else 1):
            t0 = time.perf_counter()
            out = ref.terrain.get_terrain_attribute(d, attrs, resolution=5.0, surface_fit=fit, engine=engine)
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best  # first numba call includes JIT: best of 3
        print(f"{name:28s} reference engine={engine:5s} {d.shape[0]}^2: {best:8.3f} s  {d.size / best / 1e6:8.2f} Mpix/s")
    best = None
    for rep in range(3):
        t0 = time.perf_counter()
        c_oracle.surface_attributes(dem, 5.0, attrs, surface_fit=fit)
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    print(f"{name:28s} C/OpenMP port (bench.py's CPU arm, {c_oracle.num_threads()} threads) {n}^2: {best:8.3f} s  "
          f"{dem.size / best / 1e6:8.2f} Mpix/s")
