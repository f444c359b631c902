"""1 read + n writes streaming mix: st.global.cs float4 stores vs TMA bulk tensor stores from shared memory."""
import sys, torch
sys.path.insert(0, ".")
from xdem_b200 import _lib
L = _lib.lib()
S = 32768
src = torch.randn((S, S), device="cuda")
dst = torch.empty((4, S, S), device="cuda")
cur = torch.cuda.current_stream().cuda_stream
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for planes in (1, 2, 3, 4):
    ms_a = timeit(lambda: _lib.check(L.xb_probe_stream(src.data_ptr(), dst.data_ptr(), S * S, planes, cur)))
    ms_b = timeit(lambda: _lib.check(L.xb_probe_stream_tma(src.data_ptr(), dst.data_ptr(), S, S, planes, cur)))
    gb = S * S * 4 * (1 + planes) / 1e9
    print(f"planes {planes}: st.global.cs {gb/ms_a*1e3:7.0f} GB/s ({ms_a:.3f} ms)   TMA store {gb/ms_b*1e3:7.0f} GB/s ({ms_b:.3f} ms)")
    # correctness of the TMA variant
    for p in range(planes):
        ref = src.clone(); ref[:, 0::4] += float(p)
        assert torch.equal(dst[p], ref), p
print("copy (torch):", end=" ")
ms = timeit(lambda: dst[0].copy_(src))
print(f"{S*S*8/1e9/ms*1e3:7.0f} GB/s")
