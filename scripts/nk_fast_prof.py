import sys, torch
sys.path.insert(0, ".")
from xdem_b200 import coreg
S = 16384
yy = torch.arange(S, device="cuda", dtype=torch.float32)[:, None]
xx = torch.arange(S, device="cuda", dtype=torch.float32)[None, :]
def surf(x, y):
    return 300 * torch.sin(x * 0.0037) * torch.cos(y * 0.0027) + 120 * torch.sin((x + 0.6 * y) * 0.0153) + 0.02 * x
ref = surf(xx, yy)
tba = surf(xx + 0.37, yy - 0.61) + 1.5 + 0.02 * torch.randn((S, S), device="cuda")
st = coreg._NKState(ref, tba, None)
st.fast_eligible(72)
for _ in range(3):
    st.iteration_fast(0.37, -0.61, 72)
torch.cuda.synchronize()
