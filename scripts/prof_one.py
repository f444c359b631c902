"""Launch the fused terrain kernel a few times on one workload (for ncu captures)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from xdem_b200 import _engine
size = int(sys.argv[1]); fit = sys.argv[2]; attrs = sys.argv[3].split(','); n = int(sys.argv[4]) if len(sys.argv) > 4 else 3
win = sys.argv[5].split(',') if len(sys.argv) > 5 and sys.argv[5] else []
dev = torch.device('cuda')
g = torch.Generator(device=dev).manual_seed(42)
z = torch.empty((size, size), device=dev)
carry = torch.zeros((1, size), device=dev)
for r0 in range(0, size, 4096):
    nn = torch.randn((min(4096, size - r0), size), generator=g, device=dev)
    blk = torch.cumsum(nn, 0) + carry; carry = blk[-1:].clone()
    z[r0:r0 + blk.shape[0]] = 1000 + 0.05 * torch.cumsum(blk, 1)
out = None
for _ in range(n):
    out = _engine.terrain_fused(z, 5.0, surface_attributes=[a for a in attrs if a], windowed_indexes=win, surface_fit=fit, degrees=True, clip_hillshade=True, out=out)
torch.cuda.synchronize()
