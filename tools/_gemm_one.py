import sys, torch
sys.path.insert(0, '/root/repo')
from dgn_b200 import ops
a = torch.randn(3008, 1984, device='cuda'); b = torch.randn(64, 1984, device='cuda'); out = torch.empty(3008, 64, device='cuda')
for _ in range(3): ops.gemm(a, b, out=out)
torch.cuda.synchronize()
