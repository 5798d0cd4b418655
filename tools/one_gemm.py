import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from framedipt_b200 import runtime
ctx = runtime.Context()
M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
a = torch.randn(1, M, K, device="cuda"); b = torch.randn(1, N, K, device="cuda")
for _ in range(4):
    ctx.matmul(a, b, True)
torch.cuda.synchronize()
