import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from variational_mmt_b200 import ops
M, N, K = [int(x) for x in sys.argv[1:4]]
a = torch.randn(M, K, device="cuda"); b = torch.randn(N, K, device="cuda"); c = torch.empty(M, N, device="cuda")
for _ in range(3):
    ops.gemm(a, b, c, M, N, K)
torch.cuda.synchronize()
