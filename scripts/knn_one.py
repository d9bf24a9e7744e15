"""one launch of each thread-per-query kNN kernel (for ncu)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vcr_net_b200 import ops
torch.manual_seed(0)
x3 = torch.rand(24, 3, 768, device="cuda") - 0.5
xf = torch.randn(24, 768, 64, device="cuda")
xop = ops.to_operand(xf.view(-1, 64), "h3")
for _ in range(2):
    ops.knn_topk(xf, 20, token_major=True)
    ops.knn_topk_tc(xf, xop, 20)
torch.cuda.synchronize()
