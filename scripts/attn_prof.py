"""GPU profiling driver: one flash attention call and one key-statistic call at the headline shapes (48 x 4 x 768 x 768, parity
mode), for `ncu --set full --import-source on -k regex:...`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import math
import torch
from vcr_net_b200 import ops
dev = "cuda:0"
torch.manual_seed(0)
B, H, N, dk, mode = 48, 4, 768, 128, "h3"
q = torch.randn(B * N, H * dk, device=dev); k = torch.randn(B * N, H * dk, device=dev)
vt = torch.randn(B * H * dk, N, device=dev)
Q, K, VT = ops.to_operand(q, mode), ops.to_operand(k, mode), ops.to_operand(vt, mode)
out = ops.Operand.empty(B * N, H * dk, mode, dev)
for _ in range(3):
    ops.flash_attn_tc(Q, K, VT, out, B, H, N, N, dk, 1.0 / math.sqrt(dk))
    cs = ops.attn_colsum_tc(Q, K, B, H, N, N, dk, 1.0 / math.sqrt(dk))
torch.cuda.synchronize()
