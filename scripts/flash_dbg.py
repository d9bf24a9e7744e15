import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import math, torch
from vcr_net_b200 import ops
dev = "cuda:0"
torch.manual_seed(0)
B, H, Nq, Nk, dk, mode = int(sys.argv[1]), 4, int(sys.argv[2]), int(sys.argv[3]), 128, "h3"
q = torch.randn(B * Nq, H * dk, device=dev); k = torch.randn(B * Nk, H * dk, device=dev); vt = torch.randn(B * H * dk, Nk, device=dev)
Q, K, VT = ops.to_operand(q, mode), ops.to_operand(k, mode), ops.to_operand(vt, mode)
res = {}
for org in (2, 3):
    ops.set_flash_warps(org)
    out = ops.Operand.empty(B * Nq, H * dk, mode, dev)
    out.buf.view(torch.int16).fill_(0x7e00)        # fp16 NaN poison
    ops.flash_attn_tc(Q, K, VT, out, B, H, Nq, Nk, dk, 1.0 / math.sqrt(dk))
    torch.cuda.synchronize()
    res[org] = out.to_float().clone()
a, b = res[2], res[3]
bad = ~(torch.isfinite(b)) | ((a - b).abs() > 1e-6)
print("bad elements", int(bad.sum()), "of", bad.numel())
if bad.any():
    rows = bad.any(dim=1).nonzero().flatten()
    cols = bad.any(dim=0).nonzero().flatten()
    r = rows.cpu().numpy(); c = cols.cpu().numpy()
    print("bad rows", len(r), "first", r[:10], "last", r[-5:], " items (row//128):", sorted(set((r // 128).tolist()))[:40])
    print("bad cols", len(c), c[:8], c[-4:])
    print("nan count", int((~torch.isfinite(b)).sum()))
