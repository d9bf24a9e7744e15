"""GPU diagnostic: key statistic (attention column sums): error against an fp64 product and timing at the headline shape.
(Round 2 ran it with every n-th exponential on the FMA pipe, n = 8, 4, 3, 2: profiles/r02_colsum_diag.txt.)  Run under `timeout`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import math
import torch
from vcr_net_b200 import ops
dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
torch.manual_seed(0)


def case(B, H, Nq, Nk, spread, iters=10):
    dk = 128
    q = torch.randn(B * Nq, H * dk, device=dev) * spread; k = torch.randn(B * Nk, H * dk, device=dev)
    Q, K = ops.to_operand(q, "h3"), ops.to_operand(k, "h3")
    q64 = q.double().view(B, Nq, H, dk).permute(0, 2, 1, 3); k64 = k.double().view(B, Nk, H, dk).permute(0, 2, 1, 3)
    ref = torch.zeros(B, Nk, dtype=torch.float64, device=dev)
    for b in range(min(B, 4)):
        s = torch.matmul(q64[b], k64[b].transpose(-1, -2)) / math.sqrt(dk)
        ref[b] = torch.softmax(s, -1).sum(dim=(0, 1))
    line = f"B={B:3d} H={H} Nq={Nq:5d} Nk={Nk:5d} spread={spread:4.1f} "
    for n in (0,):
        out = ops.attn_colsum_tc(Q, K, B, H, Nq, Nk, dk, 1.0 / math.sqrt(dk))
        torch.cuda.synchronize()
        nb = min(B, 4)
        err = float(((out[:nb].double() - ref[:nb]).abs() / ref[:nb].abs().clamp_min(1e-30)).max())
        ms = 0.0
        for _ in range(iters):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.attn_colsum_tc(Q, K, B, H, Nq, Nk, dk, 1.0 / math.sqrt(dk)); e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        line += f"| {ms / iters * 1e3:7.1f} us, max rel err vs fp64 {err:.1e} "
    print(line, flush=True)


case(2, 4, 200, 332, 1.0, iters=2)
case(24, 4, 768, 768, 1.0)
case(48, 4, 768, 768, 1.0)
case(48, 4, 768, 768, 3.0)
case(8, 4, 4096, 4096, 1.0, iters=3)
