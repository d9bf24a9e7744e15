"""GPU diagnostic: 3-d kNN, on-the-fly distance kernel (knn3_kernel) vs the generic tile kernel: identical indices, timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vcr_net_b200 import ops
from oracle import canon
dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rs = np.random.RandomState(0)


def case(B, N, tm, grid=False, iters=10, check=False):
    x = rs.rand(B, 3, N).astype(np.float32) - 0.5
    if grid:
        x = np.round(x * 32) / 32                      # heavy ties / duplicates
    xt = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 2, 1) if tm else x)).to(dev)
    res = {}
    for on in (False, True):
        ops.set_knn3_direct(on)
        idx = ops.knn_topk(xt, 20, token_major=tm)
        torch.cuda.synchronize()
        ms = 0.0
        for _ in range(iters):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.knn_topk(xt, 20, token_major=tm); e1.record(); torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        res[on] = (idx.clone(), ms / iters)
    ops.set_knn3_direct(True)
    same = torch.equal(res[False][0], res[True][0])
    ok = ""
    if check:
        ok = " canon=" + str(np.array_equal(res[True][0].cpu().numpy(), canon.knn(x, 20)))
    print(f"B={B:3d} N={N:5d} tm={int(tm)} grid={int(grid)}  identical={same}{ok}  tile {res[False][1]*1e3:7.1f} us | direct {res[True][1]*1e3:7.1f} us  "
          f"x{res[False][1]/res[True][1]:4.2f}", flush=True)
    return same


ok = True
for (B, N, tm, grid) in [(2, 21, False, False), (3, 33, True, False), (2, 777, False, True), (2, 1024, True, True), (1, 1500, False, True),
                         (2, 100, False, True)]:
    ok &= case(B, N, tm, grid, iters=1, check=True)
print("ALL IDENTICAL" if ok else "MISMATCH", flush=True)
case(48, 768, False)
case(48, 768, True)
case(32, 1024, False)
case(32, 4096, False, iters=3)
case(256, 1024, False, iters=3)
