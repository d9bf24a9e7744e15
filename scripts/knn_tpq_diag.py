"""GPU diagnostic (round 2): thread-per-query tensor-core kNN (csrc/knn_tpq.cuh) vs the warp-per-query kernels and oracle/canon.c.
Bit-exactness on grid (heavy ties) / float / adversarial inputs, then CUDA-event timings with the L2 flushed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vcr_net_b200 import ops
from vcr_net_b200._lib import lib
from oracle import canon
dev = "cuda:0"
L = lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rs = np.random.RandomState(0)


def timed(fn, iters=10):
    fn(); torch.cuda.synchronize()
    ms = 0.0
    for _ in range(iters):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    return ms / iters * 1e3


def casef(B, N, D, grid=False, iters=10, check=False, k=20, dup=False):
    x = rs.randn(B, N, D).astype(np.float32)
    if grid:
        x = np.round(x * 4) / 4
    if dup:
        x[:, N // 2:] = x[:, :N - N // 2]                    # every point twice: exact ties straddle every boundary
    xt = torch.from_numpy(x).to(dev)
    xop = ops.to_operand(xt.view(B * N, D), "h3")
    ref = ops.knn_topk(xt, k, token_major=True).clone()
    us_simt = timed(lambda: ops.knn_topk(xt, k, token_major=True), iters)
    res = {}
    for v in (0, 1):
        L.vcr_set_knn_tc_tpq(v)
        idx, fl = ops.knn_topk_tc(xt, xop, k, want_flagged=True)
        res[v] = (idx.clone(), int(fl.item()), timed(lambda: ops.knn_topk_tc(xt, xop, k), iters))
    L.vcr_set_knn_tc_tpq(2)
    same = torch.equal(ref, res[0][0]) and torch.equal(ref, res[1][0])
    ok = ""
    if check:
        ok = " canon=" + str(np.array_equal(ref.cpu().numpy(), canon.knn(np.ascontiguousarray(x.transpose(0, 2, 1)), k)))
    print(f"D={D:3d} B={B:3d} N={N:5d} k={k} grid={int(grid)} dup={int(dup)} identical={same}{ok}  simt {us_simt:7.1f} | tc warp {res[0][2]:7.1f} "
          f"(redo {res[0][1]}) | tc thread {res[1][2]:7.1f} (redo {res[1][1]}) us  x{us_simt/res[1][2]:4.2f} vs simt", flush=True)
    return same and (not check or "True" in ok)


ok = True
for (B, N, D, grid, dup) in [(2, 64, 64, False, False), (2, 333, 64, False, False), (2, 512, 64, True, False), (2, 700, 32, False, True),
                             (1, 1030, 128, False, False), (2, 256, 16, True, True), (1, 2100, 64, False, False)]:
    ok &= casef(B, N, D, grid, iters=1, check=True, dup=dup)
ok &= casef(2, 300, 64, False, iters=1, check=True, k=5)
ok &= casef(2, 300, 64, False, iters=1, check=True, k=30)
print("ALL EXACT" if ok else "MISMATCH", flush=True)
for (B, N, D) in [(24, 768, 64), (48, 768, 64), (16, 1024, 64), (32, 1024, 64), (8, 4096, 64), (2, 16384, 64), (32, 1024, 128)]:
    casef(B, N, D, iters=5)
