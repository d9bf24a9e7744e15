"""GPU diagnostic: latency of one registration call at small batches, eager (Python / ctypes enqueue of every C-ABI
launch) vs vcr_net_b200.graph.GraphedRegistration (one CUDA-graph launch).  Wall clock around call + synchronize, i.e. what
a caller waiting for the pose sees; inputs resident on the device; median of `reps` calls."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import vcr_net_b200 as V
from vcr_net_b200.graph import GraphedRegistration
from oracle import synth
from oracle.ref_harness import default_args
from bench import load_ckpt

dev = "cuda:0"
ckpt = synth.checkpoint_to_torch(load_ckpt())
reps = 30
print(f"{'workload':8s} {'batch':>5s} {'eager ms':>9s} {'graph ms':>9s} {'speed-up':>8s} {'pairs/s (graph)':>16s}  identical")
for partial, iters in ((False, 1), (True, 3)):
    net = V.VCRNet(default_args(partial=partial, overlap2=synth.OVERLAP2_0575 if partial else 0.75)).to(dev).eval()
    net.load_state_dict(ckpt, strict=True)
    for B in (1, 2, 4, 8, 16):
        p = synth.make_pairs(B, 1024, partial=partial, reserve=synth.RESERVE_0575 if partial else 1.0, first_item=900)
        src, tgt = torch.from_numpy(p["src"]).to(dev), torch.from_numpy(p["tgt"]).to(dev)
        reg = GraphedRegistration(net, batch=B, num_points=src.shape[2], iter=iters)

        def timed(fn):
            ts = []
            for _ in range(reps):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                out = fn()
                torch.cuda.synchronize()
                ts.append((time.perf_counter() - t0) * 1e3)
            return float(np.median(ts)), out

        with torch.no_grad():
            te, oe = timed(lambda: V.vcrnetIter(net, src, tgt, iter=iters))
        tg, og = timed(lambda: reg(src, tgt))
        same = all(torch.equal(a, b) for a, b in zip(oe, og))
        print(f"{'partial' if partial else 'whole':8s} {B:5d} {te:9.3f} {tg:9.3f} {te/tg:8.2f} {B/tg*1e3:16.1f}  {same}", flush=True)
