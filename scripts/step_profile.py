"""Per-C-ABI-call CUDA-event timing of one registration step (GPU diagnostic)."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vcr_net_b200 as V
from vcr_net_b200 import config
from vcr_net_b200._lib import lib
from oracle import synth
from oracle.ref_harness import default_args
prec = sys.argv[1] if len(sys.argv) > 1 else "h3"
partial = len(sys.argv) > 2 and sys.argv[2] == "partial"
config.set_precision(prec)
config.cuda_graph = False          # per-call timing needs the eager launch sequence
iters = 3 if partial else 1         # the benchmark configs: partial --iter 3 (loop invariants hoisted), whole --iter 1
dev = "cuda:0"
lpd = dict(np.load("tests/golden/lpd_pretrained_weights.npz"))
ck = synth.make_checkpoint(1234, emb_weights=lpd)
net = V.VCRNet(default_args(partial=partial, overlap2=synth.OVERLAP2_0575 if partial else 0.75)).to(dev).eval()
net.load_state_dict(synth.checkpoint_to_torch(ck))
B = 24 if partial else 16
p = synth.make_pairs(B, 1024, partial=partial)
s, t = torch.from_numpy(p["src"]).to(dev), torch.from_numpy(p["tgt"]).to(dev)
for _ in range(2): V.vcrnetIter(net, s, t, iter=iters)
torch.cuda.synchronize()
import time; t0 = time.perf_counter()
V.vcrnetIter(net, s, t, iter=iters)
host = time.perf_counter() - t0          # host enqueue of the whole call, WITHOUT the profiling events
torch.cuda.synchronize()
L = lib(); L.profile_begin()
torch.cuda._sleep(20_000_000)             # ~10 ms: the host runs ahead, event pairs bracket pure kernel time
V.vcrnetIter(net, s, t, iter=iters)
prof = L.profile_end()
tot = 0
agg = collections.OrderedDict()
for name, ms, a in prof:
    key = name
    if name in ("vcr_gemm_tc", "vcr_gemm_f32"):
        M, N, K, nbo, nbi = a[18:23]; key = f"{name} M{M} N{N} K{K} nb{nbo*nbi}"
    d = agg.setdefault(key, [0, 0.0, 0.0]); d[0] += 1; d[1] += ms
    if name in ("vcr_gemm_tc", "vcr_gemm_f32"): d[2] += 2.0 * M * N * K * nbo * nbi
    tot += ms
print(f"precision={prec} iters={iters} hoist={config.hoist} calls={len(prof)} sum_ms={tot:.3f} host_enqueue_ms={host*1e3:.2f} (whole call)")
for k, (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:8.3f} ms  x{n:3d}  {k}" + (f"   {fl/ms/1e9:7.1f} TFLOP/s" if fl else ""))
