"""GPU diagnostic: the e2e bracket of bench.py split into H2D | vcrnetIter | D2H, next to the device-resident step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import vcr_net_b200 as V
dev = torch.device("cuda:0")
sys.argv = sys.argv[:1]
a = bench.parse()
cfg = bench.workload_cfg(a, "partial", 1)
net = bench.build_net(cfg, dev)
from vcr_net_b200 import synthetic
B = cfg["batch"]
src_ = synthetic.PairSource(2 * B, dev, num_points=cfg["num_points"], partial=True, reserve=cfg["reserve"], base_points=2048, aligned=False)
b0 = src_.batch(0, B)
s, t = b0["src"], b0["tgt"]
hs, ht = s.cpu().pin_memory(), t.cpu().pin_memory()
sb, tb = torch.empty_like(s), torch.empty_like(t)
hp = [torch.empty(sh).pin_memory() for sh in ((B, 3, 3), (B, 3), (B, 3, 3), (B, 3))]
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
with torch.no_grad():
    for _ in range(4):
        V.vcrnetIter(net, s, t, iter=cfg["iters"]); V.vcrnetIter(net, sb, tb, iter=cfg["iters"])
    torch.cuda.synchronize()
    n = 20
    acc = [0.0] * 5
    for i in range(n):
        flush.fill_(float(i))
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        sb.copy_(hs, non_blocking=True); tb.copy_(ht, non_blocking=True)
        e[1].record()
        out = V.vcrnetIter(net, sb, tb, iter=cfg["iters"])
        e[2].record()
        for h, o in zip(hp, out[2:6]): h.copy_(o, non_blocking=True)
        e[3].record()
        flush.fill_(float(i))
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(); V.vcrnetIter(net, s, t, iter=cfg["iters"]); f1.record()
        torch.cuda.synchronize()
        acc[0] += e[0].elapsed_time(e[1]); acc[1] += e[1].elapsed_time(e[2]); acc[2] += e[2].elapsed_time(e[3]); acc[3] += e[0].elapsed_time(e[3])
        acc[4] += f0.elapsed_time(f1)
print("per step (ms): H2D %.3f | vcrnetIter %.3f | D2H %.3f | e2e bracket %.3f || device-resident step %.3f" % tuple(x / n for x in acc))
