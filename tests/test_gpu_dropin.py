"""The reference's REAL entry on the B200 path (VERDICT r1 "missing #2", reference util/initPara.py:237-260, main.py):

  * ``nn.DataParallel(VCRNet(args))`` -- the wrapper the reference always applies (util/initPara.py:260) -- on 1 device
    and, when the box has them, 2 devices: same bits as the bare module, twice in a row (exercises ``replicate()``'s
    per-forward shallow copies against the packed-weight caches, and the per-device statics of the C library from
    ``parallel_apply``'s threads);
  * LPD pre-training (--model=lpd) through DataParallel replicas: gradients reach the original parameters;
  * the unmodified ``main.py --eval`` through ``vcr_net_b200.dropin`` (oracle/run_main.py supplies the harness-side
    ``tensorboardX`` / ``h5py`` stubs, the synthetic HDF5-shaped dataset and the synthetic ``.t7``): the metrics it PRINTS
    are compared with the live reference's own (tests/golden/headline_cfg*.npz, produced by the reference's test_one_epoch).
    Needs the staged reference (oracle/_ref, `python -m oracle.build_ref`); skipped when it is absent.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import vcr_net_b200 as V
from oracle import build_ref, synth
from oracle.ref_harness import default_args

DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _net(ckpt, partial):
    net = V.VCRNet(default_args(partial=partial, overlap2=synth.OVERLAP2_0575 if partial else 0.75)).to(DEV).eval()
    net.load_state_dict(synth.checkpoint_to_torch(ckpt), strict=True)
    return net


@pytest.mark.parametrize("partial", [False, True])
def test_dataparallel_one_device_same_bits(ckpt, partial):
    net = _net(ckpt, partial)
    p = synth.make_pairs(4, 512, partial=partial, first_item=300)
    src, tgt = cu(p["src"]), cu(p["tgt"])
    bare = net(src, tgt)
    dp = nn.DataParallel(net, device_ids=[0])
    for _ in range(2):
        out = dp(src, tgt)
        for a, b in zip(bare, out):
            assert torch.equal(a, b)
    # the reference's loop calls vcrnetIter(net = the DataParallel wrapper, ...) (model/vcrnet_model.py:561-563)
    it_bare = V.vcrnetIter(net, src, tgt, iter=2)
    it_dp = V.vcrnetIter(dp, src, tgt, iter=2)
    for a, b in zip(it_bare, it_dp):
        assert torch.equal(a, b)


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("partial", [False, True])
def test_dataparallel_two_devices_same_bits(ckpt, partial):
    """replicate() + scatter + parallel_apply (one Python thread per device) + gather: pairs are independent, so each
    replica's half must carry the bare module's bits, call after call."""
    net = _net(ckpt, partial)
    p = synth.make_pairs(6, 512, partial=partial, first_item=310)
    src, tgt = cu(p["src"]), cu(p["tgt"])
    bare = V.vcrnetIter(net, src, tgt, iter=2)
    dp = nn.DataParallel(net, device_ids=[0, 1])
    for _ in range(3):
        out = V.vcrnetIter(dp, src, tgt, iter=2)
        for a, b in zip(bare, out):
            assert a.device == b.device and torch.equal(a, b)
    # weights updated in place between calls: replicas must see the new values (no stale packed weights)
    with torch.no_grad():
        net.pointer.model.decoder.norm.b_2.add_(0.25)
    changed_bare = net(src, tgt)
    changed_dp = dp(src, tgt)
    for a, b in zip(changed_bare, changed_dp):
        assert torch.equal(a, b)


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_lpd_pretraining_through_dataparallel_replicas():
    """--model=lpd under nn.DataParallel (util/initPara.py:239, 260): replicas have no parameters() of their own, the
    training path must still be taken and the gradients must reach the wrapped module's parameters."""
    lpd = load_golden("lpd_pretrained_weights")
    net = V.LPD(default_args(model="lpd", num_points=256)).to(DEV).train()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in lpd.items()}, strict=True)
    p = synth.make_pairs(4, 256, aligned=True, first_item=60)
    src, tgt = cu(p["src"]), cu(p["tgt"])
    net.zero_grad()
    loss_bare = net(src, tgt)[2]
    loss_bare.backward()
    g_bare = {k: v.grad.clone() for k, v in net.named_parameters()}
    net.zero_grad()
    dp = nn.DataParallel(net, device_ids=[0, 1])
    out = dp(src, tgt)
    loss = out[2].mean()                      # trainLPD: loss.sum() over the gathered per-replica losses
    loss.backward()
    for k, v in net.named_parameters():
        assert v.grad is not None and torch.isfinite(v.grad).all(), k
    # each replica's loss is a mean over its own half: the average of the two is close to (not equal to) the full-batch loss
    assert abs(float(loss) - float(loss_bare)) < 0.2 * abs(float(loss_bare)) + 1e-3


def _run_main(*argv, dropin=True, timeout=900):
    cmd = [sys.executable, "-m", "oracle.run_main"] + (["--dropin"] if dropin else []) + ["--"] + list(argv)
    env = dict(os.environ, VCR_CUDA_GRAPH="1")                    # the product default (tests/conftest.py turns it off)
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


needs_ref = pytest.mark.skipif(not build_ref.staged(), reason="reference not staged under oracle/_ref (python -m oracle.build_ref)")


@needs_ref
@pytest.mark.parametrize("cfg", ["cfg1", "cfg2"])
def test_unmodified_main_py_eval_runs_on_the_b200_path(cfg, capsys):
    """python main.py --eval ... with the hot-path symbols rebound (vcr_net_b200.dropin): runs para() -> initNet ->
    load_state_dict(.t7) -> nn.DataParallel -> testVCRNet -> test_one_epoch unchanged and prints the reference's report line.
    Its numbers against the live reference's own metrics for the same 48 items."""
    g = load_golden("headline_" + cfg)
    argv = ["--eval", "--iter", str(int(g["iters"])), "--test_batch_size", str(int(g["batch"])),
            "--num_points", str(int(g["num_points"]))]
    if cfg == "cfg2":
        argv += ["--partial", "--overlap", "0.575"]
    res = _run_main(*argv, dropin=True)
    assert res["finished"] and res["metrics"] is not None and res["rebound"]
    errs = {k: abs(v - float(g["m." + k])) / max(abs(float(g["m." + k])), 1e-12) for k, v in res["metrics"].items()
            if "m." + k in g and abs(float(g["m." + k])) > 0}
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    rec = {"printed_by_main_py": res["metrics"], "reference_metrics": {k: float(g["m." + k]) for k in res["metrics"] if "m." + k in g},
           "rel_err": errs, "wall_s": res["wall_s"], "argv": res["argv"]}
    with open(os.path.join(out, f"dropin_main_py_{cfg}.json"), "w") as f:
        json.dump(rec, f, indent=1, sort_keys=True)
    with capsys.disabled():
        print("\n[dropin main.py] " + cfg + " " + json.dumps(rec, sort_keys=True))
    bound = 5e-4 if cfg == "cfg1" else 2e-2
    # the report prints %f (6 decimals): allow that quantisation on top of the relative bound
    bad = {k: e for k, e in errs.items() if e > bound + 1e-6 / max(abs(float(g["m." + k])), 1e-12)}
    assert not bad, bad
