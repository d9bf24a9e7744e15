"""CPU checks of the boundary: the C-ABI library loads, exports every symbol include/vcr_b200.h
declares, the product never touches oracle/ and has no CPU fallback, state_dict layout matches."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden


@pytest.fixture(scope="module", autouse=True)
def built():
    import __graft_entry__ as g
    g.build()


def test_header_symbols_exported():
    from vcr_net_b200._lib import LIB_PATH, parse_header
    protos = parse_header()
    assert len(protos) >= 28
    cdll = ctypes.CDLL(LIB_PATH)
    for name in protos:
        assert hasattr(cdll, name), name


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "vcr_net_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f
                assert "/root/reference" not in txt, f


def test_no_cpu_fallback():
    import vcr_net_b200 as V
    from oracle.ref_harness import default_args
    net = V.VCRNet(default_args())
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 64), torch.zeros(1, 3, 64))
    with pytest.raises(RuntimeError):
        V.knn(torch.zeros(1, 3, 64), 20)
    with pytest.raises(RuntimeError):
        V.farthest_point_sample(torch.zeros(1, 3, 64), 4)
    if not torch.cuda.is_available():
        from vcr_net_b200.graph import GraphedRegistration
        with pytest.raises(RuntimeError):
            GraphedRegistration(net, batch=1, num_points=64)


def test_gemm_pair_policy_switch_roundtrip():
    """The CTA-pair policy is a process-wide setting of the library (no GPU needed to set / read it back)."""
    from vcr_net_b200 import ops, config
    assert config.GEMM_PAIR_CODES[str(config.gemm_pair)] in (0, 1, 2)
    first = ops.set_gemm_pair(False)
    try:
        assert ops.set_gemm_pair(True) == 0
        assert ops.set_gemm_pair("auto") == 1
        assert ops.set_gemm_pair(3) == 2
        assert ops.set_gemm_pair(0) == 3
    finally:
        ops.set_gemm_pair(first)


def test_state_dict_layout_and_t7_roundtrip(tmp_path, ckpt):
    import vcr_net_b200 as V
    from oracle import synth
    from oracle.ref_harness import default_args
    g = load_golden("state_dict_layout")
    net = V.VCRNet(default_args())
    sd = net.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    assert [str(tuple(v.shape)) for v in sd.values()] == [str(s) for s in g["shapes"]]
    # legacy (non-zip) pickle, the reference's .t7 format; loads with strict=False like util/initPara.py:254
    path = str(tmp_path / "model.t7")
    torch.save(synth.checkpoint_to_torch(ckpt), path, _use_new_zipfile_serialization=False)
    res = net.load_state_dict(torch.load(path, map_location="cpu"), strict=False)
    assert not res.missing_keys and not res.unexpected_keys
    lpd = {k: torch.from_numpy(v) for k, v in load_golden("lpd_pretrained_weights").items()}
    res = net.load_state_dict(lpd, strict=False)          # lpd-pretrained seeds emb_nn only
    assert not res.unexpected_keys and len(res.missing_keys) == 59 - 12
    lpdm = V.LPD(default_args())
    lpdm.load_state_dict(lpd, strict=True)
    assert lpdm.emb_nn.negative_slope == 0.2 and net.emb_nn.negative_slope == 0.0


def test_lpdnet_transform_net_state_dict_layout():
    """--t3d / --tfea: key order of the live reference's LPDNet(t3d, tfea).state_dict() (tests/golden/tnet.npz)."""
    from vcr_net_b200.model.lpdnet_model import LPDNet
    from oracle import synth
    from oracle.ref_harness import default_args
    g = load_golden("tnet")
    for name, t3d, tfea in (("both", True, True), ("t3d", True, False)):
        net = LPDNet(default_args(t3d=t3d, tfea=tfea, emb_dims=128))
        assert list(net.state_dict().keys()) == [str(k) for k in g[f"{name}.keys"]]
        res = net.load_state_dict(synth.checkpoint_to_torch(synth.make_tnet_lpdnet_weights(21, t3d, tfea, 128)), strict=False)
        assert not res.unexpected_keys and all(k.endswith("num_batches_tracked") for k in res.missing_keys)


def test_host_svd_matches_numpy():
    from vcr_net_b200._lib import lib
    L = lib()
    rs = np.random.RandomState(0)
    for trial in range(200):
        H = rs.randn(3, 3)
        if trial % 5 == 0:
            H[:, 2] = H[:, 0] * 0.5 + H[:, 1]            # rank 2
        if trial % 17 == 0:
            H = np.outer(rs.randn(3), rs.randn(3))       # rank 1
        U, S, V = np.zeros((3, 3)), np.zeros(3), np.zeros((3, 3))
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        L.vcr_host_svd3(p(H), p(U), p(S), p(V))
        assert np.allclose(S, np.linalg.svd(H)[1], atol=1e-12)
        assert np.allclose(U @ np.diag(S) @ V.T, H, atol=1e-12)
        assert np.allclose(U.T @ U, np.eye(3), atol=1e-10) and np.allclose(V.T @ V, np.eye(3), atol=1e-10)


def test_dropin_rebinds_reference_symbols():
    """The reference's own modules end up exporting OUR hot-path classes (main.py unchanged)."""
    from oracle import ref_harness
    if not ref_harness.available():
        pytest.skip("reference tree not mounted on this box")
    ref_harness._install_shims()
    import sys
    from vcr_net_b200 import dropin
    import vcr_net_b200 as V
    done = dropin.install(ref_harness.REF_ROOT)
    import model.vcrnet_model as rvm
    import model.lpdnet_model as rlm
    import util.util as ru
    assert rvm.VCRNet is V.VCRNet and rvm.vcrnetIter is V.vcrnetIter and rvm.LPDNet is V.LPDNet
    assert rlm.LPD is V.LPD and ru.knn is V.knn and rlm.knn is V.knn
    assert callable(rvm.testVCRNet) and rvm.testVCRNet.__module__ == "model.vcrnet_model"   # loops stay the reference's
    assert set(done) == set(dropin.HOT_PATH)
    for name in list(sys.modules):                    # leave no reference modules behind for other tests
        if name in ("model", "util") or name.startswith(("model.", "util.")):
            del sys.modules[name]


def test_every_entry_point_rejects_null_arguments():
    """Error behaviour of the boundary: every stream-taking entry point validates its arguments before it touches CUDA
    and answers an all-null / all-zero call with a negative code (so this runs without a GPU)."""
    import ctypes
    from vcr_net_b200._lib import lib
    L = lib()
    checked = 0
    for name, (_, args) in L.protos.items():
        if not args or args[-1][1] != "stream":
            continue
        vals = [None if t is ctypes.c_void_p else t(0) for t, _ in args]
        rc = getattr(L.cdll, name)(*vals)
        assert isinstance(rc, int) and rc < 0, (name, rc)
        checked += 1
    assert checked >= 50


def _cpu_replicate(module):
    """What torch.nn.parallel.replicate() does to a module tree, on the CPU: shallow __dict__ copies flagged _is_replica, with
    empty _parameters and the (broadcast) parameter copies set as plain tensor attributes."""
    import torch
    mods = list(module.modules())
    idx = {m: i for i, m in enumerate(mods)}
    reps = [m._replicate_for_data_parallel() for m in mods]
    for m, r in zip(mods, reps):
        for key, child in m._modules.items():
            r._modules[key] = None if child is None else reps[idx[child]]
        for key, p in m._parameters.items():
            if p is not None:
                c = p.detach().clone()
                c.requires_grad_(p.requires_grad)
                setattr(r, key, c * 1.0 if p.requires_grad else c)       # non-leaf copy, like Broadcast's outputs
    return reps[0]


def test_dataparallel_replica_code_paths():
    """inside nn.DataParallel replicas parameters() is empty and __dict__ is a shallow copy of the original's.
    The LPD training-mode test and the 12 trainable tensors must come from the conv attributes, and the packed-weight cache
    of a replica must be its own dict, not the original's (shared by every replica thread)."""
    import torch
    import vcr_net_b200 as V
    from vcr_net_b200 import functional as Fn
    from oracle.ref_harness import default_args
    net = V.LPDNet(default_args(), negative_slope=0.2)
    Fn.packed(net, "probe", [net.conv1_lpd.weight], lambda: "original")
    rep = _cpu_replicate(net)
    assert list(rep.parameters()) == [] and rep._is_replica
    ts = Fn.lpdnet_param_tensors(rep)
    assert len(ts) == 12 and all(isinstance(t, torch.Tensor) and t.requires_grad for t in ts)
    assert [tuple(t.shape) for t in ts] == [tuple(dict(net.named_parameters())[n].shape) for n in Fn.LPDNET_PARAM_ORDER]
    assert rep.__dict__["_vcr_packed"] is net.__dict__["_vcr_packed"]          # the shared dict the replica must not use
    calls = []
    v1 = Fn.packed(rep, "probe", [rep.conv1_lpd.weight], lambda: calls.append(1) or "replica")
    v2 = Fn.packed(rep, "probe", [rep.conv1_lpd.weight], lambda: calls.append(1) or "replica")
    assert v1 == v2 == "replica" and len(calls) == 1                           # cached for the replica's lifetime ...
    assert net.__dict__["_vcr_packed"]["probe"][1] == "original"               # ... without touching the original's cache
    assert "_vcr_packed_replica" not in net.__dict__
    rep2 = _cpu_replicate(net)                                                  # the next forward's replica starts clean
    assert "_vcr_packed_replica" not in rep2.__dict__


def test_product_synthetic_helpers_match_the_test_infrastructure():
    """bench.py builds its arguments, checkpoint and inputs from vcr_net_b200.synthetic (product side); they are the same
    bytes oracle/synth (test infrastructure, the thing the golden vectors were generated with) produces."""
    import numpy as np
    from vcr_net_b200 import synthetic
    from oracle import synth
    from oracle.ref_harness import default_args
    lpd = dict(np.load(os.path.join(ROOT, "tests", "golden", "lpd_pretrained_weights.npz")))
    a, b = synthetic.state_dict(1234, emb_weights=lpd), synth.make_checkpoint(1234, emb_weights=lpd)
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert np.array_equal(a[k].numpy(), b[k]), k
    assert vars(synthetic.default_args(partial=True, overlap2=0.7)) == vars(default_args(partial=True, overlap2=0.7))
    r, o2 = synthetic.reserve_overlap2(0.575)
    assert abs(r - synth.RESERVE_0575) < 1e-8 and abs(o2 - synth.OVERLAP2_0575) < 1e-8
