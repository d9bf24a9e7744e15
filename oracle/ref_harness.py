"""Import the LIVE reference with harness-only shims.

TEST INFRASTRUCTURE ONLY.  The reference is looked up at ``$VCR_REFERENCE_ROOT``, then ``/root/reference`` (build
container only), then ``oracle/_ref/VCR-Net`` (the unmodified files staged by ``python -m oracle.build_ref``; git-ignored,
travels to the GPU box with the snapshot).  Nothing under ``vcr_net_b200/`` imports this module.  It exists to
(1) generate tests/golden/ with oracle/make_golden.py, (2) let ``tests/test_ref_live.py`` (skipped when no reference can
be found) compare the GPU path with the live reference at BASELINE's sizes, (3) time the reference's own CPU path for
``bench.py --impl reference`` and (4) run the reference's unmodified ``main.py`` (oracle/run_main.py).

Shims; no reference source is modified:
  * stub ``pynvml`` when NVML is missing (util/util.py:13 calls nvmlInit at import)
  * ``Rotation.from_dcm`` -> ``from_matrix`` (util/util.py:102; removed from scipy)
  * ``torch.cuda.FloatTensor`` on CPU (model/lpdnet_model.py:186)
  * stub ``tensorboardX`` / ``h5py`` (oracle/stubs/) when the real packages are absent (util/initPara.py:15, util/data.py:9)
  * ``cpu_mode()``: ``Tensor.cuda`` / ``Module.cuda`` become no-ops and ``torch.cuda.is_available`` says False, so that the
    reference's GPU-only loops (unconditional ``.cuda()``, model/vcrnet_model.py:550-555, util/initPara.py:234-241) run on
    the host cores -- this is how the reference's own ``test_one_epoch`` produces the golden metrics.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types
from argparse import Namespace

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(HERE, "_ref", "VCR-Net")


def _find_root():
    for cand in (os.environ.get("VCR_REFERENCE_ROOT"), "/root/reference", STAGED_ROOT):
        if cand and os.path.isfile(os.path.join(cand, "model", "vcrnet_model.py")):
            return cand
    return os.environ.get("VCR_REFERENCE_ROOT", "/root/reference")


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "model", "vcrnet_model.py"))


def _install_shims():
    import importlib.util

    import torch
    try:
        import pynvml
        pynvml.nvmlInit()
    except Exception:
        stub = types.ModuleType("pynvml")
        stub.nvmlInit = lambda: None
        stub.nvmlDeviceGetHandleByIndex = lambda i: None
        stub.nvmlDeviceGetMemoryInfo = lambda h: types.SimpleNamespace(used=0)
        sys.modules["pynvml"] = stub
    from scipy.spatial.transform import Rotation
    if not hasattr(Rotation, "from_dcm"):
        Rotation.from_dcm = Rotation.from_matrix
    if not torch.cuda.is_available():
        torch.cuda.FloatTensor = lambda data: torch.tensor(data, dtype=torch.float32)
    stubs = os.path.join(HERE, "stubs")
    if any(importlib.util.find_spec(m) is None for m in ("tensorboardX", "h5py")) and stubs not in sys.path:
        sys.path.append(stubs)                       # appended: a real installation always wins


@contextlib.contextmanager
def cpu_mode():
    """Run GPU-only reference code on the host: ``.cuda()`` is the identity, CUDA reports unavailable."""
    import torch
    saved = (torch.Tensor.cuda, torch.nn.Module.cuda, torch.cuda.is_available, torch.cuda.device_count,
             torch.cuda.manual_seed_all, getattr(torch.cuda, "FloatTensor", None))
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.is_available = lambda: False
    torch.cuda.device_count = lambda: 0
    torch.cuda.manual_seed_all = lambda s: None
    torch.cuda.FloatTensor = lambda data: torch.tensor(data, dtype=torch.float32)
    try:
        yield
    finally:
        (torch.Tensor.cuda, torch.nn.Module.cuda, torch.cuda.is_available, torch.cuda.device_count,
         torch.cuda.manual_seed_all) = saved[:5]
        if saved[5] is not None:
            torch.cuda.FloatTensor = saved[5]


def _purge_foreign_packages():
    # the reference uses top-level package names ``model`` and ``util``
    for name in list(sys.modules):
        if name in ("model", "util") or name.startswith(("model.", "util.")):
            mod = sys.modules[name]
            if not (getattr(mod, "__file__", None) or "").startswith(REF_ROOT):
                del sys.modules[name]


def import_reference():
    """Returns a namespace with the reference's hot-path modules."""
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT} (stage it with `python -m oracle.build_ref`)")
    _install_shims()
    sys.dont_write_bytecode = True
    _purge_foreign_packages()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import model.lpdnet_model as lpdnet_model
    import model.transformer as transformer
    import model.vcrnet_model as vcrnet_model
    import util.util as util
    import model.icp_model as icp_model
    return types.SimpleNamespace(lpdnet_model=lpdnet_model, transformer=transformer,
                                 vcrnet_model=vcrnet_model, util=util, icp_model=icp_model)


def default_args(partial=False, overlap2=0.75, **kw):
    """The fields the module constructors read, util/initPara.py defaults."""
    a = dict(emb_dims=512, cycle=False, emb_nn="lpdnet", pointer="transformer", vcp_nn="topK",
             t3d=False, tfea=False, n_blocks=1, dropout=0.0, ff_dims=1024, n_heads=4,
             overlap2=overlap2, partial=partial, num_points=1024, iter=1, model="vcrnet", loss="point",
             max_iterations=50)
    a.update(kw)
    return Namespace(**a)
