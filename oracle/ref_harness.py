"""Import the LIVE reference (read-only /root/reference) with harness-only shims.

TEST INFRASTRUCTURE ONLY, and usable only in the build container: /root/reference does not
exist on the GPU box, so nothing run there (``-m gpu`` tests, smoke(), bench.py) may import
this module.  It exists to (1) generate tests/golden/ with oracle/make_golden.py and
(2) let ``tests/test_ref_live.py`` (skipped when the reference is absent) re-validate the
oracle against the live reference.

Shims: a stub ``pynvml`` (util/util.py:13 calls nvmlInit at import),
``Rotation.from_dcm`` -> ``from_matrix`` (util/util.py:102), ``torch.cuda.FloatTensor`` on CPU
(model/lpdnet_model.py:186).  No reference source is modified or copied.
"""
from __future__ import annotations

import os
import sys
import types
from argparse import Namespace

REF_ROOT = os.environ.get("VCR_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "model", "vcrnet_model.py"))


def _install_shims():
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


def import_reference():
    """Returns a namespace with the reference's hot-path modules."""
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    _install_shims()
    sys.dont_write_bytecode = True
    # the reference uses top-level package names ``model`` and ``util``
    for name in list(sys.modules):
        if name in ("model", "util") or name.startswith(("model.", "util.")):
            mod = sys.modules[name]
            if not getattr(mod, "__file__", "").startswith(REF_ROOT):
                del sys.modules[name]
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
             overlap2=overlap2, partial=partial, num_points=1024, iter=1, model="vcrnet")
    a.update(kw)
    return Namespace(**a)
