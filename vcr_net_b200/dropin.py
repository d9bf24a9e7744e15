"""Drop-in installer: make the reference's unchanged ``main.py`` run on the B200-native path.

The reference resolves its hot-path classes by import path (``main.py:6-10``, ``util/initPara.py:18-22``:
``model.vcrnet_model.VCRNet``, ``model.lpdnet_model.LPD``, ``util.util.knn`` ...).  ``install()`` imports the
reference's own ``model`` / ``util`` packages (so its loops, datasets and argparse stay byte-for-byte
the reference's) and rebinds exactly the hot-path symbols to this package's implementations.

    python -m vcr_net_b200.dropin /path/to/VCR-Net --eval --model_path=pretrained/vcrnet-whole.t7 ...
"""
from __future__ import annotations

import importlib
import os
import runpy
import sys

HOT_PATH = {
    "model.vcrnet_model": ["VCRNet", "VcpTopK", "VcpAtt", "VcpByDis", "SVDHead", "Identity", "vcrnetIter", "vcrnetIcpNet", "DGCNN",
                           "PointNet"],
    "model.lpdnet_model": ["LPDNet", "LPD", "TranformNet"],
    "model.icp_model": ["ICP"],
    "model.transformer": ["Transformer", "MultiHeadedAttention", "PositionwiseFeedForward", "LayerNorm",
                          "EncoderDecoder", "Encoder", "Decoder", "EncoderLayer", "DecoderLayer",
                          "SublayerConnection", "clones"],
    "util.util": ["knn", "get_graph_feature", "farthest_point_sample", "transform_point_cloud", "quat2mat",
                  "npmat2euler"],
}
_OURS = {
    "model.vcrnet_model": "vcr_net_b200.model.vcrnet_model",
    "model.lpdnet_model": "vcr_net_b200.model.lpdnet_model",
    "model.icp_model": "vcr_net_b200.model.icp_model",
    "model.transformer": "vcr_net_b200.model.transformer",
    "util.util": "vcr_net_b200.util.util",
}


def install(reference_root: str):
    """Import the reference packages from ``reference_root`` and rebind the hot-path symbols.
    Returns {reference module name: [rebound symbols]}."""
    reference_root = os.path.abspath(reference_root)
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    done = {}
    # util.util first: model.* does `from util.util import knn, ...` at import time
    for ref_name in ("util.util", "model.transformer", "model.lpdnet_model", "model.icp_model", "model.vcrnet_model"):
        ours = importlib.import_module(_OURS[ref_name])
        ref = importlib.import_module(ref_name)
        for sym in HOT_PATH[ref_name]:
            setattr(ref, sym, getattr(ours, sym))
        done[ref_name] = list(HOT_PATH[ref_name])
    # modules that copied the names with `from X import Y` before the rebind
    for mod in list(sys.modules.values()):
        f = getattr(mod, "__file__", None) or ""
        if not f.startswith(reference_root):
            continue
        for ref_name, syms in HOT_PATH.items():
            ours = sys.modules[_OURS[ref_name]]
            for sym in syms:
                if sym in getattr(mod, "__dict__", {}) and getattr(getattr(mod, sym), "__module__", "").startswith(
                        ("model.", "util.")):
                    setattr(mod, sym, getattr(ours, sym))
    return done


def main():
    if len(sys.argv) < 2:
        raise SystemExit("usage: python -m vcr_net_b200.dropin <reference_root> [main.py args...]")
    root = os.path.abspath(sys.argv[1])
    install(root)
    sys.argv = [os.path.join(root, "main.py")] + sys.argv[2:]
    os.chdir(root)
    runpy.run_path(sys.argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
