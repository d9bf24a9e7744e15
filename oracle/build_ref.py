"""Stage the LIVE reference under oracle/_ref/ so it can travel to the GPU box.

TEST INFRASTRUCTURE ONLY.  The reference (qiaozhijian/VCR-Net) is pure Python: "building" it means placing its
unmodified files where a fresh box can import them.  ``/root/reference`` exists only in the build container, while
``oracle/_ref/`` is git-ignored (never enters history) but NOT gpurun-ignored, so it rides along with the repo snapshot
like the built ``.so`` files.  Nothing under ``vcr_net_b200/`` reads it; it is used by

  * ``bench.py --impl reference`` / ``cpu_baseline``  -> kind "reference": the reference's own ``vcrnetIter`` on the box's host cores
  * ``tests/test_ref_live.py``                        -> GPU path vs the live reference at BASELINE's sizes
  * ``oracle/run_main.py``                            -> the reference's unmodified ``main.py --eval`` (stock or drop-in)

Layout written:  oracle/_ref/VCR-Net/{main.py, model/, util/, pretrained/lpd-pretrained.t7}
                 oracle/_ref/dataset/modelnet40_ply_hdf5_2048/ply_data_{train,test}0.h5   (synthetic, see make_dataset)
                 oracle/_ref/MANIFEST.json   (sha256 of every staged reference file)

Run:  python -m oracle.build_ref          (no-op with a message when /root/reference is absent)
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("VCR_REFERENCE_ROOT", "/root/reference")
STAGE = os.path.join(HERE, "_ref")
STAGED_ROOT = os.path.join(STAGE, "VCR-Net")
DATASET_DIR = os.path.join(STAGE, "dataset", "modelnet40_ply_hdf5_2048")

FILES = ["main.py", "model/__init__.py", "model/dcp_model.py", "model/icp_model.py", "model/lpdnet_model.py",
         "model/transformer.py", "model/vcrnet_model.py", "util/__init__.py", "util/data.py", "util/fps.py",
         "util/icp.py", "util/initPara.py", "util/util.py", "pretrained/lpd-pretrained.t7"]

# synthetic ModelNet40-shaped dataset (util/data.py:30-47 layout: data [n,2048,3] f32, label [n,1]); the same base clouds
# oracle/synth.make_pairs draws (RandomState(1234).rand(n, 2048, 3) - 0.5, model/icp_model.py:124 distribution)
N_TEST, N_TRAIN, BASE_POINTS, DATA_SEED = 48, 8, 2048, 1234


def make_dataset(n_test: int = N_TEST, n_train: int = N_TRAIN) -> str:
    """Files named *.h5 holding an .npz payload: the harness-side stub ``h5py`` (oracle/stubs/h5py) reads them; the
    reference's loader (util/data.py:30-47) is unchanged and only ever calls File(...)['data'][:] / ['label'][:]."""
    os.makedirs(DATASET_DIR, exist_ok=True)
    base = np.random.RandomState(DATA_SEED).rand(n_test, BASE_POINTS, 3).astype(np.float32) - 0.5
    with open(os.path.join(DATASET_DIR, "ply_data_test0.h5"), "wb") as f:
        np.savez(f, data=base, label=np.zeros((n_test, 1), np.int64))
    tr = np.random.RandomState(DATA_SEED + 1).rand(n_train, BASE_POINTS, 3).astype(np.float32) - 0.5
    with open(os.path.join(DATASET_DIR, "ply_data_train0.h5"), "wb") as f:
        np.savez(f, data=tr, label=np.zeros((n_train, 1), np.int64))
    return DATASET_DIR


def staged() -> bool:
    return os.path.isfile(os.path.join(STAGED_ROOT, "model", "vcrnet_model.py"))


def build(verbose: bool = True) -> bool:
    """Copy the reference's files (unmodified) into oracle/_ref/VCR-Net and write the synthetic dataset next to it."""
    if not os.path.isfile(os.path.join(REF_SRC, "model", "vcrnet_model.py")):
        if verbose:
            print(f"oracle.build_ref: {REF_SRC} not present; keeping whatever is staged ({'yes' if staged() else 'nothing'})")
        return staged()
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(STAGED_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(dst):
            os.chmod(dst, 0o644)
        shutil.copyfile(src, dst)
        os.chmod(dst, 0o644)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    make_dataset()
    with open(os.path.join(STAGE, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_SRC, "files": manifest}, f, indent=1)
    if verbose:
        print(f"oracle.build_ref: staged {len(FILES)} reference files under {STAGED_ROOT}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
