"""Run the reference's UNMODIFIED ``main.py --eval`` end to end -- stock, or with the B200 drop-in installed.

TEST INFRASTRUCTURE ONLY (VERDICT r1 "missing #2").  The reference's entry needs things this image
does not have; all of them are supplied from the harness side, none by product code:

  * ``tensorboardX`` / ``h5py``          -> oracle/stubs/ (only when the real packages are missing)
  * the ModelNet40 HDF5 files            -> oracle/_ref/dataset/modelnet40_ply_hdf5_2048/ply_data_{train,test}0.h5, a
                                            48-item synthetic test partition written by oracle/build_ref.make_dataset
  * ``pretrained/vcrnet-*.t7``           -> absent upstream (.MISSING_LARGE_BLOBS); a synthetic 59-key checkpoint in the
                                            reference's legacy-pickle format is written next to lpd-pretrained.t7
  * a GPU                                -> ``--cpu`` maps ``.cuda()`` to the identity (ref_harness.cpu_mode) so the same
                                            entry runs on host cores (how the golden metrics are reproduced here)

    python -m oracle.run_main [--dropin] [--cpu] -- --eval --partial --overlap 0.575 --iter 3 --test_batch_size 24

The last stdout line is one JSON object: the metrics parsed from the reference's own "EPOCH::" report line(s)
(model/vcrnet_model.py:783-799) plus wall time.  ``--dropin`` calls vcr_net_b200.dropin.install(root) first, so every
hot-path symbol main.py resolves (VCRNet, vcrnetIter, knn, ...) is the B200-native one while the loop, the dataset, argparse,
nn.DataParallel and the metric code stay the reference's own bytes.
"""
from __future__ import annotations

import io
import json
import os
import re
import runpy
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REPORT = re.compile(r"EPOCH:: (-?\d+), Loss: (\S+), test_LossPose: (\S+), Cycle Loss: (\S+), MSE: (\S+), RMSE: (\S+), "
                    r"MAE: (\S+), rot_MSE: (\S+), rot_RMSE: (\S+), rot_MAE: (\S+), trans_MSE: (\S+), trans_RMSE: (\S+), "
                    r"trans_MAE: (\S+)")
KEYS = ("loss", "loss_pose", "cycle_loss", "mse_ab", "rmse_ab", "mae_ab", "r_mse_ab", "r_rmse_ab", "r_mae_ab", "t_mse_ab",
        "t_rmse_ab", "t_mae_ab")


def write_checkpoint(root: str, name: str = "vcrnet-synth.t7") -> str:
    """The synthetic 59-key state_dict in the reference's .t7 layout (torch.save, legacy non-zip pickle)."""
    import numpy as np
    import torch
    from oracle import synth
    lpd = dict(np.load(os.path.join(ROOT, "tests", "golden", "lpd_pretrained_weights.npz")))
    sd = synth.checkpoint_to_torch(synth.make_checkpoint(1234, emb_weights=lpd))
    path = os.path.join(root, "pretrained", name)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save(sd, path, _use_new_zipfile_serialization=False)
    return path


class _Tee(io.TextIOBase):
    def __init__(self, real):
        self.real, self.lines, self._buf = real, [], ""

    def write(self, s):
        self.real.write(s)
        self._buf += s
        while "\n" in self._buf:
            ln, self._buf = self._buf.split("\n", 1)
            self.lines.append(ln)
        return len(s)

    def flush(self):
        self.real.flush()


def run_main(argv, dropin=False, cpu=False):
    from oracle import build_ref, ref_harness
    if not build_ref.staged():
        build_ref.build(verbose=False)
    root = build_ref.STAGED_ROOT
    if not build_ref.staged():
        raise RuntimeError("reference not staged: run `python -m oracle.build_ref` where /root/reference exists")
    if not os.path.isfile(os.path.join(build_ref.DATASET_DIR, "ply_data_test0.h5")):
        build_ref.make_dataset()
    ckpt = write_checkpoint(root)
    if not any(a.startswith("--model_path") for a in argv):
        argv = list(argv) + ["--model_path", os.path.relpath(ckpt, root)]
    ref_harness.REF_ROOT = root
    ref_harness._install_shims()
    ref_harness._purge_foreign_packages()
    sys.dont_write_bytecode = True
    if root not in sys.path:
        sys.path.insert(0, root)
    installed = None
    if dropin:
        from vcr_net_b200 import dropin as D
        installed = D.install(root)
    old_argv, old_cwd, old_out = sys.argv, os.getcwd(), sys.stdout
    tee = _Tee(old_out)
    sys.argv = [os.path.join(root, "main.py")] + list(argv)
    os.chdir(root)
    sys.stdout = tee
    t0 = time.perf_counter()
    try:
        if cpu:
            with ref_harness.cpu_mode():
                runpy.run_path(sys.argv[0], run_name="__main__")
        else:
            runpy.run_path(sys.argv[0], run_name="__main__")
    finally:
        sys.stdout = old_out
        sys.argv = old_argv
        os.chdir(old_cwd)
    wall = time.perf_counter() - t0
    reports = [dict(zip(KEYS, map(float, m.groups()[1:]))) for m in map(REPORT.search, tee.lines) if m]
    return {"impl": "dropin" if dropin else "reference", "device": "cpu" if cpu else "cuda", "argv": list(argv),
            "wall_s": wall, "metrics": reports[0] if reports else None, "finished": any("FINISH" in ln for ln in tee.lines),
            "rebound": {k: len(v) for k, v in installed.items()} if installed else None}


def main():
    args = sys.argv[1:]
    rest = args[args.index("--") + 1:] if "--" in args else []
    head = args[:args.index("--")] if "--" in args else args
    res = run_main(rest, dropin="--dropin" in head, cpu="--cpu" in head)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
