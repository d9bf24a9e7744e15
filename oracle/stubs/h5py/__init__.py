"""Harness-side stand-in for ``h5py`` (absent from this image).  TEST INFRASTRUCTURE ONLY.

The reference's loader (util/data.py:30-47) only does ``f = h5py.File(name, mode='r'); f['data'][:]; f['label'][:];
f.close()``.  The synthetic dataset written by oracle/build_ref.make_dataset stores an .npz payload under the *.h5
names, which this reads.  Put on sys.path by oracle/ref_harness only when the real h5py cannot be imported."""
import numpy as np


class File:
    def __init__(self, name, mode="r"):
        if mode != "r":
            raise NotImplementedError("stub h5py: read-only")
        self._z = np.load(name)

    def __getitem__(self, key):
        return self._z[key]

    def close(self):
        self._z.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
