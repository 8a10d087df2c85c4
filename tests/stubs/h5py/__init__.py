"""Minimal stand-in for h5py (absent from this image, no network) for running the reference's
scripts/datagen_denoise.py UNCHANGED as an acceptance test: File(name, 'w').create_dataset(name, shape=,
dtype=) returning an array that supports `ds[i] = image`. The data are memory-mapped .npy files next to the would-be
.h5 file (<file>.<dataset>.npy), so that the test can read back what the script wrote."""
import numpy as np


class File:
    def __init__(self, name, mode="r", **kw):
        self.filename, self.mode, self._sets = name, mode, {}
        open(name, "ab").close()           # the script checks / removes the .h5 path itself

    def create_dataset(self, name, shape=None, dtype=None, data=None, **kw):
        arr = np.lib.format.open_memmap(f"{self.filename}.{name}.npy", mode="w+", dtype=np.dtype(dtype), shape=tuple(int(s) for s in shape))
        if data is not None:
            arr[...] = data
        self._sets[name] = arr
        return arr

    def __getitem__(self, name):
        return self._sets[name]

    def flush(self):
        for a in self._sets.values():
            a.flush()

    def close(self):
        self.flush()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
