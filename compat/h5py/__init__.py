"""Stand-in for the slice of the `h5py` API the reference's I/O uses (dataloader.py:41-51,161-169;
one_step_test.py:36-60; render_rollout.py:184-188; generate_metadata.py:7-13; rollout_conversion.py:80-106):
`File(path, mode)` as a context manager, `keys()`, `f[name]` datasets with `.shape` / `.ndim` / `.dtype` / slicing,
`create_dataset`, `attrs`.  Used only when h5py itself is not installed (it is not in this image and no HDF5 library
is on disk); a real install found elsewhere on sys.path takes this module's place.

Storage is a NumPy `.npz` archive under the file name the caller gave (so `*.hdf5` globs keep working); this is a
container for synthetic test data, NOT an HDF5 reader -- real simulation files need the real h5py.  Not on the hot path.
"""
import json
import os
import sys

from _shim import prefer_real  # noqa: E402

if not prefer_real(__name__):
    import numpy as np

    __version__ = "0.0+cgnn-npz-shim"
    _ATTRS = "__attrs__"

    class Dataset:
        def __init__(self, array, attrs=None):
            self._a = array
            self.attrs = {} if attrs is None else attrs

        shape = property(lambda self: self._a.shape)
        ndim = property(lambda self: self._a.ndim)
        dtype = property(lambda self: self._a.dtype)
        size = property(lambda self: self._a.size)

        def __len__(self):
            return len(self._a)

        def __getitem__(self, idx):
            return self._a[idx]

        def __setitem__(self, idx, value):
            self._a[idx] = value

        def __array__(self, dtype=None, copy=None):
            return np.asarray(self._a, dtype=dtype)

    class File:
        def __init__(self, name, mode="r", **_):
            self.filename, self.mode = str(name), mode
            self._sets, self.attrs, self._open = {}, {}, True
            if mode in ("r", "r+", "a") and (mode != "a" or os.path.exists(self.filename)):
                with np.load(self.filename, allow_pickle=False) as z:
                    meta = json.loads(str(z[_ATTRS])) if _ATTRS in z.files else {"file": {}, "sets": {}}
                    self.attrs = meta.get("file", {})
                    for key in z.files:
                        if key != _ATTRS:
                            self._sets[key] = Dataset(z[key], meta.get("sets", {}).get(key, {}))
            elif mode not in ("w", "w-", "x", "a"):
                raise ValueError(f"h5py shim: unsupported mode {mode!r}")

        # -- reading ------------------------------------------------------------------------------------
        def keys(self):
            return self._sets.keys()

        def __iter__(self):
            return iter(self._sets)

        def __contains__(self, key):
            return key in self._sets

        def __getitem__(self, key):
            return self._sets[key]

        def __len__(self):
            return len(self._sets)

        # -- writing ------------------------------------------------------------------------------------
        def create_dataset(self, name, shape=None, dtype=None, data=None, **_):
            if data is None:
                data = np.zeros(shape, dtype=dtype or np.float32)
            arr = np.array(data, dtype=dtype) if dtype is not None else np.array(data)
            self._sets[name] = Dataset(arr)
            return self._sets[name]

        def __setitem__(self, name, data):
            self.create_dataset(name, data=data)

        def flush(self):
            if self.mode == "r":
                return
            def plain(d):
                return {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in d.items()}
            meta = {"file": plain(self.attrs), "sets": {k: plain(d.attrs) for k, d in self._sets.items() if d.attrs}}
            with open(self.filename, "wb") as fh:
                np.savez(fh, **{k: d._a for k, d in self._sets.items()}, **{_ATTRS: np.array(json.dumps(meta))})

        def close(self):
            if self._open:
                self.flush()
                self._open = False

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            self.close()
            return False
