"""No-op `matplotlib.pyplot` (see the package docstring)."""
import base64

_PNG = base64.b64decode(b"iVBORw0KGgoAAAANSUhEUgAAAAEAAAABCAQAAAC1HAwCAAAAC0lEQVR42mNkYAAAAAYAAjCB0C8AAAAASUVORK5CYII=")


class _Null:
    """Absorbs any attribute access, call, index or iteration of a figure / axes / artist."""

    def __getattr__(self, name):
        return _Null()

    def __call__(self, *a, **k):
        return _Null()

    def __getitem__(self, key):
        return _Null()

    def __iter__(self):
        return iter(())


def _write_png(path):
    if isinstance(path, (str, bytes)) or hasattr(path, "__fspath__"):
        with open(path, "wb") as f:
            f.write(_PNG)
    elif hasattr(path, "write"):
        path.write(_PNG)


class _Figure(_Null):
    def savefig(self, path, *a, **k):
        _write_png(path)


def figure(*a, **k):
    return _Figure()


def subplots(nrows=1, ncols=1, *a, **k):
    fig = _Figure()
    if nrows == 1 and ncols == 1:
        return fig, _Null()
    if nrows == 1 or ncols == 1:
        return fig, [_Null() for _ in range(nrows * ncols)]
    return fig, [[_Null() for _ in range(ncols)] for _ in range(nrows)]


def savefig(path, *a, **k):
    _write_png(path)


def __getattr__(name):          # GridSpec, tight_layout, close, axvline, axhline, plot, ...
    return _Null()
