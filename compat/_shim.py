"""Helper for the import shims in this directory: a shim only stands in when the real package is absent.

`prefer_real(name, shim_file)` looks for `name` on sys.path OUTSIDE this directory; when it is there, the real
package is imported in the shim's place (sys.modules[name] is replaced) and True is returned."""
import importlib.machinery
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def prefer_real(name: str) -> bool:
    paths = [p for p in sys.path if os.path.abspath(p or ".") != _HERE]
    spec = importlib.machinery.PathFinder.find_spec(name, paths)
    if spec is None or spec.loader is None:
        return False
    module = importlib.util.module_from_spec(spec)
    sys.modules[name] = module
    spec.loader.exec_module(module)
    return True
