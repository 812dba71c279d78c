"""Same contract as ``dfa3D/ext_loader.py:25-29`` of the reference: ``load_ext('_ext', names)`` imports
``dfa3D._ext`` and asserts the requested functions exist."""
import importlib


def load_ext(name, funcs):
    ext = importlib.import_module('dfa3D.' + name)
    for fun in funcs:
        assert hasattr(ext, fun), f'{fun} miss in module {name}'
    return ext


def check_ops_exist() -> bool:
    try:
        importlib.import_module('dfa3D._ext')
        return True
    except Exception:
        return False
