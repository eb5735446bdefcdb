"""TEST INFRASTRUCTURE ONLY: a second instance of the product's ctypes binding (the source of pyrodigal_b200/_capi.py,
executed under another module name) bound to tests/emu/libpgpu_emu.so, the host emulation of the CUDA library."""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
_cached = None


def load():
    global _cached
    if _cached is None:
        sys.path.insert(0, HERE)
        import build_emu
        lib = build_emu.build()
        src_path = os.path.join(ROOT, "pyrodigal_b200", "_capi.py")
        with open(src_path) as f:
            src = f.read()
        marker = 'LIB_PATH = os.path.join(_HERE, "libpyrodigal_b200.so")'
        assert marker in src
        src = src.replace(marker, f"LIB_PATH = {lib!r}")
        spec = importlib.util.spec_from_loader("pgpu_emu_capi", loader=None)
        mod = importlib.util.module_from_spec(spec)
        mod.__file__ = src_path
        exec(compile(src, src_path, "exec"), mod.__dict__)
        _cached = mod
    return _cached
