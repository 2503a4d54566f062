"""B200-native Fast Marching engine behind the drop-in ``FastMarching`` package.

Layout: ``csrc/`` hand-written sm_100a CUDA + the C ABI (``include/fm_b200.h``),
``_capi`` ctypes binding, ``engine`` torch-plumbed device API, ``batch`` multi-GPU
query sharding, ``synth`` seeded synthetic inputs, ``build`` in-tree nvcc build.
"""
from . import build as _build_module

build_library = _build_module.build
LIB_PATH = _build_module.LIB_PATH

__all__ = ["build_library", "LIB_PATH"]
