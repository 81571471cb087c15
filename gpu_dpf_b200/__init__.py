"""Importable name of the engine's Python package.

The sources live in `gpu-dpf_b200/` (the directory name this build was given; not a valid Python
identifier).  `import gpu_dpf_b200` makes that directory this package's search path, so

    from gpu_dpf_b200 import dpf, b200dpf, sharded
    d = dpf.DPF(prf=dpf.DPF.PRF_AES128)

work from the repository root with no sys.path edits by the caller.  The directory is also put on
sys.path once, because the extension module must stay importable under its top-level name
`dpf_cpp` -- that name is the reference's drop-in boundary (dpf.py:7 `import dpf_cpp`).
"""
import os
import sys

_SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpu-dpf_b200")
__path__.append(_SRC)
if _SRC not in sys.path:
    sys.path.insert(0, _SRC)

PACKAGE_DIR = _SRC
