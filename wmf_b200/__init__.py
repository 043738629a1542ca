"""wmf_b200 -- B200 (sm_100a) drop-in for the two data-parallel hot paths of nicolas998/WMF.

    from wmf_b200 import cu, models      # instead of `from cu import *; from models import *` (wmf.py:19-20)

``cu`` mirrors the f2py module built from wmf/cuencas.f90 (D8 topology: basin cut, downstream
table, flow accumulation, HAND ...), ``models`` the one built from wmf/modelosv2.f90 (the
SHIA/TETIS time loop).  Both are thin ctypes shims over libwmf_b200.so (hand-written CUDA,
C ABI in include/wmf_b200.h); PyTorch is used for device buffers and torch.distributed only.
There is no CPU fallback: calls raise when the library or a CUDA device is missing.
"""
from ._cu import CuModule
from ._lib import Context, WmfError, default_context
from ._models import ModelsModule

cu = CuModule()
models = ModelsModule()

__all__ = ["cu", "models", "CuModule", "ModelsModule", "Context", "WmfError", "default_context"]
__version__ = "0.1.0"
