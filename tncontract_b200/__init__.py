"""tncontract_b200 -- the hot path of tncontract (labelled contraction, QR,
truncated SVD, MPS/MPO sweeps, boundary-MPS contraction) on NVIDIA B200.

Same public names as the reference package (tncontract/__init__.py:18-25):
``Tensor``, ``contract``, ``tensor_svd``, ``truncated_svd``, ``con``, the
``onedim`` and ``twodim`` sub-packages.  All arithmetic runs in libtnb.so
(hand-written sm_100a CUDA behind the C ABI of include/tnb.h); importing the
package needs the built library, using it needs a CUDA device -- there is no
CPU fallback.
"""
from .version import __version__
from . import _lib

_lib.load()  # fail loudly at import time when libtnb.so has not been built

from .tensor import *            # noqa: E402,F401,F403
from .tensor import tensor_qr, tensor_lq, conjugate, ToContract  # noqa: E402,F401
from .label import *             # noqa: E402,F401,F403
from .tncon import con           # noqa: E402,F401
from .devarray import DevArray, launch_count  # noqa: E402,F401
from . import tensor             # noqa: E402,F401
from . import matrices           # noqa: E402,F401
from . import onedim             # noqa: E402,F401
from . import twodim             # noqa: E402,F401
