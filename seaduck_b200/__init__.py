"""seaduck_b200 -- B200-native implementation of seaduck's Lagrangian / interpolation hot path.

Drop-in for the hot path of ``seaduck`` (same names: ``OceData``, ``Particle``, ``Position``,
``KnW``, ``OceInterp``, ``Topology``, ``rcParam``); the work runs in hand-written sm_100a CUDA
kernels behind the C ABI of ``include/seaduck_b200.h``.
"""

__version__ = "0.1.0"

from .runtime_conf import rcParam  # noqa: E402
from .topology import Topology  # noqa: E402
from .kernel_weight import KnW  # noqa: E402
from .ocedata import OceData  # noqa: E402
from .eulerian import Position  # noqa: E402
from .lagrangian import Particle  # noqa: E402
from .oceinterp import OceInterp  # noqa: E402
from . import get_masks  # noqa: E402,F401
from .interp import pinned  # noqa: E402  (numpy arrays in page-locked memory for the host-buffer entry points)

__all__ = ["__version__", "Position", "KnW", "Particle", "OceData", "OceInterp", "rcParam", "Topology", "pinned"]
