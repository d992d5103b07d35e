"""lightweaver_b200 -- B200-native back end for Lightweaver's formal-solution /
Gamma-accumulation / statistical-equilibrium hot path.

Only what the path needs lives here: ``csrc/`` (CUDA kernels + the C-ABI of
``include/lwb200.h`` + the Lightweaver plugin shim), the ctypes binding
(``capi``), the host-side mirror of the reference's ``Context`` methods for this
path (``context``), the host data model (``problem``), the synthetic FAL C
inputs of BASELINE.json (``synth``) and the multi-GPU partitioning
(``sharding``).
"""
from . import capi
from .problem import AtomData, Problem, TransitionData

__all__ = ['capi', 'AtomData', 'Problem', 'TransitionData']
