"""sgcdet_b200: B200-native (sm_100a) implementation of SGCDet's view-transform hot path.

Public surface:
  * ``sgcdet_b200.plugin``     -- drop-in ``AdaptiveSparseHead`` / ``DenseHead`` / ``*_DFA3D`` modules (boundary B1)
  * ``sgcdet_b200.dropin``     -- a ``dfa3D`` package with the reference's ``ext_loader`` / ``_ext`` / ``ops``
                                  surface (boundary B2); ``install_dropin()`` puts it on ``sys.path``
  * ``sgcdet_b200.functional`` -- autograd Functions over the C ABI declared in ``include/sgcdet_b200.h``
  * ``sgcdet_b200.synthetic``  -- deterministic synthetic scenes / weights / roofline accounting
There is no CPU path: the CUDA library must be built (``python -m sgcdet_b200.build``) and inputs must be
CUDA tensors, otherwise calls raise.
"""
import os
import sys

__version__ = '0.1.0'


def install_dropin() -> str:
    """Make ``import dfa3D`` resolve to the sgcdet_b200 implementation (same names as
    packages/3D-deformable-attention/DFA3D/dfa3D)."""
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'dropin')
    if d not in sys.path:
        sys.path.insert(0, d)
    return d
