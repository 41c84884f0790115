"""sqmc_b200 -- B200-native (sm_100a CUDA + NCCL) sparse Hamiltonian build and H.v
for QMC-Cornell/sqmc's HCI / semistochastic hot path.  See DESIGN.md.

The compute path lives in libsqmc_b200.so (csrc/, C ABI in include/sqmc_b200.h);
this package is the thin host-side mirror of the reference's interface.
"""
from . import _lib, systems  # noqa: F401
from ._lib import SqmcError  # noqa: F401
from .api import SparseHamiltonian, dets_to_u64  # noqa: F401
from .systems import ChemSystem, HegSystem, HubbardKSystem  # noqa: F401
