import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

C2_ORBSYM = [1, 5, 3, 2, 1, 7, 6, 5, 1, 2, 3, 1, 6, 7, 5, 4, 1, 5, 3, 2, 8, 5, 1, 7, 6, 5]
C2_FCIDUMP = os.path.join(ROOT, "data", "C2_v2z_curve", "r1.24253", "FCIDUMP")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def heg_space(oracle):
    """HEG 14e r_s=0.5 cutoff 1.49 (reference e2e test), 2 HCI iterations -> 9475 dets."""
    s = oracle.System.heg(3, 0.5, 14, 7, 1.49)
    r = s.hci(1e-3, n_states=1, max_iters=2)
    return s, r


@pytest.fixture(scope="session")
def c2_space_ts(oracle):
    """C2 cc-pVDZ default input (time_sym=t, z=1), HCI with the shipped eps_var schedule, 3 iterations."""
    s = oracle.System.chem(C2_FCIDUMP, 26, 8, 4, C2_ORBSYM, time_sym=True, z=1, hf_symmetry=1)
    r = s.hci(1e-3, eps_var_sched=[2e-3, 2e-3], n_states=1, max_iters=3)
    return s, r


@pytest.fixture(scope="session")
def c2_space(oracle):
    s = oracle.System.chem(C2_FCIDUMP, 26, 8, 4, C2_ORBSYM, time_sym=False, z=1, hf_symmetry=1)
    r = s.hci(1e-3, eps_var_sched=[2e-3, 2e-3], n_states=1, max_iters=3)
    return s, r


@pytest.fixture(scope="session")
def c2_hci_full(oracle):
    """the complete oracle HCI run on the shipped C2 input (time_sym=t, n_states=1): final det list + per-iteration log"""
    s = oracle.System.chem(C2_FCIDUMP, 26, 8, 4, C2_ORBSYM, time_sym=True, z=1, hf_symmetry=1)
    r = s.hci(1e-3, eps_var_sched=[2e-3, 2e-3], n_states=1)
    return s, r


def label_sorted(r):
    """the variational wavefunction of an oracle HCI result in label order (signed 128-bit compare on up, then dn; the order
    perform_hci establishes before the PT stage, hci.f90:556-600) -> (up, dn, wts of state 1)"""
    def sgn(x):
        v = int(x[1]) << 64 | int(x[0])
        return v - (1 << 128) if v >> 127 else v
    key = [(sgn(u), sgn(d)) for u, d in zip(r["up"], r["dn"])]
    o = np.array(sorted(range(len(key)), key=lambda i: key[i]))
    return r["up"][o], r["dn"][o], r["wts"][o, 0]
