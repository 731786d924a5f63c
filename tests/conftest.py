import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ("demo_A", "poisson14_ml", "convdiff14_ml", "stokes28_ml", "neumann12_nsp")

# single-precision preconditioners hif::HIF<float,int> (lhfs* / mixed lhfsd* entry points)
GOLDEN_F32_CASES = ("poisson14_ml_f32", "stokes28_ml_f32", "neumann12_nsp_f32")

TOL_F64 = 1e-12  # north_star: ||x_gpu - x_ref|| / ||x_ref|| <= 1e-12 in double
TOL_F32 = 1e-5   # north_star: <= 1e-5 in float (against the reference's single-precision apply)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb else 1.0)


class Golden:
    """One committed fixture: inputs, per-level factors and the reference's outputs."""

    def __init__(self, name):
        from hifir_b200.levels_io import arrays_to_levels
        self.name = name
        with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
            self.d = {k: z[k] for k in z.files}
        self.levels = arrays_to_levels(self.d)
        self.A = (int(self.d["A_n"]), self.d["A_indptr"], self.d["A_indices"], self.d["A_vals"])
        self.nsp = bool(int(self.d["nsp"]))
        self.restart = int(self.d["restart"])
        self.n = self.A[0]

    def __getitem__(self, k):
        return self.d[k]


_cache = {}


def load_golden(name):
    if name not in _cache:
        _cache[name] = Golden(name)
    return _cache[name]


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return load_golden(request.param)


@pytest.fixture(params=GOLDEN_F32_CASES)
def golden_f32(request):
    return load_golden(request.param)


def have_reference():
    from oracle import refhost
    if not refhost.available():
        return False
    try:
        refhost.lib()
        return True
    except OSError:
        return False


def gpu_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
