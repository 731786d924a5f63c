"""Generate the committed golden fixtures tests/golden/*.npz by running the UNMODIFIED
reference (oracle/_ref/libhifir_ref.so, built from /root/reference by oracle/Makefile).

Run where /root/reference exists:   python tests/golden/make_golden.py

Each fixture holds the inputs (A in CSR, right-hand sides), the per-level factors exactly
as they sit in hif::Prec after hif::HIF::factorize, and the reference's own outputs:
HIF::solve (numerical and full last-level rank), HIF::hifir (both variants),
gmres_hif / fgmres_hifir (solution, flag, iterations, num_mv) and HIF::mmultiply.
qrcp_kat.npz carries the reference's known-answer test for the dense level
(tests/test_sss_qrcp.cpp: 20x20 MATLAB system) together with the QRCP state LAPACK produced.
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

from hifir_b200 import problems as P  # noqa: E402
from hifir_b200.levels_io import levels_to_arrays  # noqa: E402
from oracle import refhost as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def make_case(name, A, params, dense_thres=0, nsp=False, b_krylov=None, restart=30):
    n = A[0]
    M = R.RefHif(A, params, dense_thres=dense_thres)
    if nsp:
        M.set_nsp_const()
    levels = M.levels()
    out = levels_to_arrays(levels)
    out.update(A_n=np.array(n), A_indptr=A[1], A_indices=A[2], A_vals=A[3], nsp=np.array(int(nsp)),
               restart=np.array(restart))
    B = P.seeded_rhs(n, 0, nrhs=4)
    out["B"] = B
    out["X"] = np.stack([M.solve(B[:, k].copy()) for k in range(4)], axis=1)
    out["X_full"] = np.stack([M.solve(B[:, k].copy(), R.FULL_RANK) for k in range(4)], axis=1)
    b = B[:, 0].copy()
    if not nsp:  # the other lhf?Apply operations: S^H, M, M^H (numerical and full rank)
        for op, key in ((1, "x_SH"), (2, "x_M"), (3, "x_MH")):
            out[key] = M.apply_op(op, b)
            out[key + "_full"] = M.apply_op(op, b, R.FULL_RANK)
    out["x_hifir3"] = M.hifir(b, 3)
    if b_krylov is None:
        b_krylov = P.csr_matvec(A, np.ones(n))
    out["b_krylov"] = b_krylov
    xb, it, fl = M.hifir_betas(b_krylov, 16, [1e-8, 1e10])
    out["x_hifir_betas"], out["hifir_betas_status"] = xb, np.array([it, fl])
    for which in ("fgmres", "gmres"):
        x, flag, iters, nmv = M.krylov(b_krylov, which, restart=restart)
        out["x_" + which], out[which + "_status"] = x, np.array([flag, iters, nmv])
    if not nsp:
        out["Mx"] = M.mmultiply(out["X"][:, 0].copy())  # libhifir round trip: M (M^-1 b) = b
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    sizes = [(L["m"], L["n"], L["dense_n"]) for L in levels]
    print(f"{name}: n={n} levels={sizes} fgmres={out['fgmres_status']} gmres={out['gmres_status']} "
          f"betas={out['hifir_betas_status']} -> {os.path.getsize(path) / 1e6:.2f} MB")


def make_case_f32(name, A, params, dense_thres=0, nsp=False):
    """Single-precision preconditioner hif::HIF<float,int> factorized from the double matrix (the
    mixed-precision set-up of examples/intermediate/demo_mixedprecision.cpp): float factors and the
    reference's outputs of lhfsdApply (double vectors: S, S^H, M, M^H, refinement with the double
    matrix) and lhfsSolve (float vectors)."""
    n = A[0]
    M = R.RefHif(A, params, dense_thres=dense_thres, dtype=np.float32)
    if nsp:
        M.set_nsp_const()
    levels = M.levels()
    out = levels_to_arrays(levels)
    assert out["lv0_L_va"].dtype == np.float32 and out["lv0_s"].dtype == np.float32
    out.update(A_n=np.array(n), A_indptr=A[1], A_indices=A[2], A_vals=A[3], nsp=np.array(int(nsp)),
               restart=np.array(30))
    B = P.seeded_rhs(n, 0, nrhs=4)
    out["B"] = B
    out["X"] = np.stack([M.solve(B[:, k].copy()) for k in range(4)], axis=1)
    out["X_full"] = np.stack([M.solve(B[:, k].copy(), R.FULL_RANK) for k in range(4)], axis=1)
    out["X32"] = np.stack([M.solve_f32(B[:, k].astype(np.float32)) for k in range(4)], axis=1)
    b = B[:, 0].copy()
    if not nsp:
        for op, key in ((1, "x_SH"), (2, "x_M"), (3, "x_MH")):
            out[key] = M.apply_op(op, b)
    out["x_hifir3"] = M.hifir(b, 3)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    sizes = [(L["m"], L["n"], L["dense_n"]) for L in levels]
    print(f"{name}: n={n} levels={sizes} (float factors) -> {os.path.getsize(path) / 1e6:.2f} MB")


def qrcp_kat():
    src = open(os.path.join(REF, "tests", "test_sss_qrcp.cpp")).read()

    def arr(name, count):
        m = re.search(name + r"\[" + str(count) + r"\]\s*=\s*\{([^}]*)\}", src)
        vals = np.array([float(t) for t in m.group(1).replace("\n", " ").split(",") if t.strip()])
        assert vals.size == count
        return vals

    a, b, x_ref = arr("a", 400).reshape(20, 20), arr("b", 20), arr("x_ref", 20)
    mat, tau, jpvt, rank, x = R.qrcp_factor_solve(a, b)
    assert np.abs(x - x_ref).max() < 1e-10  # the reference's own assertion
    np.savez_compressed(os.path.join(HERE, "qrcp_kat.npz"), a=a, b=b, x_ref=x_ref, mat=mat, tau=tau, jpvt=jpvt,
                        rank=np.array(rank), x_lapack=x)
    print("qrcp_kat: rank", rank, "max|x-x_ref|", np.abs(x - x_ref).max())


def f32_cases():
    make_case_f32("poisson14_ml_f32", P.poisson3d(14), P.PDE_PARAMS, dense_thres=40)
    make_case_f32("stokes28_ml_f32", P.stokes2d_mac(28), P.PDE_PARAMS, dense_thres=60)
    make_case_f32("neumann12_nsp_f32", P.neumann3d(12), P.PDE_PARAMS, dense_thres=100, nsp=True)


if __name__ == "__main__":
    if "--f32-only" in sys.argv:  # adds the single-precision fixtures without touching the others
        f32_cases()
        sys.exit(0)
    f32_cases()
    qrcp_kat()
    A = P.read_matrix_market(os.path.join(REF, "examples/demo_inputs/A.mm"))
    b = P.read_matrix_market(os.path.join(REF, "examples/demo_inputs/b.mm"))
    make_case("demo_A", A, None, b_krylov=b)                       # config 1, default params
    make_case("poisson14_ml", P.poisson3d(14), P.PDE_PARAMS, dense_thres=40)
    make_case("convdiff14_ml", P.convdiff3d(14), P.PDE_PARAMS, dense_thres=60)
    make_case("stokes28_ml", P.stokes2d_mac(28), P.PDE_PARAMS, dense_thres=60, restart=10)
    An = P.neumann3d(12)
    make_case("neumann12_nsp", An, P.PDE_PARAMS, dense_thres=100, nsp=True,
              b_krylov=P.csr_matvec(An, np.sin(0.37 * np.arange(An[0]))))
