"""CPU oracle for the TriTD-ADMM hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This file is a float64 numpy restatement, statement for statement, of the
reference MATLAB code (all paths relative to /root/reference):

  fast_robust_triple_tensor/triple_decomp_ADMM.m:15-68   main loop
  fast_robust_triple_tensor/triple_decomp_ADMM.m:73-95   update_A / update_B / update_C
  fast_robust_triple_tensor/triple_decomp_ADMM.m:97-109  unfold
  fast_robust_triple_tensor/triple_decomp_ADMM.m:111-130 reshape_*_from_*
  fast_robust_triple_tensor/triple_decomp_ADMM.m:132-160 buildF / buildG / buildH
  fast_robust_triple_tensor/triple_product.m:1-8
  fast_robust_triple_tensor/soft_threshold.m:2
  fast_robust_triple_tensor/triple_decomp_ALS.m:1-40     (ALS, "next" row of SURVEY 8f)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import it, and only as the checker / timed CPU baseline.
The product (libtritd.so) never calls into this file and has no CPU fallback.

PARITY STATUS: the reference ships no golden vectors, no known-answer tests
and no fixtures for this path, and neither MATLAB nor Octave is
available to run it.  What pins this oracle:
  * the reference's OWN, UNMODIFIED .m sources executed in the build container
    by the MATLAB-subset interpreter oracle/mlab.py (tests/golden/ref_m/*.npz,
    generator tests/golden/make_ref_golden.py, checked in
    tests/test_reference_pin.py): statement order, index conventions, operator
    association, function resolution and the printed lines come from the
    reference's source text;
  * the definitional scalar loops the reference holds as comments (buildF.m:5-16,
    buildG.m:5-16, buildH.m:5-16, origin_triple_tensor/triple_decomp_ADMM.m:125-143),
    restated below as pure-Python loops (``*_loops``), tests/test_oracle.py;
  * a MATLAB / Octave dump when one is dropped into tests/golden/matlab_*.mat
    (tools/reference_dump.m); none is committed.
STILL UNPINNED: MathWorks' built-ins themselves (pinv = LAPACK SVD with cutoff
max(size)*eps(norm(G)), mtimes = BLAS dgemm, norm = dnrm2; MATLAB version not
pinned by the reference) -- numpy / OpenBLAS / LAPACK stand in for them here.

All arrays are column-major (``order='F'``) like MATLAB's.
"""
from __future__ import annotations

import numpy as np

EPS = np.finfo(np.float64).eps


# --------------------------------------------------------------------------
# L2 tensor kernels
# --------------------------------------------------------------------------
def unfold(X: np.ndarray, mode: int) -> np.ndarray:
    """triple_decomp_ADMM.m:97-109 / unfold.m:1-14 (mode is 1-based)."""
    n1, n2, n3 = X.shape
    if mode == 1:
        return np.reshape(X, (n1, n2 * n3), order="F")
    if mode == 2:
        return np.reshape(np.transpose(X, (1, 0, 2)), (n2, n1 * n3), order="F")
    if mode == 3:
        return np.reshape(np.transpose(X, (2, 0, 1)), (n3, n1 * n2), order="F")
    raise ValueError("Mode must be 1, 2, or 3.")


def buildF(B: np.ndarray, C: np.ndarray) -> np.ndarray:
    """triple_decomp_ADMM.m:132-140.  B: r x n2 x r, C: r x r x n3 -> r^2 x (n2 n3)."""
    r, n2, _ = B.shape
    n3 = C.shape[2]
    B_unfold = np.reshape(unfold(B, 2), (n2, r * r, 1), order="F")
    C_unfold = np.reshape(unfold(C, 3).T, (1, r * r, n3), order="F")
    F = B_unfold * C_unfold
    F = np.reshape(F, (n2, r, r, n3), order="F")
    return np.reshape(np.transpose(F, (1, 2, 0, 3)), (r * r, n2 * n3), order="F")


def buildG(A: np.ndarray, C: np.ndarray) -> np.ndarray:
    """triple_decomp_ADMM.m:142-150.  A: n1 x r x r, C: r x r x n3 -> r^2 x (n1 n3)."""
    n1, r, _ = A.shape
    n3 = C.shape[2]
    A_unfold = np.reshape(unfold(A, 1), (n1, r * r, 1), order="F")
    C_unfold = np.reshape(unfold(C, 3).T, (1, r * r, n3), order="F")
    G = A_unfold * C_unfold
    G = np.reshape(G, (n1, r, r, n3), order="F")
    return np.reshape(np.transpose(G, (1, 2, 0, 3)), (r * r, n1 * n3), order="F")


def buildH(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """triple_decomp_ADMM.m:152-160.  A: n1 x r x r, B: r x n2 x r -> r^2 x (n1 n2)."""
    n1, r, _ = A.shape
    n2 = B.shape[1]
    A_unfold = np.reshape(unfold(A, 1), (n1, r * r, 1), order="F")
    B_unfold = np.reshape(unfold(B, 2).T, (1, r * r, n2), order="F")
    H = A_unfold * B_unfold
    H = np.reshape(H, (n1, r, r, n2), order="F")
    return np.reshape(np.transpose(H, (1, 2, 0, 3)), (r * r, n1 * n2), order="F")


def triple_product(A: np.ndarray, B: np.ndarray, C: np.ndarray) -> np.ndarray:
    """triple_product.m:1-8."""
    n1, n2, n3 = A.shape[0], B.shape[1], C.shape[2]
    Xhat = unfold(A, 1) @ buildF(B, C)
    return np.reshape(Xhat, (n1, n2, n3), order="F")


def soft_threshold(X: np.ndarray, lam: float) -> np.ndarray:
    """soft_threshold.m:2."""
    return np.sign(X) * np.maximum(np.abs(X) - lam, 0)


def pinv_matlab(G: np.ndarray) -> np.ndarray:
    """MATLAB pinv(G): SVD, singular values <= max(size(G))*eps(norm(G)) zeroed."""
    U, s, Vt = np.linalg.svd(G, full_matrices=False)
    tol = max(G.shape) * np.spacing(s[0]) if s.size else 0.0
    keep = s > tol
    sinv = np.zeros_like(s)
    sinv[keep] = 1.0 / s[keep]
    return (Vt.T * sinv) @ U.T


def pinv_truncations(G: np.ndarray) -> int:
    """How many singular values MATLAB's pinv would zero (diagnostic for tests)."""
    s = np.linalg.svd(G, compute_uv=False)
    return int(np.sum(s <= max(G.shape) * np.spacing(s[0])))


# reshape_*_from_* (triple_decomp_ADMM.m:111-130) -- pure relabelling
def reshape_A_from_A1(A1: np.ndarray, n1: int, r: int) -> np.ndarray:
    A = np.zeros((n1, r, r), order="F")
    for i in range(n1):
        A[i, :, :] = np.reshape(A1[i, :], (r, r), order="F")
    return A


def reshape_B_from_B2(B2: np.ndarray, n2: int, r: int) -> np.ndarray:
    B = np.zeros((r, n2, r), order="F")
    for j in range(n2):
        B[:, j, :] = np.reshape(B2[j, :], (r, r), order="F")
    return B


def reshape_C_from_C3(C3: np.ndarray, n3: int, r: int) -> np.ndarray:
    C = np.zeros((r, r, n3), order="F")
    for t in range(n3):
        C[:, :, t] = np.reshape(C3[t, :], (r, r), order="F")
    return C


# --------------------------------------------------------------------------
# factor updates (triple_decomp_ADMM.m:73-95)
# --------------------------------------------------------------------------
def update_A(X, A, B, C, alphaA):
    X1 = unfold(X, 1)
    F = buildF(B, C)
    G = F @ F.T + alphaA * np.eye(F.shape[0])
    A1 = (X1 @ F.T) @ pinv_matlab(G)
    return reshape_A_from_A1(A1, A.shape[0], A.shape[1])


def update_B(X, A, B, C, alphaB):
    X2 = unfold(X, 2)
    G = buildG(A, C)
    B_old_unf = (X2 @ G.T) @ pinv_matlab(G @ G.T + alphaB * np.eye(G.shape[0]))
    return reshape_B_from_B2(B_old_unf, B.shape[1], B.shape[0])


def update_C(X, A, B, C, alphaC=1e-9):
    X3 = unfold(X, 3)
    H = buildH(A, B)
    C_old_unf = (X3 @ H.T) @ pinv_matlab(H @ H.T + alphaC * np.eye(H.shape[0]))
    return reshape_C_from_C3(C_old_unf, C.shape[2], C.shape[0])


# --------------------------------------------------------------------------
# the solver (triple_decomp_ADMM.m:1-70)
# --------------------------------------------------------------------------
REQUIRED_OPTS = ("mu", "rho", "lambda", "lambda2", "maxIter", "tol", "disp")


def triple_decomp_ADMM(D, r, opts, A0=None, B0=None, C0=None, rng=None,
                       return_state=False, on_iter=None):
    """[A,B,C,O,errHist] = triple_decomp_ADMM(D, r, opts).

    ``opts`` is a dict with the reference's field names (mu, rho, lambda,
    lambda2, maxIter, tol, disp); a missing field raises KeyError like MATLAB's
    "Unrecognized field name".  A0/B0/C0 inject the initial factors that the
    reference draws with randn at :23 (order A, B, C); if absent they are
    drawn from ``rng`` in that order.
    """
    for k in REQUIRED_OPTS:
        if k not in opts:
            raise KeyError(f"Unrecognized field name \"{k}\".")
    D = np.asfortranarray(D, dtype=np.float64)
    n1, n2, n3 = D.shape
    muL = opts["mu"]; rhoL = opts["rho"]; muL_max = opts["mu"] * 1e6
    muO = opts["mu"]; rhoO = opts["rho"]; muO_max = opts["mu"] * 1e6
    lam = opts["lambda"]
    lambda2 = opts["lambda2"]
    maxIter = int(opts["maxIter"]); tol = opts["tol"]; disp = opts["disp"]

    if A0 is None:
        rng = rng or np.random.default_rng(0)
        A0 = rng.standard_normal((n1, r, r))
        B0 = rng.standard_normal((r, n2, r))
        C0 = rng.standard_normal((r, r, n3))
    A = np.array(A0, dtype=np.float64, order="F")
    B = np.array(B0, dtype=np.float64, order="F")
    C = np.array(C0, dtype=np.float64, order="F")
    O = np.zeros((n1, n2, n3), order="F"); E = O.copy()
    Y_L = np.zeros((n1, n2, n3), order="F")
    Y_O = np.zeros((n1, n2, n3), order="F")

    normD = np.linalg.norm(D.ravel(order="K"))
    errHist = np.zeros(maxIter)
    errL_hist = np.zeros(maxIter); errO_hist = np.zeros(maxIter)

    k = 0
    for k in range(1, maxIter + 1):
        # 1) update L (A,B,C) from T = D - O + Y_L/muL
        T = D - O + (1 / muL) * Y_L
        A = update_A(T, A, B, C, lambda2)
        B = update_B(T, A, B, C, lambda2)
        C = update_C(T, A, B, C)

        L = triple_product(A, B, C)

        # 2) O
        R1 = D - L + (1 / muL) * Y_L
        R2 = E - (1 / muO) * Y_O
        O = (muL * R1 + muO * R2) / (muL + muO)

        # 3) E
        R3 = O + (1 / muO) * Y_O
        E = np.sign(R3) * np.maximum(np.abs(R3) - lam / muO, 0)

        # 4) duals
        resL = D - L - O
        resO = O - E
        Y_L = Y_L + muL * resL
        Y_O = Y_O + muO * resO

        # 5) mu
        muL = min(muL * rhoL, muL_max)
        muO = min(muO * rhoO, muO_max)

        eL = np.linalg.norm(resL.ravel(order="K")) / normD
        eO = np.linalg.norm(resO.ravel(order="K")) / normD
        errL_hist[k - 1] = eL; errO_hist[k - 1] = eO
        errHist[k - 1] = eL + eO
        if disp and k % 10 == 0:
            print("Iter %d, errL=%.2e, errO=%.2e" % (k, eL, eO))
        if on_iter is not None:
            on_iter(k, A, B, C, O, E, Y_L, Y_O)
        if k > 1 and abs(errHist[k - 1] - errHist[k - 2]) < tol * errHist[k - 2]:
            break

    errHist = errHist[:k]
    if return_state:
        return A, B, C, O, errHist, dict(E=E, Y_L=Y_L, Y_O=Y_O, L=L, muL=muL, muO=muO,
                                         errL=errL_hist[:k], errO=errO_hist[:k])
    return A, B, C, O, errHist


def triple_ADMM_masked(Y, mask, r, opts, A0, B0, C0):
    """[A,B,C,O,E,Out] = triple_ADMM_masked(Y, mask, r, opts): the completion variant the reference's drivers name in
    a comment (traffic_triple_comparison.m:53, video_triple_comparison.m:52) but do not ship -- there is NO reference
    code for it, so this function IS the specification; it is an opt-in extension.

    mask True = observed.  On the observed entries every statement is the one of triple_decomp_ADMM.m:33-59.  An
    unobserved entry carries no data-fit constraint: O = E = Y_L = Y_O = 0 there, it does not enter ||D||, resL,
    resO, and the low-rank target is imputed with the current estimate, T = L (zero before the first iteration: the
    drivers' own zero-fill, traffic_triple_comparison.m:34-35).  With an all-true mask the statements reduce to
    triple_decomp_ADMM exactly."""
    for k in REQUIRED_OPTS:
        if k not in opts:
            raise KeyError(f"Unrecognized field name \"{k}\".")
    Y = np.asfortranarray(Y, dtype=np.float64)
    M = np.asfortranarray(mask).astype(bool)
    n1, n2, n3 = Y.shape
    muL = opts["mu"]; rhoL = opts["rho"]; muL_max = opts["mu"] * 1e6
    muO = opts["mu"]; rhoO = opts["rho"]; muO_max = opts["mu"] * 1e6
    lam = opts["lambda"]; lambda2 = opts["lambda2"]
    maxIter = int(opts["maxIter"]); tol = opts["tol"]; disp = opts["disp"]
    A = np.array(A0, dtype=np.float64, order="F"); B = np.array(B0, dtype=np.float64, order="F")
    C = np.array(C0, dtype=np.float64, order="F")
    D = np.where(M, Y, 0.0)
    O = np.zeros((n1, n2, n3), order="F"); E = O.copy(); Y_L = O.copy(); Y_O = O.copy()
    L = O.copy()                                       # imputation before the first iteration: zero fill
    normD = np.linalg.norm(D.ravel(order="K"))
    errHist = np.zeros(maxIter)
    k = 0
    for k in range(1, maxIter + 1):
        T = np.where(M, D - O + (1 / muL) * Y_L, L)
        A = update_A(T, A, B, C, lambda2)
        B = update_B(T, A, B, C, lambda2)
        C = update_C(T, A, B, C)
        L = triple_product(A, B, C)
        R1 = D - L + (1 / muL) * Y_L
        R2 = E - (1 / muO) * Y_O
        O = np.where(M, (muL * R1 + muO * R2) / (muL + muO), 0.0)
        R3 = O + (1 / muO) * Y_O
        E = np.where(M, np.sign(R3) * np.maximum(np.abs(R3) - lam / muO, 0), 0.0)
        resL = np.where(M, D - L - O, 0.0)
        resO = np.where(M, O - E, 0.0)
        Y_L = Y_L + muL * resL
        Y_O = Y_O + muO * resO
        muL = min(muL * rhoL, muL_max)
        muO = min(muO * rhoO, muO_max)
        eL = np.linalg.norm(resL.ravel(order="K")) / normD
        eO = np.linalg.norm(resO.ravel(order="K")) / normD
        errHist[k - 1] = eL + eO
        if disp and k % 10 == 0:
            print("Iter %d, errL=%.2e, errO=%.2e" % (k, eL, eO))
        if k > 1 and abs(errHist[k - 1] - errHist[k - 2]) < tol * errHist[k - 2]:
            break
    return A, B, C, O, E, dict(errHist=errHist[:k], L=L)


def triple_decomp_ALS(X, r, opts, A0=None, B0=None, C0=None, rng=None, disp=False):
    """[A,B,C,errHist] = triple_decomp_ALS(X, r, opts)  (triple_decomp_ALS.m:1-40)."""
    maxIter = int(opts["maxIter"]); tol = opts["tol"]
    X = np.asfortranarray(X, dtype=np.float64)
    n1, n2, n3 = X.shape
    Xnorm = np.linalg.norm(X.ravel(order="K"))
    if A0 is None:
        rng = rng or np.random.default_rng(0)
        A0 = rng.standard_normal((n1, r, r))
        B0 = rng.standard_normal((r, n2, r))
        C0 = rng.standard_normal((r, r, n3))
    A = np.array(A0, order="F"); B = np.array(B0, order="F"); C = np.array(C0, order="F")
    errHist = np.zeros(maxIter)
    k = 0
    for k in range(1, maxIter + 1):
        Xhat = triple_product(A, B, C)
        errHist[k - 1] = np.linalg.norm((X - Xhat).ravel(order="K")) / Xnorm
        if disp and k % 5 == 0:
            print("Iteration %d, relative error = %.4e" % (k, errHist[k - 1]))
        if k > 1 and abs(errHist[k - 1] - errHist[k - 2]) < tol * errHist[k - 2]:
            errHist = errHist[:k]
            break
        A = update_A(X, A, B, C, 1e-9)
        B = update_B(X, A, B, C, 1e-9)
        C = update_C(X, A, B, C, 1e-9)
    return A, B, C, errHist


# --------------------------------------------------------------------------
# definitional scalar loops (the only "known answers" the reference holds)
# --------------------------------------------------------------------------
def buildF_loops(B, C):
    """buildF.m:5-16 (commented definition): F(q+(s-1)r, j+(t-1)n2) = B(q,j,s)*C(q,s,t)."""
    r, n2, _ = B.shape
    n3 = C.shape[2]
    F = np.zeros((r * r, n2 * n3))
    for j in range(n2):
        for t in range(n3):
            col = j + t * n2
            for q in range(r):
                for s in range(r):
                    F[q + s * r, col] = B[q, j, s] * C[q, s, t]
    return F


def buildG_loops(A, C):
    """buildG.m:5-16: G(p+(s-1)r, i+(t-1)n1) = A(i,p,s)*C(p,s,t)."""
    n1, r, _ = A.shape
    n3 = C.shape[2]
    G = np.zeros((r * r, n1 * n3))
    for i in range(n1):
        for t in range(n3):
            col = i + t * n1
            for p in range(r):
                for s in range(r):
                    G[p + s * r, col] = A[i, p, s] * C[p, s, t]
    return G


def buildH_loops(A, B):
    """buildH.m:5-16: H(p+(q-1)r, i+(j-1)n1) = A(i,p,q)*B(p,j,q)."""
    n1, r, _ = A.shape
    n2 = B.shape[1]
    H = np.zeros((r * r, n1 * n2))
    for i in range(n1):
        for j in range(n2):
            col = i + j * n1
            for p in range(r):
                for q in range(r):
                    H[p + q * r, col] = A[i, p, q] * B[p, j, q]
    return H


def triple_product_loops(A, B, C):
    """origin_triple_tensor/triple_decomp_ADMM.m:125-143 (five nested loops)."""
    n1, n2, n3 = A.shape[0], B.shape[1], C.shape[2]
    X = np.zeros((n1, n2, n3), order="F")
    for i in range(n1):
        for j in range(n2):
            for t in range(n3):
                s = 0.0
                for p in range(A.shape[1]):
                    for q in range(A.shape[2]):
                        s = s + A[i, p, q] * B[p, j, q] * C[p, q, t]
                X[i, j, t] = s
    return X


# --------------------------------------------------------------------------
# the n x R "unfolded factor" layout the CUDA library keeps on device
# --------------------------------------------------------------------------
# --------------------------------------------------------------------------
# the ORIGINAL (Qi) triple decomposition's design matrices and product
# (origin_triple_tensor/buildF.m:4-6, buildG.m:9-11, buildH.m:9-11, triple_product.m)
# --------------------------------------------------------------------------
def buildF_qi(B, C):
    """origin_triple_tensor/buildF.m:4-6:  F(q+(s-1)r, j+(t-1)n2) = sum_p B(p,j,s) C(p,q,t)."""
    r, n2, _ = B.shape
    n3 = C.shape[2]
    F = np.reshape(np.transpose(B, (1, 2, 0)), (n2 * r, r), order="F") @ np.reshape(C, (r, r * n3), order="F")
    F = np.transpose(np.reshape(F, (n2, r, r, n3), order="F"), (2, 1, 0, 3))
    return np.reshape(F, (r * r, n2 * n3), order="F")


def buildG_qi(A, C):
    """origin_triple_tensor/buildG.m:9-11:  G(p+(s-1)r, i+(t-1)n1) = sum_q A(i,q,s) C(p,q,t)."""
    n1, r, _ = A.shape
    n3 = C.shape[2]
    G = np.reshape(np.transpose(A, (0, 2, 1)), (n1 * r, r), order="F") @ np.reshape(np.transpose(C, (1, 0, 2)), (r, r * n3), order="F")
    G = np.transpose(np.reshape(G, (n1, r, r, n3), order="F"), (2, 1, 0, 3))
    return np.reshape(G, (r * r, n1 * n3), order="F")


def buildH_qi(A, B):
    """origin_triple_tensor/buildH.m:9-11:  H(p+(q-1)r, i+(j-1)n1) = sum_s A(i,q,s) B(p,j,s)."""
    n1, r, _ = A.shape
    n2 = B.shape[1]
    H = np.reshape(A, (n1 * r, r), order="F") @ np.reshape(np.transpose(B, (2, 0, 1)), (r, r * n2), order="F")
    H = np.transpose(np.reshape(H, (n1, r, r, n2), order="F"), (2, 1, 0, 3))
    return np.reshape(H, (r * r, n1 * n2), order="F")


def triple_product_qi(A, B, C):
    """origin_triple_tensor/triple_product.m:  X(i,j,t) = sum_{p,q,s} A(i,q,s) B(p,j,s) C(p,q,t)."""
    return np.asfortranarray(np.einsum("iqs,pjs,pqt->ijt", A, B, C))


def design_qi_loops(which, U, V):
    """The commented scalar definitions (buildG.m:5-6, buildH.m:5-6 and the analogous one for F) as plain loops."""
    if which == 0:      # F: U = B, V = C
        r, na, _ = U.shape; nb = V.shape[2]
    elif which == 1:    # G: U = A, V = C
        na, r, _ = U.shape; nb = V.shape[2]
    else:               # H: U = A, V = B
        na, r, _ = U.shape; nb = V.shape[1]
    out = np.zeros((r * r, na * nb), order="F")
    for b in range(nb):
        for a in range(na):
            for k1 in range(r):
                for k0 in range(r):
                    s = 0.0
                    for m in range(r):
                        if which == 0:
                            s += U[m, a, k1] * V[m, k0, b]        # sum_p B(p,j,s) C(p,q,t), row q + r s
                        elif which == 1:
                            s += U[a, m, k1] * V[k0, m, b]        # sum_q A(i,q,s) C(p,q,t), row p + r s
                        else:
                            s += U[a, k1, m] * V[k0, b, m]        # sum_s A(i,q,s) B(p,j,s), row p + r q
                    out[k0 + r * k1, a + na * b] = s
    return out


def factors_to_unfolded(A, B, C):
    """A1 (n1 x R) = unfold(A,1); B2 (n2 x R) = unfold(B,2); C3 (n3 x R) = unfold(C,3)."""
    return unfold(A, 1), unfold(B, 2), unfold(C, 3)


def mttkrp(T, A1, B2, C3, mode):
    """RHS of the mode-k ridge solve, X_(k) * M^T, via the CP-R form."""
    if mode == 1:
        return np.einsum("ijt,jk,tk->ik", T, B2, C3, optimize=True)
    if mode == 2:
        return np.einsum("ijt,ik,tk->jk", T, A1, C3, optimize=True)
    if mode == 3:
        return np.einsum("ijt,ik,jk->tk", T, A1, B2, optimize=True)
    raise ValueError("Mode must be 1, 2, or 3.")
