"""Mode-3 sharded restatement of the TriTD-ADMM iteration -- TEST INFRASTRUCTURE, NOT PRODUCT.

Same algorithm as oracle/tritd_oracle.py (reference:
fast_robust_triple_tensor/triple_decomp_ADMM.m:15-68, :73-95), rewritten in the
form the CUDA library uses so the host-side sharding logic can be checked on
CPU (gloo, world_size 2) against the plain oracle:

  * factors kept as n x R matrices A1, B2, C3 (R = r^2);
  * RHS of the ridge solves as MTTKRPs, Gram matrices through the Hadamard
    identity  F F' = (B2'B2) o (C3'C3)  etc.  (SURVEY.md fact 1);
  * each rank holds a slab t in [t0,t1) of D, O, E, Y_L, Y_O and the matching
    rows of C3; partial [RHS_A ; C3'C3], RHS_B and the two residual norms are
    summed with ``allreduce`` (identity on one rank); the C update is slab-local.

PARITY UNPINNED (see tritd_oracle.py): the reference has no fixtures for this path.
"""
from __future__ import annotations

import numpy as np

from tritd_oracle import pinv_matlab


def _ridge_solve(rhs, G):
    return rhs @ pinv_matlab(G)


def admm_sharded(D, r, opts, A1, B2, C3, allreduce=lambda x: x, iters=None):
    """Runs the iteration on one slab.  D: n1 x n2 x n3_local (column-major),
    A1 (n1 x R), B2 (n2 x R) replicated, C3 (n3_local x R) local rows.
    Returns A1, B2, C3, O, errHist (errHist identical on every rank)."""
    R = r * r
    D = np.asfortranarray(D, dtype=np.float64)
    A1 = np.array(A1, dtype=np.float64); B2 = np.array(B2, dtype=np.float64); C3 = np.array(C3, dtype=np.float64)
    muL = muO = opts["mu"]
    mu_max = opts["mu"] * 1e6
    rho = opts["rho"]; lam = opts["lambda"]; lambda2 = opts["lambda2"]
    maxIter = int(opts["maxIter"]) if iters is None else iters
    tol = opts["tol"]
    O = np.zeros_like(D); E = np.zeros_like(D); Y_L = np.zeros_like(D); Y_O = np.zeros_like(D)
    normD = np.sqrt(allreduce(np.array([np.sum(D * D)]))[0])
    errHist = []
    I = np.eye(R)
    for k in range(1, maxIter + 1):
        T = D - O + (1 / muL) * Y_L
        # A: partial MTTKRP over the local slices + partial C3'C3, one all-reduce
        buf = np.concatenate([np.einsum("ijt,jk,tk->ik", T, B2, C3, optimize=True).ravel(), (C3.T @ C3).ravel()])
        buf = allreduce(buf)
        rhsA = buf[:A1.size].reshape(A1.shape); SC = buf[A1.size:].reshape(R, R)
        A1 = _ridge_solve(rhsA, (B2.T @ B2) * SC + lambda2 * I)
        # B with the new A: P is shared with the C update
        P = np.einsum("ijt,ik->tjk", T, A1, optimize=True)                 # P[t][j][k]
        rhsB = allreduce(np.einsum("tk,tjk->jk", C3, P))
        SA = A1.T @ A1
        B2 = _ridge_solve(rhsB, SA * SC + lambda2 * I)
        # C: slab-local
        rhsC = np.einsum("jk,tjk->tk", B2, P)
        C3 = _ridge_solve(rhsC, SA * (B2.T @ B2) + 1e-9 * I)
        L = np.einsum("ik,jk,tk->ijt", A1, B2, C3, optimize=True)
        R1 = D - L + (1 / muL) * Y_L
        R2 = E - (1 / muO) * Y_O
        O = (muL * R1 + muO * R2) / (muL + muO)
        R3 = O + (1 / muO) * Y_O
        E = np.sign(R3) * np.maximum(np.abs(R3) - lam / muO, 0)
        resL = D - L - O
        resO = O - E
        Y_L = Y_L + muL * resL
        Y_O = Y_O + muO * resO
        muL = min(muL * rho, mu_max); muO = min(muO * rho, mu_max)
        n2s = allreduce(np.array([np.sum(resL * resL), np.sum(resO * resO)]))
        errHist.append(np.sqrt(n2s[0]) / normD + np.sqrt(n2s[1]) / normD)
        if k > 1 and abs(errHist[-1] - errHist[-2]) < tol * errHist[-2]:
            break
    return A1, B2, C3, O, np.array(errHist)
