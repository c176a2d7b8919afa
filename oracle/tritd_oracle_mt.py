"""Multi-threaded CPU port of the TriTD-ADMM iteration -- TEST INFRASTRUCTURE / TIMED CPU BASELINE, NOT PRODUCT.

Same statements, same passes over the data as oracle/tritd_oracle.py (and therefore as
fast_robust_triple_tensor/triple_decomp_ADMM.m:15-68, :73-95, :97-160; triple_product.m:6-7), but written with
torch CPU tensors so that every N-sized pass -- the element-wise block (:33, :41-53), the permuted ``unfold`` copies
(:97-109), the materialised design matrices (:132-160) and the dgemm calls (:78, :86, :93, triple_product.m:6) -- uses
all host threads, the way MATLAB's implicitly multi-threaded built-ins do.  numpy's element-wise passes are
single-threaded, which makes the plain numpy oracle an unfairly slow *timing* baseline; this port exists only for
bench.py's ``cpu_baseline`` / ``--impl reference`` legs.  It is checked against the numpy oracle in
tests/test_oracle.py (same errHist and factors to 1e-10); parity of the product is always judged against
oracle/tritd_oracle.py.

Storage: an n1 x n2 x n3 column-major array is a contiguous torch tensor of shape (n3, n2, n1), so
``X.reshape(n3*n2, n1)`` is unfold(X,1) transposed without a copy and the two other unfoldings are real permuted
copies, exactly the passes the reference makes.  The small r^2 x r^2 pinv is the numpy oracle's.
"""
from __future__ import annotations

import os

import numpy as np
import torch

import tritd_oracle as orc


def _rev(X):
    """numpy column-major (n1,n2,n3) -> torch (n3,n2,n1) contiguous sharing the same memory order."""
    return torch.from_numpy(np.ascontiguousarray(np.transpose(np.asfortranarray(X, dtype=np.float64), (2, 1, 0))))


def _unrev(T):
    return np.asfortranarray(np.transpose(T.numpy(), (2, 1, 0)))


def _pinv(G):
    return torch.from_numpy(orc.pinv_matlab(G.numpy()))


def triple_decomp_ADMM(D, r, opts, A0, B0, C0, on_iter=None, threads=None):
    """[A,B,C,O,errHist] = triple_decomp_ADMM(D, r, opts) with injected initial factors (multi-threaded port)."""
    for k in orc.REQUIRED_OPTS:
        if k not in opts:
            raise KeyError(f"Unrecognized field name \"{k}\".")
    torch.set_num_threads(int(threads or os.cpu_count() or 1))
    n1, n2, n3 = D.shape
    R = r * r
    muL = opts["mu"]; rhoL = opts["rho"]; muL_max = opts["mu"] * 1e6
    muO = opts["mu"]; rhoO = opts["rho"]; muO_max = opts["mu"] * 1e6
    lam = opts["lambda"]; lambda2 = opts["lambda2"]
    maxIter = int(opts["maxIter"]); tol = opts["tol"]
    Dt = _rev(D)
    # unfolded factors (reshape_*_from_* are relabellings): A1[i, p + r s] = A(i,p,s), B2[j, p + r s] = B(p,j,s),
    # C3[t, p + r s] = C(p,s,t)
    A1 = torch.from_numpy(np.reshape(np.asfortranarray(A0, dtype=np.float64), (n1, R), order="F").copy())
    B2 = torch.from_numpy(np.reshape(np.transpose(np.asarray(B0, dtype=np.float64), (1, 0, 2)), (n2, R), order="F").copy())
    C3 = torch.from_numpy(np.reshape(np.asfortranarray(C0, dtype=np.float64), (R, n3), order="F").T.copy())
    O = torch.zeros_like(Dt); E = torch.zeros_like(Dt); Y_L = torch.zeros_like(Dt); Y_O = torch.zeros_like(Dt)
    normD = torch.linalg.vector_norm(Dt).item()
    errHist = np.zeros(maxIter)
    I = torch.eye(R, dtype=torch.float64)
    k = 0
    for k in range(1, maxIter + 1):
        T = Dt - O + (1 / muL) * Y_L                                                    # :33
        # update_A (:73-81): X1 = unfold(T,1) (a view), F = buildF(B,C) materialised, dgemm, pinv
        FT = (C3[:, None, :] * B2[None, :, :]).reshape(n3 * n2, R)                      # F' : (n2 n3) x R, row (t, j)
        A1 = (T.reshape(n3 * n2, n1).T @ FT) @ _pinv(FT.T @ FT + lambda2 * I)
        # update_B (:83-88): X2 = unfold(T,2) is a permuted copy, G = buildG(A,C)
        X2T = T.permute(0, 2, 1).contiguous().reshape(n3 * n1, n2)                      # X2' : (n1 n3) x n2, row (t, i)
        GT = (C3[:, None, :] * A1[None, :, :]).reshape(n3 * n1, R)
        B2 = (X2T.T @ GT) @ _pinv(GT.T @ GT + lambda2 * I)
        # update_C (:90-95): X3 = unfold(T,3) is a permuted copy, H = buildH(A,B), ridge 1e-9
        X3T = T.permute(1, 2, 0).contiguous().reshape(n2 * n1, n3)                      # X3' : (n1 n2) x n3, row (j, i)
        HT = (B2[:, None, :] * A1[None, :, :]).reshape(n2 * n1, R)
        C3 = (X3T.T @ HT) @ _pinv(HT.T @ HT + 1e-9 * I)
        # L = triple_product(A,B,C) (:38): rebuilds F, dgemm, reshape
        FT = (C3[:, None, :] * B2[None, :, :]).reshape(n3 * n2, R)
        L = (FT @ A1.T).reshape(n3, n2, n1)
        R1 = Dt - L + (1 / muL) * Y_L                                                   # :41
        R2 = E - (1 / muO) * Y_O                                                        # :42
        O = (muL * R1 + muO * R2) / (muL + muO)                                         # :43
        R3 = O + (1 / muO) * Y_O                                                        # :46
        E = torch.sign(R3) * torch.clamp(torch.abs(R3) - lam / muO, min=0)              # :47
        resL = Dt - L - O                                                               # :50
        resO = O - E                                                                    # :51
        Y_L = Y_L + muL * resL                                                          # :52
        Y_O = Y_O + muO * resO                                                          # :53
        muL = min(muL * rhoL, muL_max)                                                  # :56
        muO = min(muO * rhoO, muO_max)                                                  # :57
        eL = torch.linalg.vector_norm(resL).item() / normD                              # :59
        eO = torch.linalg.vector_norm(resO).item() / normD
        errHist[k - 1] = eL + eO
        if on_iter is not None:
            on_iter(k)
        if k > 1 and abs(errHist[k - 1] - errHist[k - 2]) < tol * errHist[k - 2]:       # :63-65
            break
    A = np.reshape(A1.numpy(), (n1, r, r), order="F")
    B = np.transpose(np.reshape(B2.numpy(), (n2, r, r), order="F"), (1, 0, 2))
    C = np.reshape(C3.numpy().T, (r, r, n3), order="F")
    return np.asfortranarray(A), np.asfortranarray(B), np.asfortranarray(C), _unrev(O), errHist[:k]
