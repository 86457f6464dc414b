"""NumPy model of the schedule of the pivoted-LU fallback (cheq_b200/csrc/kernels_lu.cu; test infrastructure, CPU only):
32-column panels with partial pivoting inside 128-column outer blocks, row interchanges applied to the columns right of
a panel AND to the outer block's columns left of it, eager updates inside the outer block, one delayed update of
everything right of the block (on the device: U12 by four 32-row triangular solves, then A22 -= L21 U12 with K = 128 on
the tensor cores).  Restates launch_lu_solve step for step so that the schedule itself is pinned without a GPU."""
import numpy as np

NB, OB = 32, 128


def solve(A, b):
    n1p = A.shape[0]
    assert n1p % OB == 0
    ncols = n1p + 1
    F = np.zeros((n1p, ncols))
    F[:, :n1p] = A
    F[:, n1p] = b
    ipiv = np.zeros(n1p, dtype=int)
    for K0 in range(0, n1p, OB):
        K1 = K0 + OB
        for k0 in range(K0, K1, NB):
            for j in range(NB):                                   # lu_panel_kernel
                col = k0 + j
                p = col + int(np.argmax(np.abs(F[col:n1p, col])))
                ipiv[col] = p
                if p != col:
                    F[[col, p], k0:k0 + NB] = F[[p, col], k0:k0 + NB]
                F[col + 1:n1p, col] /= F[col, col]
                for c in range(j + 1, NB):
                    F[col + 1:n1p, k0 + c] -= F[col + 1:n1p, col] * F[col, k0 + c]
            for cb, ce in ((k0 + NB, ncols), (K0, k0)):           # lu_swap_kernel, twice
                for j in range(NB):
                    r, p = k0 + j, ipiv[k0 + j]
                    if p != r:
                        F[[r, p], cb:ce] = F[[p, r], cb:ce]
            if K1 > k0 + NB:                                      # eager: lu_trsm_kernel + lu_gemm_kernel
                L11 = np.tril(F[k0:k0 + NB, k0:k0 + NB], -1) + np.eye(NB)
                F[k0:k0 + NB, k0 + NB:K1] = np.linalg.solve(L11, F[k0:k0 + NB, k0 + NB:K1])
                F[k0 + NB:n1p, k0 + NB:K1] -= F[k0 + NB:n1p, k0:k0 + NB] @ F[k0:k0 + NB, k0 + NB:K1]
        for jb in range(OB // NB):                                # delayed: U12 by 32-row blocks
            kk = K0 + jb * NB
            L11 = np.tril(F[kk:kk + NB, kk:kk + NB], -1) + np.eye(NB)
            F[kk:kk + NB, K1:ncols] = np.linalg.solve(L11, F[kk:kk + NB, K1:ncols])
            if jb + 1 < OB // NB:
                F[kk + NB:K1, K1:ncols] -= F[kk + NB:K1, kk:kk + NB] @ F[kk:kk + NB, K1:ncols]
        if n1p > K1:                                              # transposed copy + gemm_nt_kernel (K = 128)
            F[K1:n1p, K1:ncols] -= F[K1:n1p, K0:K1] @ F[K0:K1, K1:ncols]
    return np.linalg.solve(np.triu(F[:n1p, :n1p]), F[:n1p, n1p])   # lu_back_* kernels
