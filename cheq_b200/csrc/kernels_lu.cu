// Fallback: partial-pivot LU of the bordered system, the reference's own algorithm
// (`work_matrix.partial_piv_lu().solve(&rhs)`, /root/reference/src/solver/implementation.rs:574-576).
//
// The Cholesky path needs the hardness + Coulomb matrix to be positive definite.  That holds for physical
// geometries but not for every input the reference accepts (its own C2H4 fixture, tests/hydrogen.rs:175-207,
// has two hydrogens 0.21 Angstrom apart and an indefinite matrix).  When the Cholesky meets a non-positive pivot the
// engine re-solves with this blocked right-looking LU of [[A, 1], [1^T, 0]] with the right-hand side carried as
// an extra column.  Two-level blocking: 32-column panels (pivot search, row interchanges, eager updates) inside outer
// blocks of 128 columns; everything right of an outer block is updated once per block with K = 128 on the FP64 tensor
// cores (gemm_nt_kernel, DMMA) from a transposed copy of U12 -- the bulk of the (2/3) N^3 flops.
#include <cuda_runtime.h>

#include "device_types.h"
#include "kernels.h"

namespace cheq {
namespace {

constexpr int NB = 32;    // panel width
constexpr int OB = 128;   // outer block: delayed (tensor-core) update of everything right of it

// F (n1p x (n1p + 128), ld n1p) <- bordered matrix in internal order.  n1p = NP + 128; column n1p is the right-hand side,
// the 127 columns after it are zero padding (the tensor-core update works on whole 128-column tiles).
//   F[i, j] = A[max, min] (i, j < NP);  F[i, NP] = F[NP, i] = one_i;  F[NP, NP] = 0;  identity padding beyond;
//   hydrogen diagonal: hardness * (1 + clamp(q) * factor) (implementation.rs:567-572);
//   column n1p = right-hand side: c_i (row NP of the trapezoid's rhs rows), total charge at row NP.
__global__ void lu_build_kernel(const double* A, long ld, int NP, int n1p, double* F, const double* q,
                                const double* hard, const uint8_t* type, int nxp, int scf, double total_charge) {
    const int i = blockIdx.y * blockDim.x + threadIdx.x;
    const int j = blockIdx.x;
    if (i >= n1p) return;
    double v = 0.0;
    if (j < NP) {
        if (i < NP) {
            v = (i >= j) ? A[(long)i + (long)j * ld] : A[(long)j + (long)i * ld];
            if (i == j && scf && i >= nxp && type[i] != kPadType) {
                const double qc = fmin(fmax(q[i], -0.95), 0.95);
                v = hard[i] * (1.0 + qc * kHChargeFactor);
            }
        } else if (i == NP) {
            v = (type[j] != kPadType) ? 1.0 : 0.0;
        }
    } else if (j == NP) {
        if (i < NP) v = (type[i] != kPadType) ? 1.0 : 0.0;
    } else if (j < n1p) {
        v = (i == j) ? 1.0 : 0.0;
    } else if (j == n1p) {   // right-hand side
        if (i < NP) v = A[(long)NP + (long)i * ld];
        else if (i == NP) v = total_charge;
    }
    F[(long)i + (long)j * n1p] = v;
}

// Panel: columns [k0, k0 + NB), rows [k0, n1p).  One CTA.
__global__ void __launch_bounds__(1024) lu_panel_kernel(double* F, int n1p, int k0, int* ipiv, int* info) {
    __shared__ double red_v[32];
    __shared__ int red_i[32];
    __shared__ int piv_row;
    __shared__ double piv_val;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const long ld = n1p;
    for (int j = 0; j < NB; ++j) {
        const int col = k0 + j;
        double* Fc = F + (long)col * ld;
        // pivot search: first maximum of |F[r, col]|, r >= col
        double best = -1.0;
        int bi = col;
        for (int r = col + t; r < n1p; r += 1024) {
            const double a = fabs(Fc[r]);
            if (a > best) {
                best = a;
                bi = r;
            }
        }
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (ov > best || (ov == best && oi < bi)) {
                best = ov;
                bi = oi;
            }
        }
        if (lane == 0) {
            red_v[warp] = best;
            red_i[warp] = bi;
        }
        __syncthreads();
        if (t == 0) {
            double b = red_v[0];
            int ix = red_i[0];
            for (int w = 1; w < 32; ++w)
                if (red_v[w] > b || (red_v[w] == b && red_i[w] < ix)) {
                    b = red_v[w];
                    ix = red_i[w];
                }
            piv_row = ix;
            ipiv[col] = ix;
            if (!(b > 0.0) || !(b < 1.7e308)) atomicExch(info, col + 1);
        }
        __syncthreads();
        const int p = piv_row;
        if (p != col && t < NB) {
            double* a = F + (long)(k0 + t) * ld;
            const double tmp = a[col];
            a[col] = a[p];
            a[p] = tmp;
        }
        __syncthreads();
        if (t == 0) piv_val = Fc[col];
        __syncthreads();
        const double pv = piv_val;
        const double inv = (pv != 0.0) ? 1.0 / pv : 0.0;
        for (int r = col + 1 + t; r < n1p; r += 1024) Fc[r] *= inv;
        __syncthreads();
        // rank-1 update of the remaining panel columns
        for (int c = j + 1; c < NB; ++c) {
            double* Fd = F + (long)(k0 + c) * ld;
            const double u = Fd[col];
            for (int r = col + 1 + t; r < n1p; r += 1024) Fd[r] -= Fc[r] * u;
        }
        __syncthreads();
    }
}

// Row interchanges of one panel applied to the columns [c_begin, c_end): those right of it (including the rhs column)
// and the columns of the current outer block left of it (their L21 parts feed the delayed update in final row order).
__global__ void lu_swap_kernel(double* F, int n1p, int c_begin, int c_end, int k0, const int* ipiv) {
    const int c = c_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c_end) return;
    double* a = F + (long)c * n1p;
    for (int j = 0; j < NB; ++j) {
        const int r = k0 + j, p = ipiv[r];
        if (p != r) {
            const double tmp = a[r];
            a[r] = a[p];
            a[p] = tmp;
        }
    }
}

// U12 = L11^-1 A12 (unit lower triangular 32 x 32 at (k0, k0)) for the columns [c_begin, c_end), one thread per column.
__global__ void __launch_bounds__(128) lu_trsm_kernel(double* F, int n1p, int c_begin, int c_end, int k0) {
    __shared__ double L[NB][NB + 1];
    for (int idx = threadIdx.x; idx < NB * NB; idx += blockDim.x) {
        const int i = idx % NB, j = idx / NB;
        L[i][j] = F[(long)(k0 + i) + (long)(k0 + j) * n1p];
    }
    __syncthreads();
    const int c = c_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c_end) return;
    double* a = F + (long)c * n1p + k0;
    double x[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) x[i] = a[i];
#pragma unroll
    for (int j = 0; j < NB; ++j)
#pragma unroll
        for (int i = j + 1; i < NB; ++i) x[i] -= L[i][j] * x[j];
#pragma unroll
    for (int i = 0; i < NB; ++i) a[i] = x[i];
}

// A[r, c] -= L[r, k0 : k0 + 32] U[k0 : k0 + 32, c] for rows [r_begin, r_end) and columns [c_begin, c_end): 64 x 64
// tiles, K = 32, 4 x 4 outputs per thread.
__global__ void __launch_bounds__(256) lu_gemm_kernel(double* F, int n1p, int r_begin, int r_end, int c_begin, int c_end,
                                                      int k0) {
    __shared__ double Ls[NB][64];
    __shared__ double Us[NB][64 + 1];
    const int r0 = r_begin + blockIdx.x * 64, c0 = c_begin + blockIdx.y * 64;
    const int ncols = c_end;
    const int t = threadIdx.x;
    const long ld = n1p;
    for (int idx = t; idx < NB * 64; idx += 256) {
        const int r = idx & 63, j = idx >> 6;
        Ls[j][r] = (r0 + r < r_end) ? F[(long)(r0 + r) + (long)(k0 + j) * ld] : 0.0;
    }
    for (int idx = t; idx < NB * 64; idx += 256) {
        const int j = idx & (NB - 1), c = idx / NB;
        Us[j][c] = (c0 + c < ncols) ? F[(long)(k0 + j) + (long)(c0 + c) * ld] : 0.0;
    }
    __syncthreads();
    const int tr = (t & 15) * 4, tc = (t >> 4) * 4;
    double acc[4][4] = {};
#pragma unroll 8
    for (int j = 0; j < NB; ++j) {
        double l[4], u[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            l[a] = Ls[j][tr + a];
            u[a] = Us[j][tc + a];
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fma(l[a], u[b], acc[a][b]);
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int c = c0 + tc + b;
        if (c >= ncols) continue;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int r = r0 + tr + a;
            if (r < r_end) F[(long)r + (long)c * ld] -= acc[a][b];
        }
    }
}

// Back substitution on the rhs column y = F[:, n1p]: diagonal block kb (32 x 32 upper), one warp.
__global__ void lu_back_diag_kernel(double* F, int n1p, int kb) {
    __shared__ double U[NB][NB + 1];
    __shared__ double y[NB];
    const int lane = threadIdx.x;
    const int k0 = kb * NB;
    for (int j = 0; j < NB; ++j) U[lane][j] = F[(long)(k0 + lane) + (long)(k0 + j) * n1p];
    double* rhs = F + (long)n1p * n1p + k0;
    y[lane] = rhs[lane];
    __syncwarp();
    for (int j = NB - 1; j >= 0; --j) {
        if (lane == j) y[j] = y[j] / U[j][j];
        __syncwarp();
        if (lane < j) y[lane] -= U[lane][j] * y[j];
        __syncwarp();
    }
    rhs[lane] = y[lane];
}

// y[i] -= sum_j U[i, kb * 32 + j] x_j for the rows above block kb.
__global__ void lu_back_update_kernel(double* F, int n1p, int kb) {
    __shared__ double x[NB];
    const int k0 = kb * NB;
    double* rhs = F + (long)n1p * n1p;
    if (threadIdx.x < NB) x[threadIdx.x] = rhs[k0 + threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k0) return;
    double s = 0.0;
#pragma unroll 8
    for (int j = 0; j < NB; ++j) s = fma(F[(long)i + (long)(k0 + j) * n1p], x[j], s);
    rhs[i] -= s;
}

// Ut[c + k * right] = F[K0 + k, K1 + c]: the transposed copy of U12 (rows K0..K1 of the columns right of the outer
// block) that gemm_nt_kernel takes as its B operand (n x K, column-major).
__global__ void lu_transpose_kernel(const double* F, int n1p, int K0, int K1, int right, double* Ut) {
    __shared__ double tile[32][33];
    const int c0 = blockIdx.x * 32, kq = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int c = c0 + j, k = kq + threadIdx.x;          // coalesced along k (rows of F)
        tile[j][threadIdx.x] = (c < right) ? F[(long)(K0 + k) + (long)(K1 + c) * n1p] : 0.0;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int k = kq + j, c = c0 + threadIdx.x;          // coalesced along c
        if (c < right) Ut[(long)c + (long)k * right] = tile[threadIdx.x][j];
    }
}

// x (internal order, NP) and mu from the solved rhs column.
__global__ void lu_extract_kernel(const double* F, int n1p, int NP, double* v, ScfState* state) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double* rhs = F + (long)n1p * n1p;
    if (i < NP) v[i] = rhs[i];
    if (i == NP) state[0].mu = rhs[NP];
}

}  // namespace

int lu_padded_size(int NP) { return NP + kTile; }
size_t lu_matrix_bytes(int NP) { return (size_t)(NP + kTile) * (NP + 2 * kTile) * 8; }
size_t lu_workspace_bytes(int NP) { return (size_t)(NP + 2 * kTile) * OB * 8; }

cudaError_t launch_lu_solve(const double* A, long ld, int NP, int nxp, int scf, double* F, double* Ut, const double* q,
                            const double* hard, const uint8_t* type, double total_charge, int* ipiv, double* v,
                            ScfState* state, unsigned long long* launches, cudaStream_t stream) {
    const int n1p = lu_padded_size(NP);
    const int ncols = n1p + 1;          // matrix + right-hand side
    const int ncolsP = n1p + kTile;     // ... padded to whole tiles
    {
        dim3 grid(ncolsP, (n1p + 255) / 256);
        lu_build_kernel<<<grid, 256, 0, stream>>>(A, ld, NP, n1p, F, q, hard, type, nxp, scf, total_charge);
        ++*launches;
    }
    for (int K0 = 0; K0 < n1p; K0 += OB) {
        const int K1 = K0 + OB;         // n1p is a multiple of 128: whole outer blocks
        for (int k0 = K0; k0 < K1; k0 += NB) {
            lu_panel_kernel<<<1, 1024, 0, stream>>>(F, n1p, k0, ipiv, &state[0].info);
            ++*launches;
            // interchanges: everything right of the panel, and the outer block's columns left of it
            lu_swap_kernel<<<(ncols - (k0 + NB) + 127) / 128, 128, 0, stream>>>(F, n1p, k0 + NB, ncols, k0, ipiv);
            ++*launches;
            if (k0 > K0) {
                lu_swap_kernel<<<(k0 - K0 + 127) / 128, 128, 0, stream>>>(F, n1p, K0, k0, k0, ipiv);
                ++*launches;
            }
            // eager update of the rest of the outer block
            const int inner = K1 - (k0 + NB);
            if (inner > 0) {
                lu_trsm_kernel<<<(inner + 127) / 128, 128, 0, stream>>>(F, n1p, k0 + NB, K1, k0);
                dim3 grid((n1p - (k0 + NB) + 63) / 64, (inner + 63) / 64);
                lu_gemm_kernel<<<grid, 256, 0, stream>>>(F, n1p, k0 + NB, n1p, k0 + NB, K1, k0);
                *launches += 2;
            }
        }
        // delayed update of everything right of the outer block: U12 = L11^-1 A12 by 32-row blocks, then
        // A22 -= L21 U12 with K = 128 on the tensor cores (B operand = U12^T, copied out)
        const int right = ncolsP - K1;
        if (right <= 0) continue;
        for (int jb = 0; jb < OB / NB; ++jb) {
            const int kk = K0 + jb * NB;
            lu_trsm_kernel<<<(ncols - K1 + 127) / 128, 128, 0, stream>>>(F, n1p, K1, ncols, kk);
            ++*launches;
            if (jb + 1 < OB / NB) {
                dim3 grid((K1 - (kk + NB) + 63) / 64, (ncols - K1 + 63) / 64);
                lu_gemm_kernel<<<grid, 256, 0, stream>>>(F, n1p, kk + NB, K1, K1, ncols, kk);
                ++*launches;
            }
        }
        const int below = n1p - K1;
        if (below <= 0) continue;
        {
            dim3 grid((right + 31) / 32, OB / 32);
            lu_transpose_kernel<<<grid, dim3(32, 8), 0, stream>>>(F, n1p, K0, K1, right, Ut);
            ++*launches;
        }
        GemmArgs g{};
        g.A = F + (long)K1 + (long)K0 * n1p;
        g.lda = n1p;
        g.B = Ut;
        g.ldb = right;
        g.C = F + (long)K1 + (long)K1 * n1p;
        g.ldc = n1p;
        g.K = OB;
        g.subtract = 1;
        cudaError_t e = launch_gemm_nt(g, below / kTile, right / kTile, 1, stream);
        if (e != cudaSuccess) return e;
        ++*launches;
    }
    for (int kb = n1p / NB - 1; kb >= 0; --kb) {
        lu_back_diag_kernel<<<1, NB, 0, stream>>>(F, n1p, kb);
        ++*launches;
        if (kb > 0) {
            lu_back_update_kernel<<<(kb * NB + 255) / 256, 256, 0, stream>>>(F, n1p, kb);
            ++*launches;
        }
    }
    lu_extract_kernel<<<(NP + 1 + 255) / 256, 256, 0, stream>>>(F, n1p, NP, v, state);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace cheq
