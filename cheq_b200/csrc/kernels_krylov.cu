// Kernels of the stale-factor SCF iterations.
//
// The reference refactors the whole bordered matrix in every hydrogen-SCF iteration although only the diagonal of
// the hydrogen block changes (/root/reference/src/solver/implementation.rs:303-314, 567-576).  After eliminating the
// non-hydrogen block and the constraint border the iteration-k system is
//
//     K_k [x_h; mu] = [t_c; Q - h],   K_k = [[S0 + D_k, t_1], [t_1^T, -g]],   D_k = diag(J_h f clamp(q_h))
//
// with S0, t_c, t_1, g, h fixed per geometry.  Once the SCF has settled, the engine keeps the factorisation of ONE
// K_s ("frozen" at iteration s) as an explicit operator  S_s^-1 = R S R^T,  R = L^-T S  (upper triangular, built with
// the tensor-core GEMM), and solves every later K_k by preconditioned GMRES with iterative refinement: K_s^-1 K_k =
// I + K_s^-1 (D_k - D_s) has a tightly clustered spectrum (measured on S20k: 7-10 steps to 1e-15 for a factor that is
// 14 iterations old).  Each step streams the lower triangle of S0 once (tile_matvec_kernel, symmetric mode) and the
// upper triangle of R twice -- HBM-bound, no dependency chain -- instead of n_h^3 / 3 flops.
//
// All reductions have a fixed order (partials per tile, then one CTA per 128-row block), so results are
// bit-reproducible.
#include <cuda_runtime.h>

#include "device_types.h"
#include "kernels.h"

namespace cheq {
namespace {

// ---------------------------------------------------------------------------------------------------------------
// Tiled triangular matrix-vector products.  Grid (ceil(nt / 8), nt): CTA (bx, j) works on tile column j, warp w on
// tile (i, j) with  lower: i = j + 8 bx + w,  upper: i = 8 bx + w <= j.  A warp streams its 128 x 128 tile column by
// column (lane l owns rows l, l + 32, l + 64, l + 96: four coalesced 256-byte loads per column) and forms
//   N:  pN[r] = sum_c T[r, c] xN[c]        (per-lane accumulators, no communication)
//   T:  pT[c] = sum_r T[r, c] xT[r]        (one xor-shuffle reduction per column)
// Diagonal tiles are masked to the stored triangle; `sym` (symmetric matrix stored as its lower triangle) counts the
// diagonal itself only in the N product.
// ---------------------------------------------------------------------------------------------------------------
template <bool DO_N, bool DO_T, class T_ELEM>
__global__ void __launch_bounds__(256)
tile_matvec_kernel(TileMatvecArgs a) {
    const int j = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (a.col_mask && !a.col_mask[j]) return;
    const int i = a.upper ? (blockIdx.x * 8 + warp) : (j + blockIdx.x * 8 + warp);
    __shared__ double xs[kTile];
    if (DO_N) {
        if (threadIdx.x < kTile) {
            double v = a.xN[j * kTile + threadIdx.x];
            if (a.scaleN) v *= a.scaleN[j * kTile + threadIdx.x];
            xs[threadIdx.x] = v;
        }
        __syncthreads();
    }
    if (a.upper ? (i > j) : (i >= a.nt)) return;
    if (a.row_end > 0 && (i < a.row_begin || i >= a.row_end)) return;   // tile row held by another rank
    const bool diag = (i == j);
    // T_ELEM = float: the preconditioner's R is kept in single precision (half the bytes; GMRES only needs a FIXED
    // linear operator close to K_s^-1, the residuals that decide convergence are formed with the fp64 S0)
    // tile-major storage (col_base != nullptr): the stored tiles of tile column j are consecutive 128 x 128 blocks
    // starting at element col_base[j] (first stored tile row: j for a lower triangle, row_begin for an upper one), so a
    // warp streams ONE contiguous 64 / 128 KB block instead of 128 segments a whole matrix column apart -- at 100k
    // atoms the column stride is 534 KB and the column-major walk thrashes the TLB
    const long tile_ld = a.col_base ? kTile : a.ld;
    const T_ELEM* T = (const T_ELEM*)a.M +
                      (a.col_base ? a.col_base[j] + (long)(i - (a.upper ? a.row_begin : j)) * (kTile * kTile)
                                  : (long)i * kTile + (long)j * kTile * a.ld) + lane;
    double xt[4] = {0.0, 0.0, 0.0, 0.0};
    if (DO_T) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            double v = a.xT[i * kTile + lane + 32 * m];
            if (a.scaleT) v *= a.scaleT[i * kTile + lane + 32 * m];
            xt[m] = v;
        }
    }
    double accN[4] = {0.0, 0.0, 0.0, 0.0};
    double* outT = a.partT + ((long)j * a.nt + i) * kTile;
    double keepT = 0.0;
#pragma unroll 4
    for (int c = 0; c < kTile; ++c) {
        const T_ELEM* p = T + (long)c * tile_ld;
        double t[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) t[m] = (double)__ldg(p + 32 * m);   // default caching: a 3,000-atom system's S0 and R live in L2
        if (diag) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int r = lane + 32 * m;
                const bool in = a.upper ? (r <= c) : (r >= c);
                if (!in) t[m] = 0.0;
            }
        }
        if (DO_N) {
            const double xc = xs[c];
#pragma unroll
            for (int m = 0; m < 4; ++m) accN[m] = fma(t[m], xc, accN[m]);
        }
        if (DO_T) {
            if (diag && a.sym) {   // the diagonal entry itself belongs to the N product only
#pragma unroll
                for (int m = 0; m < 4; ++m)
                    if (lane + 32 * m == c) t[m] = 0.0;
            }
            double s = t[0] * xt[0];
            s = fma(t[1], xt[1], s);
            s = fma(t[2], xt[2], s);
            s = fma(t[3], xt[3], s);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if ((c & 31) == lane) keepT = s;
            if ((c & 31) == 31) outT[c - 31 + lane] = keepT;
        }
    }
    if (DO_N) {
        double* outN = a.partN + ((long)i * a.nt + j) * kTile;   // [output block][tile column]: contiguous for the reducer
#pragma unroll
        for (int m = 0; m < 4; ++m) outN[lane + 32 * m] = accN[m];
    }
}

// Rectangular product  out[col block j] = sum_i T(i, j)^T x[row block i]  over ALL tiles of an nt_rows x nt_cols tile
// matrix (W = L[hydrogen rows, non-hydrogen columns]: the hydrogen part of the back substitution's right-hand side,
// streamed at HBM speed instead of by the chain CTAs).  Same warp-per-tile scheme as tile_matvec_kernel, T mode;
// partials part[(j * nt_rows + i) * 128 + c], rows [row_begin, row_end) only (the ranks of a group split the rows).
__global__ void __launch_bounds__(256)
rect_matvec_t_kernel(const double* M, long ld, int nt_rows, int row_begin, int row_end, const double* x, double* part) {
    const int j = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = row_begin + blockIdx.x * 8 + warp;
    if (i >= row_end) return;
    const double* T = M + (long)i * kTile + (long)j * kTile * ld + lane;
    double xt[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) xt[m] = x[i * kTile + lane + 32 * m];
    double* out = part + ((long)j * nt_rows + i) * kTile;
    double keep = 0.0;
#pragma unroll 4
    for (int c = 0; c < kTile; ++c) {
        const double* p = T + (long)c * ld;
        double s = __ldg(p) * xt[0];
        s = fma(__ldg(p + 32), xt[1], s);
        s = fma(__ldg(p + 64), xt[2], s);
        s = fma(__ldg(p + 96), xt[3], s);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if ((c & 31) == lane) keep = s;
        if ((c & 31) == 31) out[c - 31 + lane] = keep;
    }
}

// out[j * 128 + r] = sum_{i in [row_begin, row_end)} part[(j * nt_rows + i) * 128 + r]   (fixed order, 4 groups)
__global__ void __launch_bounds__(512)
rect_matvec_reduce_kernel(const double* part, int nt_rows, int row_begin, int row_end, double* out) {
    const int j = blockIdx.x, r = threadIdx.x & 127, g = threadIdx.x >> 7;
    __shared__ double red[4][128];
    double s = 0.0;
    for (int i = row_begin + g; i < row_end; i += 4) s += part[((long)j * nt_rows + i) * kTile + r];
    red[g][r] = s;
    __syncthreads();
    if (g == 0) out[j * kTile + r] = (red[0][r] + red[1][r]) + (red[2][r] + red[3][r]);
}

// ---------------------------------------------------------------------------------------------------------------
// All-reduce of a short vector (n_h + 1 doubles) over NVLink / NVSwitch PEER MEMORY: three of these per GMRES step on
// a group, where ncclAllReduce costs ~55 us each (measured: 0.41 ms per step on 8 GPUs with 0.045 ms of streaming).
//   push:    every rank stores its partial vector into slot [rank] of the inbox of EVERY rank (its own included;
//            the inboxes live in the peer-mapped arenas), fences at system scope, and the last CTA to finish
//            release-stores the sequence number into that rank's flag word on every destination;
//   combine: waits until all `world` flags of the local inbox carry the sequence number, then sums the slots in rank
//            order -- the same operands in the same order on every rank, so the result is bit-identical everywhere.
// Two inboxes alternate by sequence parity: a rank can only push round s + 2 after its combine of round s + 1, which
// needed every peer's push of round s + 1, issued after that peer's combine of round s.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
p2p_push_kernel(const double* src, int len, P2PAllreduce pa, unsigned* done_counter) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < len) {
        const double v = src[i];
        for (int d = 0; d < pa.world; ++d) pa.inbox[d][(long)pa.rank * pa.slot + i] = v;
    }
    __threadfence_system();
    __shared__ int last;
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(done_counter, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (last) {
        __threadfence_system();
        if (threadIdx.x < pa.world) {
            volatile unsigned* f = pa.flags[threadIdx.x] + pa.rank;
            *f = pa.seq;
        }
        if (threadIdx.x == 0) *done_counter = 0u;   // ready for the next push on this stream
        __threadfence_system();
    }
}

__global__ void __launch_bounds__(256)
p2p_combine_kernel(P2PAllreduce pa, int len, double* out, int* error) {
    __shared__ int ok_s;
    if (threadIdx.x == 0) {
        volatile unsigned* f = pa.flags[pa.rank];
        int ok = 1;
        const long long t0 = clock64();
        for (int s = 0; s < pa.world; ++s) {
            // wrap-safe "flag >= seq"
            while ((int)(f[s] - pa.seq) < 0) {
                if (clock64() - t0 > 4000000000ll) {   // ~2 s: a peer never arrived -- report instead of hanging
                    ok = 0;
                    break;
                }
            }
            if (!ok) break;
        }
        if (!ok) atomicExch(error, 1);
        ok_s = ok;
    }
    __syncthreads();
    __threadfence_system();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= len) return;
    const double* in = pa.inbox[pa.rank];
    double s = 0.0;
    for (int r = 0; r < pa.world; ++r) s += __ldcg(in + (long)r * pa.slot + i);
    out[i] = ok_s ? s : nan("");
}

// One CTA per 128-row block b: fixed-order sum of the partials that belong to it, then the epilogue of the
// operator being applied.  Block nt (one extra CTA) computes the border entry.
//   op 0: out = sum                                   (triangular passes)
//   op 1: out = K v      = sum + d v + nu t1 ;  border: t1.v - g nu
//   op 2: out = rhs - K v                              (residual)
__global__ void __launch_bounds__(512)
tile_matvec_reduce_kernel(TileReduceArgs a) {
    // 128 rows x 4 groups: group g sums the partials with index = g (mod 4), the groups are combined in order
    const int b = blockIdx.x, r = threadIdx.x & 127, g = threadIdx.x >> 7;
    __shared__ double red[4][128];
    if (b == a.nt) {
        // border entry: t1 . v - g nu   (v has n = 128 nt entries, nu = v[n])
        if (a.op == 0) return;
        const int n = a.nt * kTile;
        double s = 0.0;
        for (int k = threadIdx.x; k < n; k += 512) s = fma(a.t1[k], a.v[k], s);
        red[g][r] = s;
        __syncthreads();
        if (g == 0) red[0][r] = (red[0][r] + red[1][r]) + (red[2][r] + red[3][r]);
        __syncthreads();
        for (int off = 64; off > 0; off >>= 1) {
            if (threadIdx.x < off) red[0][threadIdx.x] += red[0][threadIdx.x + off];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const double y = red[0][0] - a.sc->g * a.v[n];
            a.out[n] = a.border_owner ? ((a.op == 2) ? (a.rhs[n] - y) : y) : 0.0;
        }
        return;
    }
    const int rb = a.row_begin, re = (a.row_end > 0) ? a.row_end : a.nt;
    double s = 0.0;
    if (a.partN && b >= rb && b < re) {
        // N partials of tiles (b, j): lower j <= b, upper j >= b
        const int j0 = a.upper ? b : 0, j1 = a.upper ? a.nt - 1 : b;
        for (int j = j0 + g; j <= j1; j += 4) {
            if (a.col_mask && !a.col_mask[j]) continue;
            s += a.partN[((long)b * a.nt + j) * kTile + r];
        }
    }
    if (a.partT) {
        // T partials of tiles (i, b): lower i >= b, upper i <= b
        if (!a.col_mask || a.col_mask[b]) {
            const int i0 = max(a.upper ? 0 : b, rb), i1 = min(a.upper ? b : a.nt - 1, re - 1);
            for (int i = i0 + g; i <= i1; i += 4) s += a.partT[((long)b * a.nt + i) * kTile + r];
        }
    }
    red[g][r] = s;
    __syncthreads();
    if (g != 0) return;
    s = (red[0][r] + red[1][r]) + (red[2][r] + red[3][r]);
    const int k = b * kTile + r;
    if (a.op != 0) {
        const double nu = a.v[a.nt * kTile];
        if (!a.col_mask || a.col_mask[b]) s += a.d[k] * a.v[k] + nu * a.t1[k];   // owner adds the diagonal terms once
        if (a.op == 2) s = ((!a.col_mask || a.col_mask[b]) ? a.rhs[k] : 0.0) - s;
    }
    a.out[k] = s;
}

// ---------------------------------------------------------------------------------------------------------------
// small vector kernels (vectors have n + 1 entries: n hydrogen-block entries + the border unknown)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_256(double s, double* red) {
    const int t = threadIdx.x;
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((t & 31) == 0) red[t >> 5] = s;
    __syncthreads();
    double tot = 0.0;
    if (t == 0)
        for (int w = 0; w < 8; ++w) tot += red[w];
    return tot;   // valid in thread 0
}

// out[k] = V_k . w  for k < m (CTA k), out[m] = w . w (CTA m, only if with_norm)
__global__ void __launch_bounds__(256)
multi_dot_kernel(const double* V, long stride, const double* w, int len, double* out, int m) {
    __shared__ double red[8];
    const int k = blockIdx.x;
    const double* a = (k < m) ? V + (long)k * stride : w;
    double s = 0.0;
    for (int i = threadIdx.x; i < len; i += 256) s = fma(a[i], w[i], s);
    const double tot = block_sum_256(s, red);
    if (threadIdx.x == 0) out[k] = tot;
}

// w -= sum_k h[k] V_k ;  hsum[k] += h[k]
__global__ void __launch_bounds__(256)
gs_update_kernel(const double* V, long stride, double* w, int len, const double* h, double* hsum, int m, int first) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < len) {
        double s = w[i];
        for (int k = 0; k < m; ++k) s = fma(-h[k], V[(long)k * stride + i], s);
        w[i] = s;
    }
    if (blockIdx.x == 0)
        for (int k = threadIdx.x; k < m; k += 256) hsum[k] = first ? h[k] : hsum[k] + h[k];
}

// dst = src / sqrt(*norm2)
__global__ void __launch_bounds__(256)
scale_by_rnorm_kernel(const double* src, double* dst, int len, const double* norm2) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= len) return;
    const double nrm = sqrt(*norm2);
    dst[i] = (nrm > 0.0) ? src[i] / nrm : 0.0;
}

// x += sum_k y[k] V_k
__global__ void __launch_bounds__(256)
combine_kernel(const double* V, long stride, double* x, int len, const double* y, int m) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= len) return;
    double s = x[i];
    for (int k = 0; k < m; ++k) s = fma(y[k], V[(long)k * stride + i], s);
    x[i] = s;
}

// Border part of the frozen preconditioner K_s^-1 [r; rho]:  u0 = S_s^-1 r is given,
//   nu = (t1 . u0 - rho) / den,   out = [u0 - nu w_s; nu]          (one CTA: a dot product and an axpy over n)
__global__ void __launch_bounds__(1024)
precond_border_kernel(const double* u0, const double* rin, const double* t1, const double* ws, const KrylovScalars* sc,
                      double* out, int n) {
    __shared__ double red[32];
    __shared__ double nu_s;
    double s = 0.0;
    for (int k = threadIdx.x; k < n; k += 1024) s = fma(t1[k], u0[k], s);
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 32; ++w) tot += red[w];
        nu_s = (tot - rin[n]) / sc->den;
    }
    __syncthreads();
    const double nu = nu_s;
    for (int k = threadIdx.x; k < n; k += 1024) out[k] = fma(-nu, ws[k], u0[k]);
    if (threadIdx.x == 0) out[n] = nu;
}

// den = t1 . w_s + g   (after w_s = S_s^-1 t1)
__global__ void __launch_bounds__(1024)
precond_den_kernel(const double* t1, const double* ws, KrylovScalars* sc, int n) {
    __shared__ double red[32];
    double s = 0.0;
    for (int k = threadIdx.x; k < n; k += 1024) s = fma(t1[k], ws[k], s);
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 32; ++w) tot += red[w];
        sc->den = tot + sc->g;
    }
}

// Per-geometry scalars from the forward-substituted right-hand-side rows of the non-hydrogen block:
//   g = 1_x^T A_xx^-1 1_x = sum_k s_k y1_k^2,   h = 1_x^T A_xx^-1 c_x = sum_k s_k y1_k yc_k      (k < nxp)
// and the contiguous copies t_c, t_1 of the reduced right-hand sides (rows nhp, nhp + 1 of the snapshot), rhs[n] = Q - h.
__global__ void __launch_bounds__(1024)
krylov_setup_kernel(const double* A, long lda, int NP, int nxp, const double* sgn, const double* S0, long ld0, int nhp,
                    double total_charge, KrylovScalars* sc, double* tc, double* t1, const uint8_t* col_mask,
                    int border_owner) {
    __shared__ double r1[32], r2[32];
    const double* fc = A + NP;
    const double* f1 = fc + 1;
    double a11 = 0.0, a1c = 0.0;
    for (int k = threadIdx.x; k < nxp; k += 1024) {
        const double y1 = f1[(long)k * lda], yc = fc[(long)k * lda], sy = sgn[k] * y1;
        a11 = fma(sy, y1, a11);
        a1c = fma(sy, yc, a1c);
    }
    for (int off = 16; off > 0; off >>= 1) {
        a11 += __shfl_xor_sync(0xffffffffu, a11, off);
        a1c += __shfl_xor_sync(0xffffffffu, a1c, off);
    }
    if ((threadIdx.x & 31) == 0) {
        r1[threadIdx.x >> 5] = a11;
        r2[threadIdx.x >> 5] = a1c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double g = 0.0, h = 0.0;
        for (int w = 0; w < 32; ++w) {
            g += r1[w];
            h += r2[w];
        }
        sc->g = g;
        sc->h = h;
        tc[nhp] = border_owner ? (total_charge - h) : 0.0;
        t1[nhp] = 0.0;
    }
    for (int k = threadIdx.x; k < nhp; k += 1024) {
        const bool mine = !col_mask || col_mask[k / kTile];
        tc[k] = mine ? S0[(long)nhp + (long)k * ld0] : 0.0;
        t1[k] = mine ? S0[(long)nhp + 1 + (long)k * ld0] : 0.0;
    }
}

// d_new[i] = hardness_i f clamp(q_i, +-0.95) (implementation.rs:567-572; 0 on padding); dmax = max |d_new - d_ref|
__global__ void __launch_bounds__(256)
h_shift_kernel(const double* hard_h, const double* q_h, const uint8_t* type_h, int nhp, double* d_new,
               const double* d_ref, KrylovScalars* sc) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    double diff = 0.0;
    if (i < nhp) {
        double d = 0.0;
        if (type_h[i] != kPadType) {
            const double qc = fmin(fmax(q_h[i], -0.95), 0.95);
            d = hard_h[i] * (qc * kHChargeFactor);
        }
        d_new[i] = d;
        diff = fabs(d - d_ref[i]);
    }
    for (int off = 16; off > 0; off >>= 1) diff = fmax(diff, __shfl_xor_sync(0xffffffffu, diff, off));
    if ((threadIdx.x & 31) == 0 && diff > 0.0)
        atomicMax(&sc->dmax_bits, (unsigned long long)__double_as_longlong(diff));
}

// R = I (n x n, ld n): zero the upper triangle by tiles, ones on the diagonal
// R points at the first LOCAL tile row (global tile row row_begin); rows [row_begin, row_end) are held here
__global__ void __launch_bounds__(256)
set_identity_tiles_kernel(double* R, long ld, int nt, int row_begin, int row_end) {
    const int ti = row_begin + blockIdx.x, tj = blockIdx.y;
    if (ti >= row_end || ti > tj) return;   // only tiles on/above the diagonal are ever touched
    double* d = R + (long)(ti - row_begin) * kTile + (long)tj * kTile * ld;
    for (int idx = threadIdx.x; idx < kTile * kTile / 2; idx += 256) {
        const int r2 = idx & 63, c = idx >> 6;
        double2 v = make_double2(0.0, 0.0);
        if (ti == tj) {
            if (2 * r2 == c) v.x = 1.0;
            if (2 * r2 + 1 == c) v.y = 1.0;
        }
        *(double2*)(d + 2 * r2 + (long)c * ld) = v;
    }
}

// v_x = yc_x - mu y1_x (- W^T x_h if wx is given) for the non-hydrogen columns, x_h copied into the solution vector,
// state.mu = mu
__global__ void __launch_bounds__(256)
krylov_finish_kernel(const double* A, long lda, int NP, int nxp, const double* sol, int nhp, double* v, double* x,
                     ScfState* state, const double* wx) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const double mu = sol[nhp];
    if (i < nxp) {
        v[i] = A[(long)NP + (long)i * lda] - mu * A[(long)NP + 1 + (long)i * lda] - (wx ? wx[i] : 0.0);
    } else if (i < NP) {
        x[i] = sol[i - nxp];
    }
    if (i == 0) state->mu = mu;
}

// sol = [x_h; mu] from the direct solve of the iteration that froze the factor
__global__ void __launch_bounds__(256)
krylov_capture_kernel(const double* x, int nxp, int nhp, const ScfState* state, double* sol) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < nhp) sol[i] = x[nxp + i];
    if (i == nhp) sol[nhp] = state->mu;
}

// matrix-free path: rhs = [-(chi + v_ext); Q], e = indicator of the real atoms (0 on padding and in the border slot)
__global__ void __launch_bounds__(256)
matfree_rhs_kernel(const double* chi, const double* vext, const uint8_t* type, int NP, double total_charge, double* rhs,
                   double* e) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < NP) {
        const bool real = type[i] != kPadType;
        rhs[i] = real ? -(chi[i] + (vext ? vext[i] : 0.0)) : 0.0;
        e[i] = real ? 1.0 : 0.0;
    } else if (i == NP) {
        rhs[NP] = total_charge;
        e[NP] = 0.0;
    }
}

// matrix-free path: d_i = hardness_i f clamp(q_i, +-0.95) on the n == 1 atoms (implementation.rs:567-572), else 0
__global__ void __launch_bounds__(256)
matfree_shift_kernel(const double* hard, const double* q, const uint8_t* hflag, int NP, double* d) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= NP) return;
    double v = 0.0;
    if (hflag[i]) v = hard[i] * (fmin(fmax(q[i], -0.95), 0.95) * kHChargeFactor);
    d[i] = v;
}

// x[0..NP) and mu from the bordered solution vector
__global__ void __launch_bounds__(256)
matfree_unpack_kernel(const double* sol, int NP, double* x, ScfState* state) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < NP) x[i] = sol[i];
    if (i == 0) state->mu = sol[NP];
}

}  // namespace

cudaError_t launch_matfree_rhs(const double* chi, const double* vext, const uint8_t* type, int NP, double total_charge,
                               double* rhs, double* e, cudaStream_t stream) {
    matfree_rhs_kernel<<<(NP + 256) / 256, 256, 0, stream>>>(chi, vext, type, NP, total_charge, rhs, e);
    return cudaGetLastError();
}

cudaError_t launch_matfree_shift(const double* hard, const double* q, const uint8_t* hflag, int NP, double* d,
                                 cudaStream_t stream) {
    matfree_shift_kernel<<<(NP + 255) / 256, 256, 0, stream>>>(hard, q, hflag, NP, d);
    return cudaGetLastError();
}

cudaError_t launch_matfree_unpack(const double* sol, int NP, double* x, ScfState* state, cudaStream_t stream) {
    matfree_unpack_kernel<<<(NP + 255) / 256, 256, 0, stream>>>(sol, NP, x, state);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
cudaError_t launch_tile_matvec(const TileMatvecArgs& a, bool do_n, bool do_t, cudaStream_t stream) {
    dim3 grid((a.nt + 7) / 8, a.nt);
    if (a.f32) {
        if (do_n && do_t) tile_matvec_kernel<true, true, float><<<grid, 256, 0, stream>>>(a);
        else if (do_n) tile_matvec_kernel<true, false, float><<<grid, 256, 0, stream>>>(a);
        else tile_matvec_kernel<false, true, float><<<grid, 256, 0, stream>>>(a);
    } else {
        if (do_n && do_t) tile_matvec_kernel<true, true, double><<<grid, 256, 0, stream>>>(a);
        else if (do_n) tile_matvec_kernel<true, false, double><<<grid, 256, 0, stream>>>(a);
        else tile_matvec_kernel<false, true, double><<<grid, 256, 0, stream>>>(a);
    }
    return cudaGetLastError();
}

// Repack the stored tiles of a column-major matrix into tile-major order (see tile_matvec_kernel), optionally
// converting to float.  src points at global tile row 0 (shifted base allowed); tile (i, j) goes to
// dst + col_base[j] + (i - i0(j)) * 128 * 128, i0(j) = upper ? row_begin : j.
template <class TOUT>
__global__ void __launch_bounds__(256)
pack_tiles_kernel(const double* src, long ld, TOUT* dst, const long long* col_base, int nt, int upper, int row_begin,
                  int row_end, const uint8_t* col_mask) {
    const int tj = blockIdx.y;
    if (col_mask && !col_mask[tj]) return;
    const int ti = (upper ? row_begin : tj) + blockIdx.x;
    if (upper ? (ti > tj || ti >= row_end) : (ti >= nt)) return;
    const double* s0 = src + (long)ti * kTile + (long)tj * kTile * ld;
    TOUT* d0 = dst + col_base[tj] + (long)blockIdx.x * (kTile * kTile);
    for (int idx = threadIdx.x; idx < kTile * kTile / 2; idx += 256) {
        const int r2 = idx & 63, c = idx >> 6;
        const double2 v = *(const double2*)(s0 + 2 * r2 + (long)c * ld);
        d0[2 * r2 + c * kTile] = (TOUT)v.x;
        d0[2 * r2 + 1 + c * kTile] = (TOUT)v.y;
    }
}

cudaError_t launch_pack_tiles(const double* src, long ld, void* dst, int to_f32, const long long* col_base, int nt,
                              int upper, int row_begin, int row_end, const uint8_t* col_mask, cudaStream_t stream) {
    const int rows = upper ? (row_end - row_begin) : nt;
    if (rows <= 0) return cudaSuccess;
    dim3 grid(rows, nt);
    if (to_f32)
        pack_tiles_kernel<float><<<grid, 256, 0, stream>>>(src, ld, (float*)dst, col_base, nt, upper, row_begin, row_end, col_mask);
    else
        pack_tiles_kernel<double><<<grid, 256, 0, stream>>>(src, ld, (double*)dst, col_base, nt, upper, row_begin, row_end, col_mask);
    return cudaGetLastError();
}

cudaError_t launch_tile_matvec_reduce(const TileReduceArgs& a, cudaStream_t stream) {
    tile_matvec_reduce_kernel<<<a.nt + (a.op != 0 ? 1 : 0), 512, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_multi_dot(const double* V, long stride, const double* w, int len, double* out, int m,
                             bool with_norm, cudaStream_t stream) {
    const int grid = m + (with_norm ? 1 : 0);
    if (grid <= 0) return cudaSuccess;
    multi_dot_kernel<<<grid, 256, 0, stream>>>(V, stride, w, len, out, m);
    return cudaGetLastError();
}

cudaError_t launch_gs_update(const double* V, long stride, double* w, int len, const double* h, double* hsum, int m,
                             bool first, cudaStream_t stream) {
    gs_update_kernel<<<(len + 255) / 256, 256, 0, stream>>>(V, stride, w, len, h, hsum, m, first ? 1 : 0);
    return cudaGetLastError();
}

cudaError_t launch_scale_by_rnorm(const double* src, double* dst, int len, const double* norm2, cudaStream_t stream) {
    scale_by_rnorm_kernel<<<(len + 255) / 256, 256, 0, stream>>>(src, dst, len, norm2);
    return cudaGetLastError();
}

cudaError_t launch_combine(const double* V, long stride, double* x, int len, const double* y, int m,
                           cudaStream_t stream) {
    if (m <= 0) return cudaSuccess;
    combine_kernel<<<(len + 255) / 256, 256, 0, stream>>>(V, stride, x, len, y, m);
    return cudaGetLastError();
}

cudaError_t launch_precond_border(const double* u0, const double* rin, const double* t1, const double* ws,
                                  const KrylovScalars* sc, double* out, int n, cudaStream_t stream) {
    precond_border_kernel<<<1, 1024, 0, stream>>>(u0, rin, t1, ws, sc, out, n);
    return cudaGetLastError();
}

cudaError_t launch_precond_den(const double* t1, const double* ws, KrylovScalars* sc, int n, cudaStream_t stream) {
    precond_den_kernel<<<1, 1024, 0, stream>>>(t1, ws, sc, n);
    return cudaGetLastError();
}

cudaError_t launch_krylov_setup(const double* A, long lda, int NP, int nxp, const double* sgn, const double* S0,
                                long ld0, int nhp, double total_charge, KrylovScalars* sc, double* tc, double* t1,
                                const uint8_t* col_mask, int border_owner, cudaStream_t stream) {
    krylov_setup_kernel<<<1, 1024, 0, stream>>>(A, lda, NP, nxp, sgn, S0, ld0, nhp, total_charge, sc, tc, t1, col_mask,
                                                border_owner);
    return cudaGetLastError();
}

cudaError_t launch_h_shift(const double* hard_h, const double* q_h, const uint8_t* type_h, int nhp, double* d_new,
                           const double* d_ref, KrylovScalars* sc, cudaStream_t stream) {
    h_shift_kernel<<<(nhp + 255) / 256, 256, 0, stream>>>(hard_h, q_h, type_h, nhp, d_new, d_ref, sc);
    return cudaGetLastError();
}

cudaError_t launch_set_identity_tiles(double* R, long ld, int nt, int row_begin, int row_end, cudaStream_t stream) {
    if (row_end <= row_begin) return cudaSuccess;
    dim3 grid(row_end - row_begin, nt);
    set_identity_tiles_kernel<<<grid, 256, 0, stream>>>(R, ld, nt, row_begin, row_end);
    return cudaGetLastError();
}

cudaError_t launch_krylov_finish(const double* A, long lda, int NP, int nxp, const double* sol, int nhp, double* v,
                                 double* x, ScfState* state, const double* wx, cudaStream_t stream) {
    krylov_finish_kernel<<<(NP + 255) / 256, 256, 0, stream>>>(A, lda, NP, nxp, sol, nhp, v, x, state, wx);
    return cudaGetLastError();
}

cudaError_t launch_p2p_allreduce(const double* src, double* out, int len, const P2PAllreduce& pa, unsigned* done_counter,
                                 int* error, cudaStream_t stream) {
    const int grid = (len + 255) / 256;
    p2p_push_kernel<<<grid, 256, 0, stream>>>(src, len, pa, done_counter);
    p2p_combine_kernel<<<grid, 256, 0, stream>>>(pa, len, out, error);
    return cudaGetLastError();
}

cudaError_t launch_rect_matvec_t(const double* M, long ld, int nt_rows, int nt_cols, int row_begin, int row_end,
                                 const double* x, double* part, double* out, cudaStream_t stream) {
    if (nt_cols <= 0) return cudaSuccess;
    if (row_end > row_begin) {
        dim3 grid((row_end - row_begin + 7) / 8, nt_cols);
        rect_matvec_t_kernel<<<grid, 256, 0, stream>>>(M, ld, nt_rows, row_begin, row_end, x, part);
    }
    rect_matvec_reduce_kernel<<<nt_cols, 512, 0, stream>>>(part, nt_rows, row_begin, row_end, out);
    return cudaGetLastError();
}

cudaError_t launch_krylov_capture(const double* x, int nxp, int nhp, const ScfState* state, double* sol,
                                  cudaStream_t stream) {
    krylov_capture_kernel<<<(nhp + 256) / 256, 256, 0, stream>>>(x, nxp, nhp, state, sol);
    return cudaGetLastError();
}

}  // namespace cheq
