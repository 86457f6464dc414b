// Dense FP64 kernels of the QEq solve for sm_100a.
//
// Replaces `work_matrix.partial_piv_lu().solve(&rhs)` (/root/reference/src/solver/implementation.rs:574-576,
// un-vendored crate faer) and the vector part of run_single_solve / run_scf_iterations (:557-603, :303-339).
// The bordered system is solved by constraint elimination on a blocked Cholesky factorisation
// (SURVEY.md App. E): the right-hand sides ride along as two extra rows of the trapezoid, so the forward
// substitution is part of the factorisation.
//
// FP64 on Blackwell has no tcgen05 kind; the tensor path is mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).  The
// GEMM below stages 128 x 32 operand slabs with cp.async (LDGSTS) into padded shared memory (row stride
// 132 doubles -> conflict-free 16-byte fragment loads), 3 stages (203 KB), 16 warps each owning a 32 x 32
// accumulator tile, 16 DMMAs per 4 LDS.128 (measured: 32-deep slabs +4 % over 16-deep x 4 stages).
#include <cuda_runtime.h>

#include <cstdlib>

#include "device_types.h"
#include "kernels.h"

namespace cheq {
namespace {

// ------------------------------------------------------------------------------------------------------------
// GEMM  C (-)= A * B^T,  A: m x K, B: n x K, column-major, all dimensions multiples of 128 (K of 16)
// ------------------------------------------------------------------------------------------------------------
#ifndef CHEQ_GEMM_BK
#define CHEQ_GEMM_BK 32
#define CHEQ_GEMM_STAGES 3
#endif
constexpr int BM = 128, BN = 128, BK = CHEQ_GEMM_BK, LDT = 132, STAGES = CHEQ_GEMM_STAGES;
constexpr int kGemmThreads = 256;
constexpr int kGemmSmem = STAGES * 2 * BK * LDT * (int)sizeof(double);  // 135168 B

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}

constexpr int kGroupM = 8;   // tile rows per rasterisation group: their A slabs stay in L2 while B streams

// C (-)= A S B^T with S = diag(+-1) over K (pivot signs of A = L S L^T).  The main loop ignores S (measured:
// scaling B fragments in the loop costs 4 % of DMMA throughput, a flagged branch 11 %); negative pivots are rare
// (one per negative eigenvalue), so the epilogue applies  A S B^T = A B^T - 2 sum_{k: s_k = -1} a_k b_k^T  as
// rank-1 corrections to the accumulators, driven by the per-tile lists the factorisation leaf writes.
// FM = 16-row fragments per warp, WM = warps along M: CTA tile (16 FM WM) x 128.
//   <4, 2>: 128 x 128, 8 warps (64 x 32 warp tiles)      <2, 4>: 128 x 128, 16 warps (32 x 32)
//   <2, 2>:  64 x 128, 8 warps: twice the CTAs for skinny updates that would otherwise leave SMs idle
template <int FM, int WM, bool PEERS = false>
__global__ void __launch_bounds__(WM * 4 * 32, 1) gemm_nt_kernel(GemmArgs g) {
    constexpr int kWarpsM = WM;
    constexpr int kThreads = kWarpsM * 4 * 32;        // 256 or 512
    constexpr int TM = 16 * FM * WM;                  // CTA tile rows: 128 or 64
    constexpr int kChunksA = TM * BK / 2 / kThreads;  // 16-byte chunks per thread per stage, A slab (TM x BK)
    constexpr int kChunksB = 64 * BK / kThreads;      //                                      B slab (128 x BK)
    // grouped rasterisation: consecutive CTAs walk the tile columns of a group of kGroupM tile rows
    const int frame = blockIdx.z;
    int tile_m, tile_n;
    {
        const int pid = blockIdx.x;
        const int per_group = kGroupM * g.n_tiles;
        const int group = pid / per_group, rem = pid - group * per_group;
        const int first = group * kGroupM;
        const int gsize = min(kGroupM, g.m_tiles - first);
        tile_n = rem / gsize;
        tile_m = first + (rem - tile_n * gsize);
    }
    if (g.n_list) tile_n = g.n_list[tile_n];                     // column-list mode: absolute tile column
    if (g.lower && g.row0 + tile_m * TM + TM - 1 < tile_n * BN) return;   // tile entirely above the diagonal
    if (g.state && !g.state[frame].active) return;
    extern __shared__ __align__(16) double smem[];
    double* As = smem;                          // [STAGES][BK][LDT]
    double* Bs = smem + STAGES * BK * LDT;      // [STAGES][BK][LDT]

    const double* A = g.A + (long)frame * g.strideA + (long)tile_m * TM;
    const double* B = g.B + (long)frame * g.strideB + (long)tile_n * BN;
    double* C = g.C + (long)frame * g.strideC + (long)tile_m * TM + (long)tile_n * BN * g.ldc;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int wm = warp % kWarpsM, wn = warp / kWarpsM;   // kWarpsM x 4 warps, warp tile (16 FM) x 32
    const int gq = lane >> 2, q = lane & 3;

    auto load_stage = [&](int stage, int ktile) {
        double* as = As + stage * BK * LDT;
        double* bs = Bs + stage * BK * LDT;
        const long k0 = (long)ktile * BK;
#pragma unroll
        for (int i = 0; i < kChunksA; ++i) {
            const int c = t + i * kThreads;
            const int kr = c / (TM / 2), mc = c % (TM / 2);
            cp_async16(as + kr * LDT + 2 * mc, A + 2 * mc + (k0 + kr) * g.lda);
        }
#pragma unroll
        for (int i = 0; i < kChunksB; ++i) {
            const int c = t + i * kThreads;   // 0..1023
            const int kr = c >> 6, mc = c & 63;
            cp_async16(bs + kr * LDT + 2 * mc, B + 2 * mc + (k0 + kr) * g.ldb);
        }
    };

    double acc[FM][2][2][2][2];  // [fm][row parity][fn][col parity][e]
#pragma unroll
    for (int i = 0; i < FM * 16; ++i) ((double*)acc)[i] = 0.0;

    const int kt = g.K / BK;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < kt) load_stage(s, s);
        cp_async_commit();
    }

    for (int it = 0; it < kt; ++it) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nxt = it + STAGES - 1;
            if (nxt < kt) load_stage(nxt % STAGES, nxt);
            cp_async_commit();
        }
        const int stage = it % STAGES;
        const double* as = As + stage * BK * LDT + wm * (16 * FM) + 2 * gq;
        const double* bs = Bs + stage * BK * LDT + wn * 32 + 2 * gq;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double2 a[FM], b[2];
#pragma unroll
            for (int fm = 0; fm < FM; ++fm) a[fm] = *(const double2*)(as + (kk * 4 + q) * LDT + fm * 16);
#pragma unroll
            for (int fn = 0; fn < 2; ++fn) b[fn] = *(const double2*)(bs + (kk * 4 + q) * LDT + fn * 16);
#pragma unroll
            for (int fm = 0; fm < FM; ++fm)
#pragma unroll
                for (int fn = 0; fn < 2; ++fn) {
                    dmma884(acc[fm][0][fn][0][0], acc[fm][0][fn][0][1], a[fm].x, b[fn].x);
                    dmma884(acc[fm][0][fn][1][0], acc[fm][0][fn][1][1], a[fm].x, b[fn].y);
                    dmma884(acc[fm][1][fn][0][0], acc[fm][1][fn][0][1], a[fm].y, b[fn].x);
                    dmma884(acc[fm][1][fn][1][0], acc[fm][1][fn][1][1], a[fm].y, b[fn].y);
                }
        }
    }
    cp_async_wait<0>();

    // negative pivots inside the K range: acc -= 2 a_k b_k^T
    if (g.neg_cnt) {
        const int* cnt = g.neg_cnt + (long)frame * g.strideNeg;
        const int* idx = g.neg_idx + (long)frame * g.strideNeg * kTile;
        for (int kt128 = 0; kt128 < g.K / kTile; ++kt128) {
            const int nneg = cnt[kt128];
            for (int j = 0; j < nneg; ++j) {
                const long k = (long)kt128 * kTile + idx[kt128 * kTile + j];
                const double* ak = A + k * g.lda + wm * (16 * FM) + 2 * gq;
                const double* bk = B + k * g.ldb + wn * 32 + 4 * q;
                double2 av[FM];
                double bv[2][2][2];   // [fn][pc][e]: column 4 q + 2 e + pc of the 16-column group
#pragma unroll
                for (int fm = 0; fm < FM; ++fm) av[fm] = *(const double2*)(ak + fm * 16);
#pragma unroll
                for (int fn = 0; fn < 2; ++fn) {
                    const double2 lo = *(const double2*)(bk + fn * 16), hi = *(const double2*)(bk + fn * 16 + 2);
                    bv[fn][0][0] = lo.x; bv[fn][1][0] = lo.y; bv[fn][0][1] = hi.x; bv[fn][1][1] = hi.y;
                }
#pragma unroll
                for (int fm = 0; fm < FM; ++fm)
#pragma unroll
                    for (int fn = 0; fn < 2; ++fn)
#pragma unroll
                        for (int pc = 0; pc < 2; ++pc)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                acc[fm][0][fn][pc][e] = fma(-2.0 * av[fm].x, bv[fn][pc][e], acc[fm][0][fn][pc][e]);
                                acc[fm][1][fn][pc][e] = fma(-2.0 * av[fm].y, bv[fn][pc][e], acc[fm][1][fn][pc][e]);
                            }
            }
        }
    }

    // epilogue: thread owns rows (2 gq, 2 gq + 1) of each 16-row group and columns 4 q + 2 e + pc of each
    // 16-column group
#pragma unroll
    for (int fm = 0; fm < FM; ++fm) {
        const int row = wm * (16 * FM) + fm * 16 + 2 * gq;
#pragma unroll
        for (int fn = 0; fn < 2; ++fn)
#pragma unroll
            for (int pc = 0; pc < 2; ++pc)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = wn * 32 + fn * 16 + 4 * q + 2 * e + pc;
                    double2* p = (double2*)(C + row + (long)col * g.ldc);
                    double2 v = make_double2(acc[fm][0][fn][pc][e], acc[fm][1][fn][pc][e]);
                    if (g.subtract) {
                        double2 old = *p;
                        v.x = old.x - v.x;
                        v.y = old.y - v.y;
                    }
                    *p = v;
                    if constexpr (PEERS) {
                        const long off = (long)((double*)p - g.C);
                        for (int pe = 0; pe < g.n_peers; ++pe) *(double2*)(g.peerC[pe] + off) = v;
                    }
                }
    }
}

// ------------------------------------------------------------------------------------------------------------
// 128 x 128 diagonal block, one CTA: A = L S L^T (S = diag(+-1): Cholesky that tolerates negative pivots, i.e.
// LDL^T without pivoting; the hardness + Coulomb matrix of a large cluster is symmetric but not always positive
// definite) together with M = S L^-1, which turns the triangular solves of the rows below into a GEMM.
//
// The whole tile lives in registers: thread (ty, tx) of a 16 x 16 grid owns the 64 elements
// (r, c) = (ty + 16 a, tx + 16 b).  For r >= c the element is A -> L; the strict upper part (r < c) holds
// X[c][r], X = L^-1 (forward elimination on the identity).  Step k scales logical column k by s/l -- it is owned by
// one half-warp, which gets the pivot by shuffle -- and publishes it as u[0..127]; then every element with c > k
// and (r >= c or r <= k) takes  e -= s u[r] u[c].  One __syncthreads per step (u is double buffered).
// ------------------------------------------------------------------------------------------------------------
// 1/sqrt(x) for x in the range of pivots (1e-4 .. 1e30): the hardware's double-precision seed
// (rsqrt.approx.ftz.f64 = MUFU.RSQ64H on the high word, ~2^-20 relative) plus ONE Newton step (-> ~2^-40); the
// callers' joint correction of l = x * y and y brings both to full precision.  No f64 <-> f32 conversions, no
// library slow path, no branches: this sits on the per-pivot dependency chain of the diagonal-tile kernels.
__device__ __forceinline__ double rsqrt_pivot(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double h = fma(-t, y, 1.0);
    y = fma(0.5 * y, h, y);
    return y;
}

// debug cycle counters of the diagonal-tile kernels (see cheq_dbg_potrf_cycles)
__device__ unsigned long long g_potrf_cycles[8];
constexpr int kPotrfThreads = 256;
constexpr int kPotrfLd = kTile + 1;
constexpr int kPotrfSmem = kTile * kPotrfLd * (int)sizeof(double);

__global__ void __launch_bounds__(kPotrfThreads)
potrf_tile_kernel(double* A_, long lda, long strideA, double* Minv_, long strideM, double* sgn_, long strideSgn,
                  int* neg_cnt_, int* neg_idx_, long strideNeg, ScfState* state, double pivot_tol, PotrfPeers peers) {
    const int frame = blockIdx.z;
    if (state && !state[frame].active) return;
    extern __shared__ __align__(16) double Ms[];   // staging of M for a coalesced store
    __shared__ double ubuf[2][kTile];
    __shared__ double dinv[kTile];
    __shared__ double sg[kTile];
    double* A = A_ + (long)frame * strideA;
    double* Minv = Minv_ + (long)frame * strideM;
    const int t = threadIdx.x;
    const int ty = t & 15, tx = t >> 4;

    double e[8][8];
#pragma unroll
    for (int b = 0; b < 8; ++b)
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int r = ty + 16 * a, c = tx + 16 * b;
            e[a][b] = (r >= c) ? A[r + (long)c * lda] : 0.0;
        }

    const bool p_diag = ty >= tx;   // within a diagonal 16 x 16 sub-block: element on or below the diagonal
    int cur = 0;
#ifdef CHEQ_POTRF_INSTRUMENT
#define POTRF_CLOCK() clock64()
#else
#define POTRF_CLOCK() 0ll
#endif
    long long tkChain = 0, tkBar = 0, tkUpd = 0, tkAll0 = POTRF_CLOCK();
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
        for (int kt = 0; kt < 16; ++kt) {
            const int k = kb * 16 + kt;
            const long long tc0 = POTRF_CLOCK();
            // pivot: element (k, k) lives in thread (ty = kt, tx = kt), register e[kb][kb]
            const double d = __shfl_sync(0xffffffffu, e[kb][kb], kt + 16 * (tx & 1));
            if (tx == kt) {
                const double ad = fabs(d);
                const bool bad = !(ad > pivot_tol) || !(ad < 1.0e30);
                const double sgnk = (d < 0.0) ? -1.0 : 1.0;
                // 1/sqrt and sqrt from one rsqrt plus a Newton step (shorter dependent chain than sqrt + divide)
                double il = rsqrt_pivot(ad);
                double l = ad * il;
                {
                    const double res = fma(-l, l, ad);        // ad - l^2
                    l = fma(0.5 * il, res, l);
                    il = fma(fma(-l, il, 1.0), il, il);
                }
                il = bad ? 1.0 : il;
                l = bad ? 1.0 : l;
                const double sil = sgnk * il;
                if (ty == kt) {
                    dinv[k] = il;
                    sg[k] = sgnk;
                    if (bad && state) atomicExch(&state[frame].info, k + 1);
                }
                // logical column k: rows r = ty + 16 a.  a < kb: row k of X (r < k); a > kb: L column (r > k);
                // a == kb: decided by ty vs kt.
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const double v = e[a][kb];
                    double u = v * sil, keep;
                    if (a < kb) keep = v * il;
                    else if (a > kb) keep = u;
                    else {
                        keep = (ty > kt) ? u : v * il;
                        if (ty == kt) {
                            u = sil;
                            keep = l;
                        }
                    }
                    e[a][kb] = keep;
                    ubuf[cur][ty + 16 * a] = u;
                }
            }
            const long long tc1 = POTRF_CLOCK();
            __syncthreads();
            const long long tc2 = POTRF_CLOCK();
            const double sk = sg[k];
            const bool p_row = ty <= kt;    // rows of block kb that are <= k
            const bool p_col = tx > kt;     // columns of block kb that are > k
            double su[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) su[a] = sk * ubuf[cur][ty + 16 * a];
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                if (b < kb) continue;                 // columns <= k: finished
                if (b == kb && !p_col) continue;      // same block, column <= k
                const double uc = ubuf[cur][tx + 16 * b];
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    // update iff (r >= c) [L part] or (r <= k) [X part]; resolved per (a, b, kb) at compile time
                    // except inside the diagonal sub-block (p_diag) and the pivot's row block (p_row)
                    bool upd;
                    if (a > b) upd = true;
                    else if (a < kb) upd = true;
                    else if (a == b && a == kb) upd = p_diag || p_row;
                    else if (a == b) upd = p_diag;
                    else if (a == kb) upd = p_row;
                    else upd = false;
                    if (upd) e[a][b] = fma(-su[a], uc, e[a][b]);
                }
            }
            cur ^= 1;
            if (tx == kt && ty == kt) {   // the pivot thread: chain, barrier wait and update time of this step
                tkChain += tc1 - tc0;
                tkBar += tc2 - tc1;
                tkUpd += POTRF_CLOCK() - tc2;
            }
        }
    }
    __syncthreads();
#ifdef CHEQ_POTRF_INSTRUMENT
    if (tkChain) {
        atomicAdd(&g_potrf_cycles[1], (unsigned long long)tkChain);
        atomicAdd(&g_potrf_cycles[2], (unsigned long long)tkBar);
        atomicAdd(&g_potrf_cycles[3], (unsigned long long)tkUpd);
    }
    if (t == 0) {
        atomicAdd(&g_potrf_cycles[0], (unsigned long long)(clock64() - tkAll0));
        atomicAdd(&g_potrf_cycles[5], 1ull);
    }
#else
    (void)tkChain; (void)tkBar; (void)tkUpd; (void)tkAll0;
#endif

    // L back to global; M = S X staged through shared memory (M[i][j] = s_i X[i][j], i > j, held as element (j, i))
#pragma unroll
    for (int b = 0; b < 8; ++b)
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int r = ty + 16 * a, c = tx + 16 * b;
            if (r >= c) {
                A[r + (long)c * lda] = e[a][b];
                for (int pe = 0; pe < peers.n; ++pe) peers.A[pe][r + (long)c * lda] = e[a][b];
            }
            // Ms[j * ld + i] = M[i][j]; the holder of (r, c) fills M[c][r]: X part below the diagonal of M,
            // the diagonal, and zeros above it
            double m = 0.0;
            if (r < c) m = sg[c] * e[a][b];
            else if (r == c) m = sg[r] * dinv[r];
            Ms[r * kPotrfLd + c] = m;
        }
    __syncthreads();
    for (int idx = t; idx < kTile * kTile; idx += kPotrfThreads) {
        const int i = idx & (kTile - 1), j = idx >> 7;
        const double m = Ms[j * kPotrfLd + i];
        Minv[idx] = m;
        for (int pe = 0; pe < peers.n; ++pe) peers.Minv[pe][idx] = m;
    }
    if (t < kTile && sgn_) {
        sgn_[(long)frame * strideSgn + t] = sg[t];
        for (int pe = 0; pe < peers.n; ++pe) peers.sgn[pe][t] = sg[t];
    }
    if (t == 0 && neg_cnt_) {   // compact list of this tile's negative pivots for the GEMM epilogues
        int n = 0;
        int* idx = neg_idx_ + (long)frame * strideNeg * kTile;
        for (int i = 0; i < kTile; ++i)
            if (sg[i] < 0.0) {
                for (int pe = 0; pe < peers.n; ++pe) peers.neg_idx[pe][n] = i;
                idx[n++] = i;
            }
        neg_cnt_[(long)frame * strideNeg] = n;
        for (int pe = 0; pe < peers.n; ++pe) peers.neg_cnt[pe][0] = n;
    }
}

// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
copy_lower_tiles_kernel(const double* src, long lds, long stride_src, double* dst, long ldd, long stride_dst,
                        const ScfState* state, const uint8_t* col_mask) {
    const int ti = blockIdx.x, tj = blockIdx.y, frame = blockIdx.z;
    if (ti < tj) return;
    if (col_mask && !col_mask[tj]) return;
    if (state && !state[frame].active) return;
    const double* s = src + (long)frame * stride_src + (long)ti * kTile + (long)tj * kTile * lds;
    double* d = dst + (long)frame * stride_dst + (long)ti * kTile + (long)tj * kTile * ldd;
    for (int idx = threadIdx.x; idx < kTile * kTile / 2; idx += 256) {
        const int r2 = idx & 63, c = idx >> 6;
        *(double2*)(d + 2 * r2 + (long)c * ldd) = *(const double2*)(s + 2 * r2 + (long)c * lds);
    }
}

__global__ void set_h_diagonal_kernel(double* A, long lda, long strideA, const double* S0, long ld0,
                                      long stride0, const double* hard_h, const double* q_h, long q_stride,
                                      const uint8_t* type_h, int nhp, const ScfState* state, const uint8_t* col_mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int frame = blockIdx.z;
    if (i >= nhp) return;
    if (col_mask && !col_mask[i / kTile]) return;   // column owned by another rank
    if (state && !state[frame].active) return;
    if (type_h[i] == kPadType) return;
    // implementation.rs:567-572: hardness * (1 + clamp(q, -0.95, 0.95) * H_CHARGE_DEPENDENCE_FACTOR)
    const double q = q_h[(long)frame * q_stride + i];
    const double qc = fmin(fmax(q, -0.95), 0.95);
    const double base = S0[(long)frame * stride0 + (long)i + (long)i * ld0];
    A[(long)frame * strideA + (long)i + (long)i * lda] = base + hard_h[i] * (qc * kHChargeFactor);
}

// mu and the combined right-hand side.  One CTA per frame; fixed summation order.
__global__ void __launch_bounds__(1024)
mu_kernel(const double* A_, long lda, long strideA, int NP, double total_charge, double* v_, long v_stride,
          const double* sgn_, long strideSgn, ScfState* state) {
    const int frame = blockIdx.z;
    if (!state[frame].active) return;
    const double* fc = A_ + (long)frame * strideA + NP;   // row NP
    const double* f1 = fc + 1;                            // row NP + 1
    double* v = v_ + (long)frame * v_stride;
    const double* sgn = sgn_ + (long)frame * strideSgn;
    __shared__ double s11[32], s1c[32];
    __shared__ double mu_s;
    double a11 = 0.0, a1c = 0.0;
    for (int j = threadIdx.x; j < NP; j += 1024) {
        // rows hold y = S L^-1 b, so b^T A^-1 c = sum_k s_k y_b[k] y_c[k]
        const double x1 = f1[(long)j * lda], xc = fc[(long)j * lda], sx1 = sgn[j] * x1;
        a11 = fma(sx1, x1, a11);
        a1c = fma(sx1, xc, a1c);
    }
    for (int off = 16; off > 0; off >>= 1) {
        a11 += __shfl_xor_sync(0xffffffffu, a11, off);
        a1c += __shfl_xor_sync(0xffffffffu, a1c, off);
    }
    if ((threadIdx.x & 31) == 0) {
        s11[threadIdx.x >> 5] = a11;
        s1c[threadIdx.x >> 5] = a1c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t11 = 0.0, t1c = 0.0;
        for (int w = 0; w < 32; ++w) {
            t11 += s11[w];
            t1c += s1c[w];
        }
        const double mu = (t1c - total_charge) / t11;
        mu_s = mu;
        state[frame].mu = mu;
    }
    __syncthreads();
    const double mu = mu_s;
    for (int j = threadIdx.x; j < NP; j += 1024) v[j] = fc[(long)j * lda] - mu * f1[(long)j * lda];
}

// Backward substitution L^T x = v, right-looking, block row kb (from the last to the first):
//   x_kb = L_kk^-T v_kb = M_kk^T (S v_kb);   v_j -= L[kb rows, j]^T x_kb  for every column j < 128 kb.
// Every CTA recomputes x_kb (128 x 128 triangular product, M from L2) and then owns 64 columns of the update, one
// warp per column group: no cross-CTA reduction, fixed summation order.  CTA 0 also stores x_kb.
constexpr int kBackCols = 64;
__global__ void __launch_bounds__(256)
backsolve_step_kernel(const double* A_, long lda, long strideA, const double* Minv_, long strideM,
                      const double* sgn_, long strideSgn, int kb, double* v_, double* x_, long v_stride,
                      const ScfState* state) {
    const int frame = blockIdx.z;
    if (state && !state[frame].active) return;
    const double* A = A_ + (long)frame * strideA;
    const double* Minv = Minv_ + (long)frame * strideM + (long)kb * kTile * kTile;
    const double* sgn = sgn_ + (long)frame * strideSgn + kb * kTile;
    double* v = v_ + (long)frame * v_stride;
    double* x = x_ + (long)frame * v_stride;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    __shared__ double tv[kTile];
    __shared__ double xs[kTile];
    if (t < kTile) tv[t] = sgn[t] * v[kb * kTile + t];
    __syncthreads();
    for (int i = warp; i < kTile; i += 8) {
        // x_i = sum_{j >= i} M[j, i] (S v)[j]
        const double* Xc = Minv + (long)i * kTile;
        double s = 0.0;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = lane + 32 * jj;
            if (j >= i) s = fma(Xc[j], tv[j], s);
        }
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) xs[i] = s;
    }
    __syncthreads();
    if (blockIdx.x == 0 && t < kTile) x[kb * kTile + t] = xs[t];
    // update of the earlier blocks: this CTA's columns [c0, c0 + 64)
    const int c0 = blockIdx.x * kBackCols;
    if (c0 >= kb * kTile) return;
    const double x0 = xs[lane], x1 = xs[lane + 32], x2 = xs[lane + 64], x3 = xs[lane + 96];
    const double* Lr = A + (long)kb * kTile + lane;   // rows of block kb
#pragma unroll
    for (int cc = 0; cc < kBackCols / 8; cc += 2) {
        const int ja = c0 + warp * (kBackCols / 8) + cc, jb = ja + 1;
        const double* La = Lr + (long)ja * lda;
        const double* Lb = Lr + (long)jb * lda;
        double sa = La[0] * x0, sb = Lb[0] * x0;
        sa = fma(La[32], x1, sa); sb = fma(Lb[32], x1, sb);
        sa = fma(La[64], x2, sa); sb = fma(Lb[64], x2, sb);
        sa = fma(La[96], x3, sa); sb = fma(Lb[96], x3, sb);
        for (int off = 16; off > 0; off >>= 1) {
            sa += __shfl_xor_sync(0xffffffffu, sa, off);
            sb += __shfl_xor_sync(0xffffffffu, sb, off);
        }
        if (lane == 0) {
            v[ja] -= sa;
            v[jb] -= sb;
        }
    }
}

// Backward substitution L^T x = v as ONE kernel: a chain of CTAs linked by release/acquire flags in global memory.
// CTA b (one per 128-column block; blocks are claimed by ticket from the last to the first) owns v_b privately:
//   for kb = last .. b + 1 (as soon as x_kb is published):  acc_b += L[kb rows, b cols]^T x_kb   (HBM stream,
//     128 KB per tile, per-lane partial sums: no cross-lane reduction inside the loop)
//   then v_b -= reduce(acc_b);  x_b = M_b^T (S v_b)  with M_b = S L_bb^-1 staged in shared memory at kernel start;
//   publish x_b, __threadfence, flag[b] = epoch.
// Only the CTA right below the chain head is latency-critical (one tile + one triangular product per link,
// ~2-3 us); all others stream their tiles at HBM speed behind it.  A CTA waits only on CTAs with a smaller
// ticket (tickets are handed out as CTAs start), so the chain cannot deadlock even when the grid exceeds the
// number of resident CTAs.  Fixed summation order: results are bit-reproducible (the replicated solves of all ranks agree).
constexpr int kChainPacked = kTile * (kTile + 1) / 2;                 // lower triangle of M_b, packed by row
constexpr int kChainRedLd = 17;
constexpr int kChainSmem = (kChainPacked + 8 * 32 * kChainRedLd + 4 * kTile) * (int)sizeof(double);

__global__ void __launch_bounds__(256, 2)
backsolve_chain_kernel(const double* A_, long lda, long strideA, const double* Minv_, long strideM,
                       const double* sgn_, long strideSgn, int nb, int nb_solve, const double* v_, double* x_,
                       long v_stride, unsigned* flags_, unsigned* tickets, const ScfState* state) {
    const int frame = blockIdx.z;
    if (state && !state[frame].active) return;
    // block index by ticket: the order in which CTAs start IS the order of the chain (no assumption about the
    // hardware's dispatch order)
    __shared__ int b_s;
    if (threadIdx.x == 0) b_s = nb_solve - 1 - (int)atomicAdd(&tickets[frame], 1u);
    __syncthreads();
    const int b = b_s;
    const unsigned epoch = 1u;
    extern __shared__ __align__(16) double csm[];
    double* Mp = csm;                                  // Mp[j (j + 1) / 2 + i] = M[j][i], i <= j
    double* red = Mp + kChainPacked;                   // [8 warps][32 lanes][17]
    double* tv = red + 8 * 32 * kChainRedLd;           // [128] S v_b
    double* xs = tv + kTile;                           // [128] x_kb of the tile being applied / x_b
    double* part = xs + kTile;                         // [2][128]
    __shared__ int ready_s;
    const double* A = A_ + (long)frame * strideA;
    const double* Minv = Minv_ + (long)frame * strideM + (long)b * kTile * kTile;
    const double* sgn = sgn_ + (long)frame * strideSgn + b * kTile;
    const double* v = v_ + (long)frame * v_stride;
    double* x = x_ + (long)frame * v_stride;
    volatile unsigned* flags = flags_ + (long)frame * nb;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;

    // stage M_b (lower triangle) while the blocks above are still being solved
    for (int idx = t; idx < kTile * kTile; idx += 256) {
        const int j = idx & (kTile - 1), i = idx >> 7;   // row j, column i of M (column-major in global memory)
        if (j >= i) Mp[j * (j + 1) / 2 + i] = Minv[idx];
    }

    // per-lane partial sums of the 16 columns of this warp: column b 128 + 16 warp + c, rows lane + 32 m
    double acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = 0.0;
    const double* Lcol = A + ((long)b * kTile + 16 * warp) * lda + lane;
    int kb = nb - 1;
    while (kb > b) {
        // how many of the next (up to 32) blocks are already published?
        if (warp == 0) {
            const int k = kb - lane;
            unsigned ok = 0;
            do {
                // blocks from nb_solve on were solved elsewhere (stale-factor iterations: x_h comes from GMRES)
                const bool mine_ready = (k > b) ? (k >= nb_solve || flags[k] == epoch) : true;
                ok = __ballot_sync(0xffffffffu, mine_ready);
            } while ((ok & 1u) == 0u);                   // at least block kb must be there
            const int run = __ffs(~ok) - 1;              // consecutive ready blocks from kb downwards (32 if all)
            if (lane == 0) ready_s = (run < 0) ? 32 : run;
        }
        __syncthreads();
        __threadfence();                                 // acquire: x of the published blocks is visible
        int nready = ready_s;
        if (nready > kb - b) nready = kb - b;
        for (int q = 0; q < nready; ++q, --kb) {
            const double* xk = x + (long)kb * kTile;
            const double x0 = __ldcg(xk + lane), x1 = __ldcg(xk + lane + 32), x2 = __ldcg(xk + lane + 64),
                         x3 = __ldcg(xk + lane + 96);
            const double* Lt = Lcol + (long)kb * kTile;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const double* p = Lt + (long)c * lda;
                const double l0 = __ldcs(p), l1 = __ldcs(p + 32), l2 = __ldcs(p + 64), l3 = __ldcs(p + 96);
                acc[c] = fma(l0, x0, acc[c]);
                acc[c] = fma(l1, x1, acc[c]);
                acc[c] = fma(l2, x2, acc[c]);
                acc[c] = fma(l3, x3, acc[c]);
            }
        }
        __syncthreads();                                 // ready_s is rewritten in the next round
    }
    // reduce the per-lane partials: red[warp][lane][c]
#pragma unroll
    for (int c = 0; c < 16; ++c) red[(warp * 32 + lane) * kChainRedLd + c] = acc[c];
    __syncthreads();
    if (t < kTile) {
        const int w = t >> 4, c = t & 15;
        double s = 0.0;
#pragma unroll 8
        for (int l = 0; l < 32; ++l) s += red[(w * 32 + l) * kChainRedLd + c];
        tv[t] = sgn[t] * (v[b * kTile + t] - s);
    }
    __syncthreads();
    {   // x_i = sum_{j >= i} M[j][i] (S v)_j : thread (i, h) takes the rows j of parity h
        const int i = t & (kTile - 1), h = t >> 7;
        double s = 0.0;
        int j = i + ((i + h) & 1);                       // first row >= i with j % 2 == h
        for (; j < kTile; j += 2) s = fma(Mp[j * (j + 1) / 2 + i], tv[j], s);
        part[h * kTile + i] = s;
    }
    __syncthreads();
    if (t < kTile) {
        const double xi = part[t] + part[kTile + t];
        __stcg(x + (long)b * kTile + t, xi);
        __threadfence();
    }
    __syncthreads();
    if (t == 0) {
        __threadfence();
        flags[b] = epoch;
    }
}

__global__ void __launch_bounds__(256)
scf_mix_kernel(const double* x_, double* q_, long stride, const uint8_t* type, int NP, ScfState* state) {
    const int frame = blockIdx.z;
    if (!state[frame].active) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double d = state[frame].damping;
    double diff = 0.0;
    if (i < NP && type[i] != kPadType) {
        const double xn = x_[(long)frame * stride + i];
        const double qo = q_[(long)frame * stride + i];
        diff = fabs(xn - qo);                                   // implementation.rs:589-594
        q_[(long)frame * stride + i] = (1.0 - d) * qo + d * xn;  // implementation.rs:596-598
    }
    for (int off = 16; off > 0; off >>= 1) diff = fmax(diff, __shfl_xor_sync(0xffffffffu, diff, off));
    __shared__ double wmax[8];
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = diff;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int w = 0; w < 8; ++w) m = fmax(m, wmax[w]);
        atomicMax(&state[frame].delta_bits, (unsigned long long)__double_as_longlong(m));
    }
}

__global__ void scf_decide_kernel(ScfState* state, double tolerance, int damping_kind, int batch,
                                  int* active_count) {
    const int frame = blockIdx.x * blockDim.x + threadIdx.x;
    if (frame >= batch) return;
    ScfState& s = state[frame];
    if (!s.active) return;
    const double delta = __longlong_as_double((long long)s.delta_bits);
    s.delta_bits = 0ull;
    s.delta = delta;
    s.iteration += 1;
    if (s.info != 0) {
        s.active = 0;
        return;
    }
    if (delta < tolerance) {            // implementation.rs:318
        s.converged = 1;
        s.active = 0;
        return;
    }
    if (damping_kind == 2) {            // implementation.rs:326-336
        double d = s.damping;
        if (delta > s.prev_delta) d *= 0.5;
        else if (delta < s.prev_delta * 0.9) d = fmin(d * 1.1, 1.0);
        if (d < 0.001) d = 0.001;
        s.damping = d;
    }
    s.prev_delta = delta;
    atomicAdd(active_count, 1);
}

__global__ void scf_init_kernel(ScfState* state, double damping, int batch) {
    const int frame = blockIdx.x * blockDim.x + threadIdx.x;
    if (frame >= batch) return;
    ScfState s;
    s.damping = damping;
    s.prev_delta = 1.7976931348623157e308;   // f64::MAX, implementation.rs:295
    s.delta = 0.0;
    s.mu = 0.0;
    s.delta_bits = 0ull;
    s.iteration = 0;
    s.converged = 0;
    s.info = 0;
    s.active = 1;
    state[frame] = s;
}

}  // namespace

// FP64 FMA throughput probe (plain DFMA pipe, not the tensor path): 8 independent chains per thread.
__global__ void __launch_bounds__(256) dfma_rate_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double x = 1.0000001, y = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
        a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// DMMA.8x8x4 issue-rate probe: 32 independent accumulator pairs per warp, operands in registers.
__global__ void dmma_rate_kernel(double* out, int iters) {
    double acc[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = threadIdx.x * 1e-9 + i;
    double a = 1.0000001 + threadIdx.x * 1e-12, b = 0.5 + threadIdx.x * 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma884(acc[i][0], acc[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Dependent-issue latencies (one warp, one chain): DFMA, shuffle of a double, DMMA, shared-memory load
__global__ void latency_probe_kernel(double* out, long long* cycles, int iters) {
    __shared__ double sm[64];
    sm[threadIdx.x] = 1.0 + threadIdx.x * 1e-9;
    sm[threadIdx.x + 32] = 0.5;
    __syncthreads();
    double a = 1.0 + threadIdx.x * 1e-9;
    const double x = 1.0000001, y = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) a = fma(a, x, y);
    long long t1 = clock64();
    double b = a;
    for (int i = 0; i < iters; ++i) b = __shfl_sync(0xffffffffu, b, (threadIdx.x + 1) & 31);
    long long t2 = clock64();
    double c0 = b, c1 = a;
    for (int i = 0; i < iters; ++i) dmma884(c0, c1, c0, x);
    long long t3 = clock64();
    int idx = threadIdx.x;
    double d = c0;
    for (int i = 0; i < iters; ++i) {
        d += sm[idx];
        idx = (idx + (int)(d > 1e300)) & 31;   // address depends on the loaded value
    }
    long long t4 = clock64();
    float f = (float)d;
    for (int i = 0; i < iters; ++i) f = rsqrtf(f) + 1.0f;
    long long t5 = clock64();
    out[threadIdx.x] = a + b + c0 + c1 + d + f;
    if (threadIdx.x == 0) {
        cycles[0] = t1 - t0;
        cycles[1] = t2 - t1;
        cycles[2] = t3 - t2;
        cycles[3] = t4 - t3;
        cycles[4] = t5 - t4;
    }
}

cudaError_t launch_latency_probe(double* out, long long* cycles, int iters, cudaStream_t stream) {
    latency_probe_kernel<<<1, 32, 0, stream>>>(out, cycles, iters);
    return cudaGetLastError();
}

cudaError_t launch_dmma_rate(double* out, int blocks, int threads, int iters, cudaStream_t stream) {
    dmma_rate_kernel<<<blocks, threads, 0, stream>>>(out, iters);
    return cudaGetLastError();
}

cudaError_t launch_dfma_rate(double* out, int blocks, int iters, cudaStream_t stream) {
    dfma_rate_kernel<<<blocks, 256, 0, stream>>>(out, iters);
    return cudaGetLastError();
}

cudaError_t potrf_debug_cycles(unsigned long long* out, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out, g_potrf_cycles, sizeof(g_potrf_cycles));
    if (e == cudaSuccess && reset) {
        unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        e = cudaMemcpyToSymbol(g_potrf_cycles, z, sizeof(z));
    }
    return e;
}

int g_gemm_warps = 16;        // warps per 128 x 128 GEMM CTA: 8 (64 x 32 warp tiles) or 16 (32 x 32)
int g_gemm_half_tiles = -1;   // -1: heuristic; 0 / 1: force 128- / 64-row tiles (tuning)
void set_gemm_warps(int w) {
    if (w == 8 || w == 16) g_gemm_warps = w;
    if (w == 64) g_gemm_half_tiles = 1;
    if (w == 128) g_gemm_half_tiles = 0;
    if (w == 0) g_gemm_half_tiles = -1;
}

// Opt-in to > 48 KB of dynamic shared memory.  The attribute is PER DEVICE: cheq_ctx_create calls this after
// cudaSetDevice for every context, so a second GPU in the same process is configured too.
cudaError_t configure_dense_kernels() {
    cudaError_t e = cudaFuncSetAttribute(gemm_nt_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_nt_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_nt_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_nt_kernel<2, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_nt_kernel<2, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(potrf_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPotrfSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(backsolve_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmem);
    return e;
}

cudaError_t launch_gemm_nt(const GemmArgs& g, int m_tiles, int n_tiles, int batch, cudaStream_t stream) {
    if (m_tiles <= 0 || n_tiles <= 0) return cudaSuccess;
    GemmArgs args = g;
    args.m_tiles = m_tiles;
    args.n_tiles = n_tiles;
    dim3 grid(m_tiles * n_tiles, 1, batch);
    // skinny problems: 64-row tiles double the CTA count when 128-row tiles would leave SMs idle
    const long tiles128 = g.lower ? (long)m_tiles * n_tiles - (long)n_tiles * (n_tiles - 1) / 2 : (long)m_tiles * n_tiles;
    const bool half = (g_gemm_half_tiles < 0) ? (tiles128 * batch <= 74) : (g_gemm_half_tiles != 0);
    if (args.n_peers > 0) {   // leaf solves of a distributed panel: tiles are also pushed to the peers
        if (half) {
            args.m_tiles = 2 * m_tiles;
            grid.x = 2 * m_tiles * n_tiles;
            gemm_nt_kernel<2, 2, true><<<grid, 256, kGemmSmem, stream>>>(args);
        } else {
            gemm_nt_kernel<2, 4, true><<<grid, 512, kGemmSmem, stream>>>(args);
        }
        return cudaGetLastError();
    }
    if (half) {
        args.m_tiles = 2 * m_tiles;
        grid.x = 2 * m_tiles * n_tiles;
        gemm_nt_kernel<2, 2><<<grid, 256, kGemmSmem, stream>>>(args);
    } else if (g_gemm_warps == 8) {
        gemm_nt_kernel<4, 2><<<grid, 256, kGemmSmem, stream>>>(args);
    } else {
        gemm_nt_kernel<2, 4><<<grid, 512, kGemmSmem, stream>>>(args);
    }
    return cudaGetLastError();
}

__global__ void signal_peers_kernel(PotrfPeers words, unsigned value) {
    // the panel's tiles were pushed by earlier kernels of this stream; the fence orders this (system-scope) flag
    // store after them for observers on other GPUs
    __threadfence_system();
    const int pe = threadIdx.x;
    if (pe < words.n) {
        volatile unsigned* w = (volatile unsigned*)words.neg_cnt[pe];
        *w = value;
    }
    __threadfence_system();
}

cudaError_t launch_signal_peers(unsigned* const* peer_words, int n_peers, unsigned value, cudaStream_t stream) {
    PotrfPeers w{};
    w.n = n_peers;
    for (int i = 0; i < n_peers; ++i) w.neg_cnt[i] = (int*)peer_words[i];
    signal_peers_kernel<<<1, 32, 0, stream>>>(w, value);
    return cudaGetLastError();
}

cudaError_t launch_potrf_tile(double* A, long lda, long strideA, double* Linv, long strideLinv, double* sgn,
                              long strideSgn, int* neg_cnt, int* neg_idx, long strideNeg, ScfState* state,
                              double pivot_tol, int batch, cudaStream_t stream, const PotrfPeers* peers) {
    PotrfPeers pp{};
    if (peers) pp = *peers;
    dim3 grid(1, 1, batch);
    potrf_tile_kernel<<<grid, kPotrfThreads, kPotrfSmem, stream>>>(A, lda, strideA, Linv, strideLinv, sgn, strideSgn,
                                                                   neg_cnt, neg_idx, strideNeg, state, pivot_tol, pp);
    return cudaGetLastError();
}

cudaError_t launch_copy_lower_tiles(const double* src, long lds, long stride_src, double* dst, long ldd,
                                    long stride_dst, int m_tiles, int n_tiles, const ScfState* state, int batch,
                                    cudaStream_t stream, const uint8_t* col_mask) {
    if (m_tiles <= 0 || n_tiles <= 0) return cudaSuccess;
    dim3 grid(m_tiles, n_tiles, batch);
    copy_lower_tiles_kernel<<<grid, 256, 0, stream>>>(src, lds, stride_src, dst, ldd, stride_dst, state, col_mask);
    return cudaGetLastError();
}

cudaError_t launch_set_h_diagonal(double* A, long lda, long strideA, const double* S0, long ld0, long stride0,
                                  const double* hard_h, const double* q_h, long q_stride, const uint8_t* type_h,
                                  int nhp, const ScfState* state, int batch, cudaStream_t stream,
                                  const uint8_t* col_mask) {
    if (nhp <= 0) return cudaSuccess;
    dim3 grid((nhp + 255) / 256, 1, batch);
    set_h_diagonal_kernel<<<grid, 256, 0, stream>>>(A, lda, strideA, S0, ld0, stride0, hard_h, q_h, q_stride,
                                                    type_h, nhp, state, col_mask);
    return cudaGetLastError();
}

cudaError_t launch_mu(const double* A, long lda, long strideA, int NP, double total_charge, double* v,
                      long v_stride, const double* sgn, long strideSgn, ScfState* state, int batch,
                      cudaStream_t stream) {
    dim3 grid(1, 1, batch);
    mu_kernel<<<grid, 1024, 0, stream>>>(A, lda, strideA, NP, total_charge, v, v_stride, sgn, strideSgn, state);
    return cudaGetLastError();
}

int backsolve_max_ctas() { return 1; }

cudaError_t launch_backsolve_chain(const double* A, long lda, long strideA, const double* Linv, long strideLinv,
                                   const double* sgn, long strideSgn, int NP, const double* v, double* x,
                                   long v_stride, unsigned* work, const ScfState* state, int batch,
                                   cudaStream_t stream, int nb_solve) {
    const int nb = NP / kTile;
    if (nb_solve < 0 || nb_solve > nb) nb_solve = nb;
    if (nb_solve == 0) return cudaSuccess;
    // work = [batch] tickets, then [batch][nb] flags
    cudaError_t e = cudaMemsetAsync(work, 0, (size_t)batch * (nb + 1) * sizeof(unsigned), stream);
    if (e != cudaSuccess) return e;
    dim3 grid(nb_solve, 1, batch);
    backsolve_chain_kernel<<<grid, 256, kChainSmem, stream>>>(A, lda, strideA, Linv, strideLinv, sgn, strideSgn, nb,
                                                              nb_solve, v, x, v_stride, work + batch, work, state);
    return cudaGetLastError();
}

cudaError_t launch_backsolve_step(const double* A, long lda, long strideA, const double* Linv, long strideLinv,
                                  const double* sgn, long strideSgn, int NP, int kb, double* v, double* x,
                                  long v_stride, const ScfState* state, int batch, cudaStream_t stream) {
    (void)NP;
    int G = (kb * kTile) / kBackCols;   // 64 columns of the update per CTA
    if (G < 1) G = 1;
    dim3 grid(G, 1, batch);
    backsolve_step_kernel<<<grid, 256, 0, stream>>>(A, lda, strideA, Linv, strideLinv, sgn, strideSgn, kb, v, x,
                                                    v_stride, state);
    return cudaGetLastError();
}

cudaError_t launch_scf_mix(const double* x, double* q, long stride, const uint8_t* type, int NP, ScfState* state,
                           int batch, cudaStream_t stream) {
    dim3 grid((NP + 255) / 256, 1, batch);
    scf_mix_kernel<<<grid, 256, 0, stream>>>(x, q, stride, type, NP, state);
    return cudaGetLastError();
}

cudaError_t launch_scf_decide(ScfState* state, double tolerance, int damping_kind, int batch, int* active_count,
                              cudaStream_t stream) {
    scf_decide_kernel<<<(batch + 127) / 128, 128, 0, stream>>>(state, tolerance, damping_kind, batch,
                                                                active_count);
    return cudaGetLastError();
}

cudaError_t launch_scf_init(ScfState* state, double damping, int batch, cudaStream_t stream) {
    scf_init_kernel<<<(batch + 127) / 128, 128, 0, stream>>>(state, damping, batch);
    return cudaGetLastError();
}

}  // namespace cheq
