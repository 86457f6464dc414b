// PARKED EXPERIMENTS -- not part of the product build (cheq_b200/build.py does not compile this file).
//
// Four variants of the 128 x 128 diagonal-tile factorisation that were written, verified against
// potrf_tile_kernel and measured SLOWER on B200 in round 1 (profiles/r1_potrf_variants.txt):
//   potrf_tile_pipe_kernel     software-pipelined owners            +4 %
//   potrf_tile_cw_kernel       dedicated chain warp                 +8 %
//   potrf_tile_dmma_kernel     rank-4 DMMA updates in registers     2.0x (local-memory spills)
//   potrf_tile_blocked_kernel  16-column blocks in shared memory    1.3x (shared-memory bound)
// Kept as a record of what was tried.  The fragment below is the body that used to sit in kernels_dense.cu
// between potrf_tile_kernel and copy_lower_tiles_kernel; it relies on that file's helpers (rsqrt_pivot,
// dmma884, g_potrf_cycles, kPotrfThreads, kPotrfLd, kPotrfSmem, PotrfPeers).
#if 0
// Software-pipelined version of potrf_tile_kernel (experimental, CHEQ_POTRF_PIPE=1): identical arithmetic, but the per-pivot dependency
// chain (pivot broadcast -> rsqrt -> column scaling) of pivot k + 1 is started by the 16 threads that own column
// k + 1 as soon as THEY have applied update k to that column, and runs while all other warps are still doing the
// bulk of update k.  One __syncthreads per pivot as before; the step time drops from chain + update to
// max(chain, update) + the owners' own share.
#define POTRF_UPD(a, b, kb) \
    ((a) > (b) ? true : (a) < (kb) ? true : ((a) == (b) && (a) == (kb)) ? (p_diag || p_row) : (a) == (b) ? p_diag : (a) == (kb) ? p_row : false)

template <int KBN>
__device__ __forceinline__ void potrf_next_pivot(double (&e)[8][8], int ty, int tx, int ktn, double* ubuf_next,
                                                 double* dinv, double* sg, ScfState* state, int frame,
                                                 double pivot_tol) {
    // executed by the 16 threads with tx == ktn (one half-warp): pivot k' = 16 KBN + ktn
    const unsigned mask = (tx & 1) ? 0xffff0000u : 0x0000ffffu;
    const int k = KBN * 16 + ktn;
    const double d = __shfl_sync(mask, e[KBN][KBN], ktn + 16 * (tx & 1));
    const double ad = fabs(d);
    const bool bad = !(ad > pivot_tol) || !(ad < 1.0e30);
    const double sgnk = (d < 0.0) ? -1.0 : 1.0;
    double il = rsqrt_pivot(ad);
    double l = ad * il;
    {
        const double res = fma(-l, l, ad);
        l = fma(0.5 * il, res, l);
        il = fma(fma(-l, il, 1.0), il, il);
    }
    il = bad ? 1.0 : il;
    l = bad ? 1.0 : l;
    const double sil = sgnk * il;
    if (ty == ktn) {
        dinv[k] = il;
        sg[k] = sgnk;
        if (bad && state) atomicExch(&state[frame].info, k + 1);
    }
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const double v = e[a][KBN];
        double u = v * sil, keep;
        if (a < KBN) keep = v * il;
        else if (a > KBN) keep = u;
        else {
            keep = (ty > ktn) ? u : v * il;
            if (ty == ktn) {
                u = sil;
                keep = l;
            }
        }
        e[a][KBN] = keep;
        ubuf_next[ty + 16 * a] = u;
    }
}

template <int KB>
__device__ __forceinline__ void potrf_pipe_block(double (&e)[8][8], int ty, int tx, bool p_diag, int& cur,
                                                 double (*ubuf)[kTile], double* dinv, double* sg, ScfState* state,
                                                 int frame, double pivot_tol) {
    constexpr int kb = KB;
    constexpr int bn = (KB < 7) ? KB + 1 : 7;
    for (int kt = 0; kt < 16; ++kt) {
        const int k = kb * 16 + kt;
        const double sk = sg[k];
        const bool p_row = ty <= kt;    // rows of block kb that are <= k
        const bool p_col = tx > kt;     // columns of block kb that are > k
        double su[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) su[a] = sk * ubuf[cur][ty + 16 * a];
        // ---- owners of column k + 1: bring it up to date and start its pivot ----
        bool done_same = false, done_next = false;
        if (kt < 15) {
            if (tx == kt + 1) {
                const double uc = ubuf[cur][tx + 16 * kb];
#pragma unroll
                for (int a = 0; a < 8; ++a)
                    if (POTRF_UPD(a, kb, kb)) e[a][kb] = fma(-su[a], uc, e[a][kb]);
                potrf_next_pivot<kb>(e, ty, tx, kt + 1, ubuf[cur ^ 1], dinv, sg, state, frame, pivot_tol);
                done_same = true;
            }
        } else if (KB < 7) {
            if (tx == 0) {
                const double uc = ubuf[cur][tx + 16 * bn];
#pragma unroll
                for (int a = 0; a < 8; ++a)
                    if (POTRF_UPD(a, bn, kb)) e[a][bn] = fma(-su[a], uc, e[a][bn]);
                potrf_next_pivot<bn>(e, ty, tx, 0, ubuf[cur ^ 1], dinv, sg, state, frame, pivot_tol);
                done_next = true;
            }
        }
        // ---- everybody: the rest of update k ----
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            if (b < kb) continue;                 // columns <= k: finished
            if (b == kb && (!p_col || done_same)) continue;   // same block: column <= k, or column k + 1 already done
            if (b == bn && KB < 7 && done_next) continue;
            const double uc = ubuf[cur][tx + 16 * b];
#pragma unroll
            for (int a = 0; a < 8; ++a)
                if (POTRF_UPD(a, b, kb)) e[a][b] = fma(-su[a], uc, e[a][b]);
        }
        __syncthreads();
        cur ^= 1;
    }
}

__global__ void __launch_bounds__(kPotrfThreads)
potrf_tile_pipe_kernel(double* A_, long lda, long strideA, double* Minv_, long strideM, double* sgn_, long strideSgn,
                       int* neg_cnt_, int* neg_idx_, long strideNeg, ScfState* state, double pivot_tol) {
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
    if (tx == 0) potrf_next_pivot<0>(e, ty, tx, 0, ubuf[0], dinv, sg, state, frame, pivot_tol);
    __syncthreads();
    potrf_pipe_block<0>(e, ty, tx, p_diag, cur, ubuf, dinv, sg, state, frame, pivot_tol);
    potrf_pipe_block<1>(e, ty, tx, p_diag, cur, ubuf, dinv, sg, state, frame, pivot_tol);
    potrf_pipe_block<2>(e, ty, tx, p_diag, cur, ubuf, dinv, sg, state, frame, pivot_tol);
    potrf_pipe_block<3>(e, ty, tx, p_diag, cur, ubuf, dinv, sg, state, frame, pivot_tol);
    potrf_pipe_block<4>(e, ty, tx, p_diag, cur, ubuf, dinv, sg, state, frame, pivot_tol);
    potrf_pipe_block<5>(e, ty, tx, p_diag, cur, ubuf, dinv, sg, state, frame, pivot_tol);
    potrf_pipe_block<6>(e, ty, tx, p_diag, cur, ubuf, dinv, sg, state, frame, pivot_tol);
    potrf_pipe_block<7>(e, ty, tx, p_diag, cur, ubuf, dinv, sg, state, frame, pivot_tol);

    // L back to global; M = S X staged through shared memory (M[i][j] = s_i X[i][j], i > j, held as element (j, i))
#pragma unroll
    for (int b = 0; b < 8; ++b)
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int r = ty + 16 * a, c = tx + 16 * b;
            if (r >= c) A[r + (long)c * lda] = e[a][b];
            double m = 0.0;
            if (r < c) m = sg[c] * e[a][b];
            else if (r == c) m = sg[r] * dinv[r];
            Ms[r * kPotrfLd + c] = m;
        }
    __syncthreads();
    for (int idx = t; idx < kTile * kTile; idx += kPotrfThreads) {
        const int i = idx & (kTile - 1), j = idx >> 7;
        Minv[idx] = Ms[j * kPotrfLd + i];
    }
    if (t < kTile && sgn_) sgn_[(long)frame * strideSgn + t] = sg[t];
    if (t == 0 && neg_cnt_) {   // compact list of this tile's negative pivots for the GEMM epilogues
        int n = 0;
        int* idx = neg_idx_ + (long)frame * strideNeg * kTile;
        for (int i = 0; i < kTile; ++i)
            if (sg[i] < 0.0) idx[n++] = i;
        neg_cnt_[(long)frame * strideNeg] = n;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Diagonal tile with a dedicated CHAIN WARP (experimental, CHEQ_POTRF_CW=1).  Same arithmetic and outputs as potrf_tile_kernel.
//
// In the rank-1 kernel the thread that owns the pivot spends 505 cycles per step on the dependency chain (pivot ->
// rsqrt -> column scaling -> publish) and then 433 cycles on its share of the rank-1 update, while the other seven
// warps wait at the barrier (profiles/r1_potrf_variants.txt).  Here the chain runs in a ninth warp:
//   update warps (8 x 32 threads, the tile in registers as before), step k:
//     the 16 owners of column k+1 first apply update k to that column and publish it raw (colbuf), their warp
//     arrives at a named barrier; then every warp does the rest of update k; __syncthreads.
//   chain warp, step k: waits on the named barrier, turns the raw column k+1 into the scaled column u (ubuf) and
//     the final column values (keepbuf: L below the pivot, X = L^-1 above it), records pivot sign and inverse;
//     __syncthreads.  The owners pick their final column up from keepbuf at the start of the next step.
// Step time: max(chain, update) instead of chain + update.
// ------------------------------------------------------------------------------------------------------------
constexpr int kCwThreads = kPotrfThreads + 32;
constexpr int kCwSoloThreads = 384;

__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__device__ __forceinline__ void cw_chain_step(int kp, int lane, const double* col, double* u_out, double* keep_out,
                                              double* dinv, double* sg, ScfState* state, int frame,
                                              double pivot_tol) {
    double v[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) v[m] = col[lane + 32 * m];
    const double d = col[kp];
    const double ad = fabs(d);
    const bool bad = !(ad > pivot_tol) || !(ad < 1.0e30);
    const double sgnk = (d < 0.0) ? -1.0 : 1.0;
    double il = rsqrt_pivot(ad);
    double l = ad * il;
    {
        const double res = fma(-l, l, ad);
        l = fma(0.5 * il, res, l);
        il = fma(fma(-l, il, 1.0), il, il);
    }
    il = bad ? 1.0 : il;
    l = bad ? 1.0 : l;
    const double sil = sgnk * il;
    if (lane == 0) {
        dinv[kp] = il;
        sg[kp] = sgnk;
        if (bad && state) atomicExch(&state[frame].info, kp + 1);
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int r = lane + 32 * m;
        double u = v[m] * sil, keep = (r > kp) ? u : v[m] * il;
        if (r == kp) {
            u = sil;
            keep = l;
        }
        u_out[r] = u;
        keep_out[r] = keep;
    }
}

template <int KB>
__device__ __forceinline__ void potrf_cw_block(double (&e)[8][8], int ty, int tx, bool p_diag, int& cur,
                                               double (*ubuf)[kTile], double (*colbuf)[kTile],
                                               double (*keepbuf)[kTile], const double* sg) {
    constexpr int kb = KB;
    constexpr int bn = (KB < 7) ? KB + 1 : 7;
    for (int kt = 0; kt < 16; ++kt) {
        const int k = kb * 16 + kt;
        if (tx == kt) {   // my column k is final: take it back from the chain warp
#pragma unroll
            for (int a = 0; a < 8; ++a) e[a][kb] = keepbuf[cur][ty + 16 * a];
        }
        const double sk = sg[k];
        const bool p_row = ty <= kt;    // rows of block kb that are <= k
        const bool p_col = tx > kt;     // columns of block kb that are > k
        double su[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) su[a] = sk * ubuf[cur][ty + 16 * a];
        // ---- owners of column k + 1: bring it up to date, hand it to the chain warp ----
        bool done_same = false, done_next = false;
        if (kt < 15) {
            if (tx == kt + 1) {
                const double uc = ubuf[cur][tx + 16 * kb];
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    if (POTRF_UPD(a, kb, kb)) e[a][kb] = fma(-su[a], uc, e[a][kb]);
                    colbuf[cur ^ 1][ty + 16 * a] = e[a][kb];
                }
                done_same = true;
            }
            if ((tx >> 1) == ((kt + 1) >> 1)) named_bar_arrive(1, 64);   // the warp that contains those owners
        } else if (KB < 7) {
            if (tx == 0) {
                const double uc = ubuf[cur][tx + 16 * bn];
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    if (POTRF_UPD(a, bn, kb)) e[a][bn] = fma(-su[a], uc, e[a][bn]);
                    colbuf[cur ^ 1][ty + 16 * a] = e[a][bn];
                }
                done_next = true;
            }
            if ((tx >> 1) == 0) named_bar_arrive(1, 64);
        }
        // ---- everybody: the rest of update k ----
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            if (b < kb) continue;                 // columns <= k: finished
            if (b == kb && (!p_col || done_same)) continue;
            if (b == bn && KB < 7 && done_next) continue;
            const double uc = ubuf[cur][tx + 16 * b];
#pragma unroll
            for (int a = 0; a < 8; ++a)
                if (POTRF_UPD(a, b, kb)) e[a][b] = fma(-su[a], uc, e[a][b]);
        }
        named_bar_sync(2, kCwThreads);   // update warps + chain warp
        cur ^= 1;
    }
}

__global__ void __launch_bounds__(kCwSoloThreads)
potrf_tile_cw_kernel(double* A_, long lda, long strideA, double* Minv_, long strideM, double* sgn_, long strideSgn,
                     int* neg_cnt_, int* neg_idx_, long strideNeg, ScfState* state, double pivot_tol) {
    const int frame = blockIdx.z;
    if (state && !state[frame].active) return;
    extern __shared__ __align__(16) double Ms[];   // staging of M for a coalesced store
    __shared__ double ubuf[2][kTile], colbuf[2][kTile], keepbuf[2][kTile];
    __shared__ double dinv[kTile];
    __shared__ double sg[kTile];
    double* A = A_ + (long)frame * strideA;
    double* Minv = Minv_ + (long)frame * strideM;
    // Roles.  288 threads: warps 0-7 update, warp 8 runs the chain.  384 threads ("solo"): warps map to SM
    // sub-partitions by warp % 4, so warps 0, 4, 8 share one; warp 0 runs the chain there ALONE (4 and 8 exit at
    // once, as does 11), the eight update warps live on the other three sub-partitions and their FP64 pipes.
    const int pw = threadIdx.x >> 5;
    bool chain;
    int lw;
    if (blockDim.x == kCwThreads) {
        chain = pw == 8;
        lw = pw;
    } else {
        if (pw == 4 || pw == 8 || pw == 11) return;
        chain = pw == 0;
        lw = pw < 4 ? pw - 1 : (pw < 8 ? pw - 2 : pw - 3);
    }
    const int t = lw * 32 + (threadIdx.x & 31);   // logical thread of the 16 x 16 tile grid (update warps)
    const int ty = t & 15, tx = t >> 4;

    if (chain) {
        // ---------------- chain warp ----------------
        const int lane = threadIdx.x & 31;
        int cur = 0;
        named_bar_sync(1, 64);
        cw_chain_step(0, lane, colbuf[0], ubuf[0], keepbuf[0], dinv, sg, state, frame, pivot_tol);
        named_bar_sync(2, kCwThreads);
#ifdef CHEQ_POTRF_INSTRUMENT
        long long w1 = 0, wc = 0, w2 = 0;
#endif
        for (int k = 0; k < kTile - 1; ++k) {
#ifdef CHEQ_POTRF_INSTRUMENT
            const long long c0 = clock64();
#endif
            named_bar_sync(1, 64);
#ifdef CHEQ_POTRF_INSTRUMENT
            const long long c1 = clock64();
#endif
            cw_chain_step(k + 1, lane, colbuf[cur ^ 1], ubuf[cur ^ 1], keepbuf[cur ^ 1], dinv, sg, state, frame,
                          pivot_tol);
#ifdef CHEQ_POTRF_INSTRUMENT
            const long long c2 = clock64();
#endif
            named_bar_sync(2, kCwThreads);
#ifdef CHEQ_POTRF_INSTRUMENT
            w1 += c1 - c0;
            wc += c2 - c1;
            w2 += clock64() - c2;
#endif
            cur ^= 1;
        }
#ifdef CHEQ_POTRF_INSTRUMENT
        if (lane == 0) {
            atomicAdd(&g_potrf_cycles[1], (unsigned long long)w1);   // wait for the raw column
            atomicAdd(&g_potrf_cycles[2], (unsigned long long)wc);   // chain
            atomicAdd(&g_potrf_cycles[3], (unsigned long long)w2);   // wait for the update warps
            atomicAdd(&g_potrf_cycles[5], 1ull);
        }
#endif
        named_bar_sync(2, kCwThreads);   // step 127 of the update warps
    } else {
        // ---------------- update warps ----------------
        double e[8][8];
#pragma unroll
        for (int b = 0; b < 8; ++b)
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int r = ty + 16 * a, c = tx + 16 * b;
                e[a][b] = (r >= c) ? A[r + (long)c * lda] : 0.0;
            }
        const bool p_diag = ty >= tx;
        int cur = 0;
        if (tx == 0) {
#pragma unroll
            for (int a = 0; a < 8; ++a) colbuf[0][ty + 16 * a] = e[a][0];
        }
        if ((tx >> 1) == 0) named_bar_arrive(1, 64);
        named_bar_sync(2, kCwThreads);
        potrf_cw_block<0>(e, ty, tx, p_diag, cur, ubuf, colbuf, keepbuf, sg);
        potrf_cw_block<1>(e, ty, tx, p_diag, cur, ubuf, colbuf, keepbuf, sg);
        potrf_cw_block<2>(e, ty, tx, p_diag, cur, ubuf, colbuf, keepbuf, sg);
        potrf_cw_block<3>(e, ty, tx, p_diag, cur, ubuf, colbuf, keepbuf, sg);
        potrf_cw_block<4>(e, ty, tx, p_diag, cur, ubuf, colbuf, keepbuf, sg);
        potrf_cw_block<5>(e, ty, tx, p_diag, cur, ubuf, colbuf, keepbuf, sg);
        potrf_cw_block<6>(e, ty, tx, p_diag, cur, ubuf, colbuf, keepbuf, sg);
        potrf_cw_block<7>(e, ty, tx, p_diag, cur, ubuf, colbuf, keepbuf, sg);

        // L back to global; M = S X staged through shared memory (M[i][j] = s_i X[i][j], i > j, held as (j, i))
#pragma unroll
        for (int b = 0; b < 8; ++b)
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int r = ty + 16 * a, c = tx + 16 * b;
                if (r >= c) A[r + (long)c * lda] = e[a][b];
                double m = 0.0;
                if (r < c) m = sg[c] * e[a][b];
                else if (r == c) m = sg[r] * dinv[r];
                Ms[r * kPotrfLd + c] = m;
            }
        named_bar_sync(3, kPotrfThreads);   // update warps only
        for (int idx = t; idx < kTile * kTile; idx += kPotrfThreads) {
            const int i = idx & (kTile - 1), j = idx >> 7;
            Minv[idx] = Ms[j * kPotrfLd + i];
        }
        if (t < kTile && sgn_) sgn_[(long)frame * strideSgn + t] = sg[t];
        if (t == 0 && neg_cnt_) {
            int n = 0;
            int* idx = neg_idx_ + (long)frame * strideNeg * kTile;
            for (int i = 0; i < kTile; ++i)
                if (sg[i] < 0.0) idx[n++] = i;
            neg_cnt_[(long)frame * strideNeg] = n;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Diagonal tile, DMMA-blocked variant (experimental, see launch_potrf_tile).  Same mathematics and outputs as potrf_tile_kernel, but the tile
// is held in registers in DMMA accumulator layout (warp w owns tile rows 2w, 2w+1 of the 16 x 16 grid of 8 x 8
// tiles; lane T holds rows T/4, columns 2 (T%4), +1 of each) and eliminated four columns at a time:
//   (a) the owners write the current 128 x 4 panel to shared memory;
//   (c) warp 0 runs the four pivots on it (lane l holds rows l, l+32, l+64, l+96), producing the final panel, the
//       scaled columns u_k and the pivot signs;
//   (e) every warp applies the rank-4 update  e(r, c) -= sum_k s_k u_k[r] u_k[c]  with one DMMA.8x8x4 per 8 x 8
//       tile (A fragment = -s u rows, B fragment = u columns), skipping finished columns and the dead part of X;
//       diagonal tiles keep their not-yet-started X entries by a lane mask, and the one tile that contains the
//       panel itself is updated with predicated scalar FMAs.
// Element semantics as above: (r >= c) holds A -> L, (r < c) holds X[c][r], X = L^-1; update iff c > k and
// (r >= c or r <= k).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884_acc(double (&c)[2], double a, double b) { dmma884(c[0], c[1], a, b); }

template <int TCB, int HALF>
__device__ __forceinline__ void potrf3_step(double (&c)[2][16][2], int warp, int lane, double (*Pn)[kTile],
                                            double (*Pk)[kTile], double (*Ut)[kTile], double* sg4, double* dinv,
                                            double* sg, ScfState* state, int frame, double pivot_tol) {
    constexpr int k0 = 8 * TCB + 4 * HALF;
    const int g = lane >> 2, q = lane & 3;
    const bool owns_panel = (q >> 1) == HALF;     // this lane's column pair of tile column TCB lies in the panel
    const int pj = (q & 1) * 2;                   // panel column of its first element
    if (owns_panel) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int r = 16 * warp + 8 * t + g;
            Pn[pj][r] = c[t][TCB][0];
            Pn[pj + 1][r] = c[t][TCB][1];
        }
    }
    __syncthreads();
    if (warp == 0) {
        double p[4][4];   // [slot][column]: row = lane + 32 slot
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int j = 0; j < 4; ++j) p[s][j] = Pn[j][lane + 32 * s];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            constexpr int dummy = 0;
            (void)dummy;
            const int k = k0 + j;
            const int ks = k >> 5, kl = k & 31;   // compile-time: slot and lane that hold row k
            const double d = __shfl_sync(0xffffffffu, p[ks][j], kl);
            const double ad = fabs(d);
            const bool bad = !(ad > pivot_tol) || !(ad < 1.0e30);
            const double sgnk = (d < 0.0) ? -1.0 : 1.0;
            double il = bad ? 1.0 : rsqrt_pivot(ad);
            double l = bad ? 1.0 : ad * il;
            if (!bad) {
                const double res = fma(-l, l, ad);
                l = fma(0.5 * il, res, l);
                il = fma(fma(-l, il, 1.0), il, il);
            }
            const double sil = sgnk * il;
            if (lane == 0) {
                dinv[k] = il;
                sg[k] = sgnk;
                sg4[j] = sgnk;
                if (bad && state) atomicExch(&state[frame].info, k + 1);
            }
            double u[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int r = lane + 32 * s;
                const double v = p[s][j];
                double uu = v * sil, keep = (r > k) ? uu : v * il;
                if (r == k) {
                    uu = sil;
                    keep = l;
                }
                p[s][j] = keep;
                u[s] = uu;
                Ut[j][r] = uu;
            }
            // the remaining columns of the panel
#pragma unroll
            for (int jj = j + 1; jj < 4; ++jj) {
                const int cc = k0 + jj;
                const double ucj = __shfl_sync(0xffffffffu, u[cc >> 5], cc & 31);
                const double f = -sgnk * ucj;
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const int r = lane + 32 * s;
                    if (r >= cc || r <= k) p[s][jj] = fma(f, u[s], p[s][jj]);
                }
            }
        }
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int j = 0; j < 4; ++j) Pk[j][lane + 32 * s] = p[s][j];
    }
    __syncthreads();
    if (owns_panel) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int r = 16 * warp + 8 * t + g;
            c[t][TCB][0] = Pk[pj][r];
            c[t][TCB][1] = Pk[pj + 1][r];
        }
    }
    // rank-4 update: lane's A element is (row 8 tr + g, pivot k0 + q), its B element (pivot k0 + q, column 8 tc + g)
    const double sq = sg4[q];
    double a[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int tr = 2 * warp + t;
        const int r = 8 * tr + g;
        a[t] = -sq * Ut[q][r];
        if (tr == TCB && r > k0 + q) a[t] = 0.0;   // rows of the panel's tile row below pivot k0+q act as X rows (zero)
    }
#pragma unroll
    for (int tc = TCB; tc < 16; ++tc) {
        if (tc == TCB && HALF == 1) continue;      // no column of this tile column lies beyond the panel
        const int cb = 8 * tc + g;
        const double b = (cb > k0 + 3) ? Ut[q][cb] : 0.0;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int tr = 2 * warp + t;
            if (tr < TCB) {
                dmma884_acc(c[t][tc], a[t], b);                     // X rows above the panel
            } else if (tr == TCB) {
                if (tc > TCB) dmma884_acc(c[t][tc], a[t], b);       // X rows inside the panel's tile row
            } else if (tc < tr) {
                dmma884_acc(c[t][tc], a[t], b);                     // L part
            } else if (tc == tr) {                                  // diagonal tile: L part only (g >= column)
                double d0 = c[t][tc][0], d1 = c[t][tc][1];
                dmma884(d0, d1, a[t], b);
                c[t][tc][0] = (g >= 2 * q) ? d0 : c[t][tc][0];
                c[t][tc][1] = (g >= 2 * q + 1) ? d1 : c[t][tc][1];
            }
        }
    }
    // the tile that contains the panel: columns beyond the panel exist only for the first half
    if (HALF == 0 && warp == TCB / 2) {
        constexpr int t = TCB % 2;
        const int r = 8 * TCB + g;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int cc = 8 * TCB + 2 * q + e;
            if (cc > k0 + 3) {
                double v = c[t][TCB][e];
#pragma unroll
                for (int kp = 0; kp < 4; ++kp)
                    if (r >= cc || r <= k0 + kp) v = fma(-sg4[kp] * Ut[kp][r], Ut[kp][cc], v);
                c[t][TCB][e] = v;
            }
        }
    }
}

template <int TCB>
__device__ __forceinline__ void potrf3_block(double (&c)[2][16][2], int warp, int lane, double (*Pn)[kTile],
                                             double (*Pk)[kTile], double (*Ut)[kTile], double* sg4, double* dinv,
                                             double* sg, ScfState* state, int frame, double pivot_tol) {
    potrf3_step<TCB, 0>(c, warp, lane, Pn, Pk, Ut, sg4, dinv, sg, state, frame, pivot_tol);
    potrf3_step<TCB, 1>(c, warp, lane, Pn, Pk, Ut, sg4, dinv, sg, state, frame, pivot_tol);
}

__global__ void __launch_bounds__(kPotrfThreads)
potrf_tile_dmma_kernel(double* A_, long lda, long strideA, double* Minv_, long strideM, double* sgn_, long strideSgn,
                       int* neg_cnt_, int* neg_idx_, long strideNeg, ScfState* state, double pivot_tol) {
    const int frame = blockIdx.z;
    if (state && !state[frame].active) return;
    extern __shared__ __align__(16) double Ms[];   // staging of M for a coalesced store
    __shared__ double Pn[4][kTile], Pk[4][kTile], Ut[4][kTile];
    __shared__ double dinv[kTile], sg[kTile], sg4[4];
    double* A = A_ + (long)frame * strideA;
    double* Minv = Minv_ + (long)frame * strideM;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int g = lane >> 2, q = lane & 3;

    double c[2][16][2];
#pragma unroll
    for (int tt = 0; tt < 2; ++tt)
#pragma unroll
        for (int tc = 0; tc < 16; ++tc)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int r = 16 * warp + 8 * tt + g, cc = 8 * tc + 2 * q + e;
                c[tt][tc][e] = (r >= cc) ? A[r + (long)cc * lda] : 0.0;
            }

    potrf3_block<0>(c, warp, lane, Pn, Pk, Ut, sg4, dinv, sg, state, frame, pivot_tol);
    potrf3_block<1>(c, warp, lane, Pn, Pk, Ut, sg4, dinv, sg, state, frame, pivot_tol);
    potrf3_block<2>(c, warp, lane, Pn, Pk, Ut, sg4, dinv, sg, state, frame, pivot_tol);
    potrf3_block<3>(c, warp, lane, Pn, Pk, Ut, sg4, dinv, sg, state, frame, pivot_tol);
    potrf3_block<4>(c, warp, lane, Pn, Pk, Ut, sg4, dinv, sg, state, frame, pivot_tol);
    potrf3_block<5>(c, warp, lane, Pn, Pk, Ut, sg4, dinv, sg, state, frame, pivot_tol);
    potrf3_block<6>(c, warp, lane, Pn, Pk, Ut, sg4, dinv, sg, state, frame, pivot_tol);
    potrf3_block<7>(c, warp, lane, Pn, Pk, Ut, sg4, dinv, sg, state, frame, pivot_tol);
    potrf3_block<8>(c, warp, lane, Pn, Pk, Ut, sg4, dinv, sg, state, frame, pivot_tol);
    potrf3_block<9>(c, warp, lane, Pn, Pk, Ut, sg4, dinv, sg, state, frame, pivot_tol);
    potrf3_block<10>(c, warp, lane, Pn, Pk, Ut, sg4, dinv, sg, state, frame, pivot_tol);
    potrf3_block<11>(c, warp, lane, Pn, Pk, Ut, sg4, dinv, sg, state, frame, pivot_tol);
    potrf3_block<12>(c, warp, lane, Pn, Pk, Ut, sg4, dinv, sg, state, frame, pivot_tol);
    potrf3_block<13>(c, warp, lane, Pn, Pk, Ut, sg4, dinv, sg, state, frame, pivot_tol);
    potrf3_block<14>(c, warp, lane, Pn, Pk, Ut, sg4, dinv, sg, state, frame, pivot_tol);
    potrf3_block<15>(c, warp, lane, Pn, Pk, Ut, sg4, dinv, sg, state, frame, pivot_tol);
    __syncthreads();

    // L back to global; M = S X staged through shared memory (Ms[j * ld + i] = M[i][j])
#pragma unroll
    for (int tt = 0; tt < 2; ++tt)
#pragma unroll
        for (int tc = 0; tc < 16; ++tc)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int r = 16 * warp + 8 * tt + g, cc = 8 * tc + 2 * q + e;
                const double v = c[tt][tc][e];
                if (r >= cc) A[r + (long)cc * lda] = v;
                double m = 0.0;
                if (r < cc) m = sg[cc] * v;
                else if (r == cc) m = sg[r] * dinv[r];
                Ms[r * kPotrfLd + cc] = m;
            }
    __syncthreads();
    for (int idx = t; idx < kTile * kTile; idx += kPotrfThreads) {
        const int i = idx & (kTile - 1), j = idx >> 7;
        Minv[idx] = Ms[j * kPotrfLd + i];
    }
    if (t < kTile && sgn_) sgn_[(long)frame * strideSgn + t] = sg[t];
    if (t == 0 && neg_cnt_) {
        int n = 0;
        int* idx = neg_idx_ + (long)frame * strideNeg * kTile;
        for (int i = 0; i < kTile; ++i)
            if (sg[i] < 0.0) idx[n++] = i;
        neg_cnt_[(long)frame * strideNeg] = n;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Diagonal tile, shared-memory blocked variant (experimental, CHEQ_POTRF_BLOCKED=1).  Same outputs as potrf_tile_kernel.
//
// The 128 x 128 tile lives in shared memory, T[c * 129 + r]; as in the rank-1 kernel the lower part (r >= c)
// becomes L and the strict upper part (r < c) holds X[c][r], X = L^-1 (forward elimination of the identity).
// Columns are eliminated in blocks of 16 (block K = columns c0 .. c0 + 15):
//   A  one warp factors the 16 x 16 diagonal block in registers (lane i = row c0 + i, pivots exchanged by shuffle):
//      D = L_D S_D L_D^T, X_D = L_D^-1.  This is the only place with a per-pivot dependency chain
//      (shuffle -> rsqrt -> scale -> update); nothing else in the CTA waits on individual pivots.
//   B  every other row r (L rows below the block, X rows above it) is transformed by the block inverse,
//      n_r = e(r, K) X_D^T; L rows store s_j n_r[j], X rows n_r[j]; W[j][r] = n_r[j] for all rows (for the rows
//      of the block itself W[j][r] = X_D[j][r]).
//   C  rank-16 update of the L-shaped region right of the block: e(r, c) -= sum_j W[j][r] L[c][j] for
//      c > c0 + 15 and (r >= c or r < c0 + 16), in 4 x 4 register tiles (rows strided by 32 so that shared-memory
//      accesses are conflict-free, W rows kept in registers across column tiles).
// ------------------------------------------------------------------------------------------------------------
// One pivot of the 16 x 16 diagonal block held by a warp (lane i = row i, e[j] = column j); recursion on the
// compile-time pivot index keeps e[] in registers.
template <int k>
__device__ __forceinline__ void pb_pivot(double (&e)[16], double& myil, int i, int lane, int c0, double* dinv,
                                         double* sg, ScfState* state, int frame, double pivot_tol) {
    const double d = __shfl_sync(0xffffffffu, e[k], k);
    const double ad = fabs(d);
    const bool bad = !(ad > pivot_tol) || !(ad < 1.0e30);
    const double sgnk = (d < 0.0) ? -1.0 : 1.0;
    double il = rsqrt_pivot(ad);
    double l = ad * il;
    {
        const double res = fma(-l, l, ad);
        l = fma(0.5 * il, res, l);
        il = fma(fma(-l, il, 1.0), il, il);
    }
    il = bad ? 1.0 : il;
    l = bad ? 1.0 : l;
    const double sil = sgnk * il;
    if (lane == k) {
        dinv[c0 + k] = il;
        sg[c0 + k] = sgnk;
        if (bad && state) atomicExch(&state[frame].info, c0 + k + 1);
    }
    if (i == k) myil = il;
    const double v = e[k];
    double u = v * sil, keep = (i > k) ? u : v * il;
    if (i == k) {
        u = sil;
        keep = l;
    }
    e[k] = keep;
    const double su = -sgnk * u;
#pragma unroll
    for (int c = k + 1; c < 16; ++c) {
        const double uc = __shfl_sync(0xffffffffu, u, c);
        if (i >= c || i <= k) e[c] = fma(su, uc, e[c]);
    }
    if constexpr (k + 1 < 16) pb_pivot<k + 1>(e, myil, i, lane, c0, dinv, sg, state, frame, pivot_tol);
}

constexpr int kPB = 16;
constexpr int kPbLd = kTile + 1;
constexpr int kPbSmem = (kTile * kPbLd + kPB * kTile + kPB * kPB + 2 * kTile) * (int)sizeof(double);

__global__ void __launch_bounds__(kPotrfThreads)
potrf_tile_blocked_kernel(double* A_, long lda, long strideA, double* Minv_, long strideM, double* sgn_,
                          long strideSgn, int* neg_cnt_, int* neg_idx_, long strideNeg, ScfState* state,
                          double pivot_tol) {
    const int frame = blockIdx.z;
    if (state && !state[frame].active) return;
    extern __shared__ __align__(16) double psm[];
    double* T = psm;                          // [128][129]
    double* W = T + kTile * kPbLd;            // [16][128]
    double* XD = W + kPB * kTile;             // [16][16]  X_D[j][m], zero above the diagonal
    double* dinv = XD + kPB * kPB;            // [128]
    double* sg = dinv + kTile;                // [128]
    double* A = A_ + (long)frame * strideA;
    double* Minv = Minv_ + (long)frame * strideM;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    long long tk0 = clock64(), tkA = 0, tkB = 0, tkC = 0;

    for (int idx = t; idx < kTile * kTile; idx += kPotrfThreads) {
        const int r = idx & (kTile - 1), c = idx >> 7;
        T[c * kPbLd + r] = (r >= c) ? A[r + (long)c * lda] : 0.0;
    }
    __syncthreads();
    const long long tkLoad = clock64() - tk0;

    for (int K = 0; K < kTile / kPB; ++K) {
        const int c0 = K * kPB;
        long long tk1 = clock64();
        // ---- A: diagonal block, warp 0 ----
        if (warp == 0) {
            const int i = lane & 15;
            double e[kPB];
#pragma unroll
            for (int j = 0; j < kPB; ++j) e[j] = T[(c0 + j) * kPbLd + c0 + i];
            double myil = 1.0;
            pb_pivot<0>(e, myil, i, lane, c0, dinv, sg, state, frame, pivot_tol);
            if (lane < kPB) {
#pragma unroll
                for (int j = 0; j < kPB; ++j) {
                    T[(c0 + j) * kPbLd + c0 + i] = e[j];
                    // X_D[j][i]: strict lower part from the transposed upper entries, diagonal 1 / l
                    const double x = (i < j) ? e[j] : ((i == j) ? myil : 0.0);
                    XD[j * kPB + i] = x;
                    W[j * kTile + c0 + i] = x;
                }
            }
        }
        __syncthreads();
        tkA += clock64() - tk1;
        tk1 = clock64();
        // ---- B: all other rows through the block inverse ----
        if (t < kTile && (t < c0 || t >= c0 + kPB)) {
            const int r = t;
            double v[kPB];
#pragma unroll
            for (int m = 0; m < kPB; ++m) v[m] = T[(c0 + m) * kPbLd + r];
            const bool lrow = r >= c0 + kPB;
#pragma unroll
            for (int j = 0; j < kPB; ++j) {
                double acc = 0.0;
#pragma unroll
                for (int m = 0; m <= j; ++m) acc = fma(v[m], XD[j * kPB + m], acc);
                W[j * kTile + r] = acc;
                T[(c0 + j) * kPbLd + r] = lrow ? sg[c0 + j] * acc : acc;
            }
        }
        __syncthreads();
        tkB += clock64() - tk1;
        tk1 = clock64();
        // ---- C: rank-16 update right of the block ----
        if (c0 + kPB < kTile) {
            const int xrows = c0 + kPB;               // rows below this bound are X rows (always updated)
            double w[4][kPB];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int j = 0; j < kPB; ++j) w[a][j] = W[j * kTile + lane + 32 * a];
            for (int c4 = xrows / 4 + warp; c4 < kTile / 4; c4 += kPotrfThreads / 32) {
                const int cb = 4 * c4;
#pragma unroll 1
                for (int b = 0; b < 4; ++b) {
                    const int c = cb + b;
                    double lc[kPB];
#pragma unroll
                    for (int j = 0; j < kPB; ++j) lc[j] = T[(c0 + j) * kPbLd + c];
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const int rlo = 32 * a;
                        if (rlo >= xrows && rlo + 31 < c) continue;   // dead part of X (warp-uniform)
                        const int r = rlo + lane;
                        if (r >= c || r < xrows) {
                            double acc = T[c * kPbLd + r];
#pragma unroll
                            for (int j = 0; j < kPB; ++j) acc = fma(-w[a][j], lc[j], acc);
                            T[c * kPbLd + r] = acc;
                        }
                    }
                }
            }
        }
        __syncthreads();
        tkC += clock64() - tk1;
    }
    const long long tkS0 = clock64();

    // L back to global (lower part); M = S X with M[i][j] = s_i X[i][j] (i > j), s_i / l_i on the diagonal
    for (int idx = t; idx < kTile * kTile; idx += kPotrfThreads) {
        const int i = idx & (kTile - 1), j = idx >> 7;
        if (i >= j) A[i + (long)j * lda] = T[j * kPbLd + i];
        double m = 0.0;
        if (i > j) m = sg[i] * T[i * kPbLd + j];
        else if (i == j) m = sg[i] * dinv[i];
        Minv[idx] = m;
    }
    if (t < kTile && sgn_) sgn_[(long)frame * strideSgn + t] = sg[t];
    if (t == 0 && neg_cnt_) {
        int n = 0;
        int* idx = neg_idx_ + (long)frame * strideNeg * kTile;
        for (int i = 0; i < kTile; ++i)
            if (sg[i] < 0.0) idx[n++] = i;
        neg_cnt_[(long)frame * strideNeg] = n;
    }
    if (t == 0) {
        atomicAdd(&g_potrf_cycles[0], (unsigned long long)tkLoad);
        atomicAdd(&g_potrf_cycles[1], (unsigned long long)tkA);
        atomicAdd(&g_potrf_cycles[2], (unsigned long long)tkB);
        atomicAdd(&g_potrf_cycles[3], (unsigned long long)tkC);
        atomicAdd(&g_potrf_cycles[4], (unsigned long long)(clock64() - tkS0));
        atomicAdd(&g_potrf_cycles[5], 1ull);
    }
}

#endif
