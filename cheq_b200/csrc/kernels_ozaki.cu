// FP64-equivalent C -= A S B^T on the 5th-generation tensor cores (tcgen05.mma.kind::i8, accumulators in TMEM).
//
// tcgen05.mma has no f64 kind, so the dense updates of the factorisation (gemm_nt_kernel, DMMA.8x8x4, ~28 TF/s in
// situ) cannot run there as they are.  This file implements the Ozaki error-free split instead:
//
//   * every row of an operand gets one power-of-two scale 2^e (e from the row's largest magnitude), and its entries
//     are written as  x = 2^(e - B) * sum_p d_p 256^(S - p),  B = 8 S - 2,  with S balanced base-256 digits
//     d_p in [-128, 127] (top digit in [-64, 64]) -- S int8 "planes" per operand (oz_slice_kernel);
//   * the slice products are exact in int32:  G_g = sum_{p + q = g} sum_k a_p[m, k] b_q[n, k]  (|G_g| < 2^31 for
//     K <= 16384), and  C[m, n] = 2^(ea_m - 6) 2^(eb_n - 6) sum_{g = 2}^{S + 1} 256^(2 - g) G_g[m, n]; the groups
//     g > S + 1 (below 2^-(8 S - 3) of rowmax * colmax * K) are dropped.  S = 7 carries 54 bits per entry relative
//     to its row maximum, i.e. fp64 grade; S = 4 carries 30 bits (fp32 grade, for preconditioner-only factors);
//   * one CTA owns a 128 x 64 tile of C and keeps ALL S group accumulators resident in TMEM (S x 64 <= 512 columns).
//     For A plane p the B planes q = 1 .. S + 1 - p are consecutive in shared memory, so their products with a_p land
//     in consecutive accumulator groups: ONE tcgen05.mma of N = 64 (S + 1 - p) columns (split at 256) does them all
//     -- 10 instructions per 32-deep K step for S = 7 instead of 28;
//   * the planes are stored in HBM already in the tensor core's canonical K-major core-matrix layout, one contiguous
//     block per (row block, K step), so a stage is filled by two 1-D bulk copies (cp.async.bulk, SASS UBLKCP) that
//     complete on an mbarrier: no tensor maps, no swizzle;
//   * warp 0 / lane 0 is the copy producer, warp 1 / lane 0 issues the MMAs (tcgen05.commit releases the stages),
//     then all four warps read their TMEM lane quarter (tcgen05.ld), recombine the groups by Horner in fp64 and
//     update C.
//
// Every spin has a clock-based timeout that raises *error and lets the grid drain: a broken pipeline can never hang
// the device.
#include <cstdint>
#include <cstdio>

#include "kernels.h"

namespace cheq {

namespace {

constexpr int kOzTM = 128;                 // rows of C per CTA = TMEM lanes
constexpr int kOzTN = 64;                  // columns of C per CTA
constexpr int kOzKB = 32;                  // K bytes per tcgen05.mma.kind::i8 instruction and per pipeline stage
constexpr int kOzPlaneA = kOzTM * kOzKB;   // 4096 B: one plane of one A stage
constexpr int kOzPlaneB = kOzTN * kOzKB;   // 2048 B
constexpr int kOzLbo = 128;                // bytes between the two 16-byte K chunks of a core-matrix row group
constexpr int kOzSbo = 256;                // bytes between consecutive 8-row groups
constexpr int kOzSmemBudget = 220 * 1024;

__host__ __device__ constexpr int oz_stages(int S) {
    return (kOzSmemBudget / (S * (kOzPlaneA + kOzPlaneB))) > 6 ? 6 : (kOzSmemBudget / (S * (kOzPlaneA + kOzPlaneB)));
}

__device__ __forceinline__ uint32_t oz_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor: no swizzle, K-major (same bit layout as cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t oz_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((kOzLbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((kOzSbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// kind::i8 instruction descriptor: s8 x s8 -> s32, both operands K-major
__host__ __device__ constexpr uint32_t oz_idesc(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void oz_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

__device__ __forceinline__ void oz_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}

__device__ __forceinline__ bool oz_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

// bounded wait: false = gave up (another role died or ~8 s of SM clock passed -- long enough to survive time-slicing with
// another process on the same GPU); *dead and *error are raised once
__device__ __forceinline__ bool oz_wait(uint32_t bar, uint32_t parity, volatile int* dead, int* error, int code) {
    if (oz_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!oz_try_wait(bar, parity)) {
        if (*dead) return false;
        if (clock64() - t0 > 16000000000ll) {
            *dead = 1;
            atomicCAS(error, 0, code);
            return false;
        }
    }
    return true;
}

__device__ __forceinline__ void oz_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

struct OzGemmArgs {
    const int8_t* PA;      // [m / 128][nkb][S][4096]
    const int8_t* PB;      // [n / 64][nkb][S][2048]  (pivot signs folded in)
    const double* sa;      // [m]  2^(e - 6) per row of A
    const double* sb;      // [n]
    double* C;
    long ldc;
    int nkb;               // K / 32
    int m_tiles, n_tiles;  // 128-row and 64-column tiles
    int lower, row0, subtract;
    const int* n_list;     // nullable: 128-column tile t of the grid is tile column n_list[t] of C (B is gathered by the slicer)
    const ScfState* state; // nullable
    int* error;            // device flag: != 0 -> the grid drains without touching C
};

template <int S>
__global__ void __launch_bounds__(128, 1) oz_gemm_kernel(const __grid_constant__ OzGemmArgs a) {
    constexpr int NST = oz_stages(S);
    constexpr int kStageA = S * kOzPlaneA, kStageB = S * kOzPlaneB, kStage = kStageA + kStageB;
    constexpr int kGM = 8;   // tile rows per rasterisation group (their A planes stay in L2 for the whole group)
    int tile_m, tile_n;
    {
        const int per_group = kGM * a.n_tiles;
        const int grp = blockIdx.x / per_group, in = blockIdx.x % per_group;
        const int first = grp * kGM;
        const int gsz = min(kGM, a.m_tiles - first);
        tile_m = first + in % gsz;
        tile_n = in / gsz;
    }
    // first column of this tile in C (column-list mode: the owned tile columns of a rank, any ascending set)
    const long ccol0 = (a.n_list ? (long)a.n_list[tile_n >> 1] * kTile : (long)(tile_n >> 1) * kTile) + (tile_n & 1) * kOzTN;
    if (a.lower && a.row0 + tile_m * kOzTM + kOzTM - 1 < ccol0) return;   // entirely above the diagonal
    if (a.state && !a.state[0].active) return;
    if (*(volatile int*)a.error != 0) return;

    extern __shared__ __align__(1024) uint8_t oz_smem[];
    __shared__ __align__(8) unsigned long long full_bar[NST], empty_bar[NST], tmem_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ int dead_s;
    const int t = threadIdx.x, warp = t >> 5;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(oz_smem_u32(&tmem_base_s)),
                     "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (t == 32) {
        for (int i = 0; i < NST; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(oz_smem_u32(&full_bar[i])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(oz_smem_u32(&empty_bar[i])), "r"(1));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(oz_smem_u32(&tmem_bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;\n");
        dead_s = 0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t tmem = tmem_base_s;
    volatile int* dead = &dead_s;
    const uint32_t smem0 = oz_smem_u32(oz_smem);

    // Every thread owns one row of the C tile.  Its 64 old values are fetched NOW, so that their latency hides behind
    // the main loop instead of sitting in the epilogue (which no other CTA on this SM can overlap: TMEM is full).
    const long row = (long)tile_m * kOzTM + t;
    double* crow = a.C + row + ccol0 * a.ldc;
    double cold[kOzTN];
    if (a.subtract) {
#pragma unroll
        for (int j = 0; j < kOzTN; ++j) cold[j] = __ldcg(crow + (long)j * a.ldc);
    } else {
#pragma unroll
        for (int j = 0; j < kOzTN; ++j) cold[j] = 0.0;
    }

    if (t == 0) {
        // ---- producer: two bulk copies per stage ----
        const int8_t* ga = a.PA + (size_t)tile_m * a.nkb * kStageA;
        const int8_t* gb = a.PB + (size_t)tile_n * a.nkb * kStageB;
        for (int kb = 0; kb < a.nkb; ++kb) {
            const int st = kb % NST, use = kb / NST;
            if (use > 0 && !oz_wait(oz_smem_u32(&empty_bar[st]), (uint32_t)((use - 1) & 1), dead, a.error, 1)) break;
            const uint32_t fb = oz_smem_u32(&full_bar[st]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(fb), "r"((uint32_t)kStage) : "memory");
            oz_bulk_load(smem0 + st * kStage, ga + (size_t)kb * kStageA, kStageA, fb);
            oz_bulk_load(smem0 + st * kStage + kStageA, gb + (size_t)kb * kStageB, kStageB, fb);
        }
    } else if (t == 32) {
        // ---- MMA issuer ----
        for (int kb = 0; kb < a.nkb; ++kb) {
            const int st = kb % NST, use = kb / NST;
            if (!oz_wait(oz_smem_u32(&full_bar[st]), (uint32_t)(use & 1), dead, a.error, 2)) break;
            asm volatile("tcgen05.fence::after_thread_sync;\n");
            const uint32_t a0 = smem0 + st * kStage, b0 = a0 + kStageA;
#pragma unroll
            for (int p = 0; p < S; ++p) {
                const uint64_t da = oz_desc(a0 + p * kOzPlaneA);
#pragma unroll
                for (int c0 = p * kOzTN; c0 < S * kOzTN; c0 += 256) {
                    const int N = (S * kOzTN - c0) < 256 ? (S * kOzTN - c0) : 256;
                    const int q0 = (c0 - p * kOzTN) / kOzTN;
                    oz_mma(tmem + (uint32_t)c0, da, oz_desc(b0 + q0 * kOzPlaneB), oz_idesc(kOzTM, N),
                           (kb > 0 || p > 0) ? 1u : 0u);
                }
            }
            oz_commit(oz_smem_u32(&empty_bar[st]));   // the stage is free once these MMAs have read it
        }
        oz_commit(oz_smem_u32(&tmem_bar));            // all MMAs done: accumulators complete
    }
    __syncwarp();

    // ---- epilogue: C (-)= sa sb sum_g 256^(2 - g) G_g ----
    const bool ready = oz_wait(oz_smem_u32(&tmem_bar), 0u, dead, a.error, 3);
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    if (ready && !*dead) {
        const double sar = a.sa[row];
        const double* sb = a.sb + (long)tile_n * kOzTN;
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int c = 0; c < kOzTN; c += 8) {
            uint32_t v[S][8];
#pragma unroll
            for (int g = 0; g < S; ++g)
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                             : "=r"(v[g][0]), "=r"(v[g][1]), "=r"(v[g][2]), "=r"(v[g][3]), "=r"(v[g][4]), "=r"(v[g][5]),
                               "=r"(v[g][6]), "=r"(v[g][7])
                             : "r"(taddr + (uint32_t)(g * kOzTN + c)));
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                double acc = (double)(int)v[S - 1][j];
#pragma unroll
                for (int g = S - 2; g >= 0; --g) acc = fma(acc, 1.0 / 256.0, (double)(int)v[g][j]);
                const double val = acc * sar * sb[c + j];
                crow[(long)(c + j) * a.ldc] = a.subtract ? cold[c + j] - val : val;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512));
}

// ---- slicing ---------------------------------------------------------------------------------------------------
// largest magnitude per row (bit pattern of a non-negative double orders like the integer): grid (rows / 128, K chunks)
// tile_list (nullable): 128-row block b of the output is row tile tile_list[b] of X
__global__ void __launch_bounds__(128) oz_rowmax_kernel(const double* __restrict__ X, long ld, int K, int kchunk,
                                                        unsigned long long* __restrict__ mx,
                                                        const int* __restrict__ tile_list) {
    const long r = (long)blockIdx.x * 128 + threadIdx.x;
    const long rsrc = (tile_list ? (long)tile_list[blockIdx.x] : (long)blockIdx.x) * 128 + threadIdx.x;
    const int k0 = blockIdx.y * kchunk, k1 = min(K, k0 + kchunk);
    const double* p = X + rsrc + (long)k0 * ld;
    // 16 independent loads in flight per thread: the launches are short (K = 512 ... 8192 in chunks of 128 columns)
    // and latency-bound otherwise
    double mm[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) mm[u] = 0.0;
    bool finite = true;   // fmax drops NaN: track it, so that a NaN / Inf entry poisons its row like it would in fp64
    int k = k0;
    for (; k + 16 <= k1; k += 16, p += 16 * ld) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const double ax = fabs(p[(long)u * ld]);
            finite = finite && (ax < INFINITY);
            mm[u] = fmax(mm[u], ax);
        }
    }
    for (; k < k1; ++k, p += ld) {
        const double ax = fabs(p[0]);
        finite = finite && (ax < INFINITY);
        mm[0] = fmax(mm[0], ax);
    }
#pragma unroll
    for (int u = 1; u < 16; ++u) mm[0] = fmax(mm[0], mm[u]);
    const double m0 = finite ? mm[0] : INFINITY;
    atomicMax(mx + r, (unsigned long long)__double_as_longlong(m0));
}

// exponent e with rowmax < 2^e (0 for an all-zero row; 1025 marks a row with a NaN / Inf entry)
__device__ __forceinline__ int oz_exponent(unsigned long long bits) {
    const int E = (int)((bits >> 52) & 0x7ff);
    if (bits == 0ull) return 0;
    return max(E - 1022, -900);
}

// One CTA: RB rows x one 128-wide K tile (4 K steps) -> S planes in the canonical layout
//   byte (r, k) of a plane block = (r / 8) * 256 + (k / 16) * 128 + (r % 8) * 16 + k % 16
// Block (rb, kb) of plane p sits at ((rb * nkb + kb) * S + p) * RB * 32.
template <int S, int RB>
__global__ void __launch_bounds__(256) oz_slice_kernel(const double* __restrict__ X, long ld, int K,
                                                       const unsigned long long* __restrict__ mx,
                                                       double* __restrict__ scale_out, int8_t* __restrict__ planes,
                                                       const int* __restrict__ neg_cnt, const int* __restrict__ neg_idx,
                                                       const int* __restrict__ tile_list) {
    constexpr int B = 8 * S - 2;
    __shared__ unsigned char neg[128];
    const int t = threadIdx.x;
    const int rb = blockIdx.x, kt = blockIdx.y, nkb = K / kOzKB;
    if (t < 128) neg[t] = 0;
    __syncthreads();
    if (neg_cnt) {
        const int cnt = neg_cnt[kt];
        for (int j = t; j < cnt; j += 256) neg[neg_idx[(long)kt * 128 + j]] = 1;
    }
    __syncthreads();
    constexpr int kItems = 8 * RB;   // (K step 0..3) x (16-byte chunk 0..1) x row
#pragma unroll 1
    for (int i = t; i < kItems; i += 256) {
        const int r = i % RB, c16 = i / RB;          // c16: 16-column chunk of the K tile, 0..7
        const long row = (long)rb * RB + r;   // row of the sliced operand; its source row in X differs in list mode
        const long rsrc = tile_list ? (long)tile_list[row >> 7] * 128 + (row & 127) : row;
        const int e = oz_exponent(mx[row]);
        // a row with a non-finite entry gets a NaN scale: every entry of C it touches becomes NaN, as in the fp64 path
        if (kt == 0 && c16 == 0) scale_out[row] = (e == 1025) ? __longlong_as_double(0x7ff8000000000000ll) : ldexp(1.0, e - 6);
        const double sc = ldexp(1.0, B - e);
        const double* src = X + rsrc + ((long)kt * 128 + c16 * 16) * ld;
        long long v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            double x = src[(long)j * ld];
            if (neg[c16 * 16 + j]) x = -x;
            v[j] = __double2ll_rn(x * sc);
        }
        const int kb = kt * 4 + c16 / 2;
        int8_t* dst = planes + ((size_t)((size_t)rb * nkb + kb) * S) * (RB * kOzKB) + (r / 8) * kOzSbo +
                      (c16 % 2) * kOzLbo + (r % 8) * 16;
#pragma unroll
        for (int p = S - 1; p >= 0; --p) {
            uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                long long d;
                if (p > 0) {
                    d = ((v[j] + 128) & 255) - 128;      // balanced digit in [-128, 127]
                    v[j] = (v[j] - d) >> 8;              // exact: v - d is a multiple of 256
                } else {
                    d = v[j];                            // what is left: |d| <= 64
                }
                w[j / 4] |= ((uint32_t)(d & 255)) << (8 * (j % 4));
            }
            *(uint4*)(dst + (size_t)p * (RB * kOzKB)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

// workspace: [row maxima of A | scales of A | planes of A] [row maxima of B | scales of B | planes of B]; the A part
// depends on (m, K, S) only, so a following call with the same A can reuse it
struct OzLayout {
    size_t mxa, sa, pa, mxb, sb, pb, bytes;
    OzLayout(int S, size_t m, size_t n, size_t K) {
        auto up = [](size_t v) { return (v + 1023) / 1024 * 1024; };
        mxa = 0;
        sa = mxa + m * 8;
        pa = up(sa + m * 8);
        mxb = up(pa + (size_t)S * m * K);
        sb = mxb + n * 8;
        pb = up(sb + n * 8);
        bytes = pb + (size_t)S * n * K;
    }
};

template <int S>
cudaError_t oz_launch_t(const GemmArgs& g, int m_tiles, int n_tiles, void* ws, size_t ws_bytes, int* error,
                        bool reuse_a, cudaStream_t stream) {
    const long m = (long)m_tiles * kTile, n = (long)n_tiles * kTile;
    const int K = g.K, nkb = K / kOzKB;
    const OzLayout L(S, (size_t)m, (size_t)n, (size_t)K);
    if (L.bytes > ws_bytes) return cudaErrorInvalidValue;
    char* w = (char*)ws;
    unsigned long long* mxa = (unsigned long long*)(w + L.mxa);
    unsigned long long* mxb = (unsigned long long*)(w + L.mxb);
    double* sa = (double*)(w + L.sa);
    double* sb = (double*)(w + L.sb);
    int8_t* PA = (int8_t*)(w + L.pa);
    int8_t* PB = (int8_t*)(w + L.pb);
    const int kchunk = 128, ksplit = (K + kchunk - 1) / kchunk;
    cudaError_t e;
    if (!reuse_a) {
        e = cudaMemsetAsync(mxa, 0, (size_t)m * 8, stream);
        if (e != cudaSuccess) return e;
        oz_rowmax_kernel<<<dim3((unsigned)(m / 128), ksplit), 128, 0, stream>>>(g.A, g.lda, K, kchunk, mxa, nullptr);
        oz_slice_kernel<S, kOzTM><<<dim3((unsigned)(m / kOzTM), K / 128), 256, 0, stream>>>(g.A, g.lda, K, mxa, sa, PA,
                                                                                           nullptr, nullptr, nullptr);
    }
    e = cudaMemsetAsync(mxb, 0, (size_t)n * 8, stream);
    if (e != cudaSuccess) return e;
    oz_rowmax_kernel<<<dim3((unsigned)(n / 128), ksplit), 128, 0, stream>>>(g.B, g.ldb, K, kchunk, mxb, g.n_list);
    oz_slice_kernel<S, kOzTN><<<dim3((unsigned)(n / kOzTN), K / 128), 256, 0, stream>>>(g.B, g.ldb, K, mxb, sb, PB,
                                                                                       g.neg_cnt, g.neg_idx, g.n_list);
    OzGemmArgs a{};
    a.PA = PA;
    a.PB = PB;
    a.sa = sa;
    a.sb = sb;
    a.C = g.C;
    a.ldc = g.ldc;
    a.nkb = nkb;
    a.m_tiles = (int)(m / kOzTM);
    a.n_tiles = (int)(n / kOzTN);
    a.lower = g.lower;
    a.row0 = g.row0;
    a.subtract = g.subtract;
    a.n_list = g.n_list;
    a.state = g.state;
    a.error = error;
    constexpr int smem = oz_stages(S) * S * (kOzPlaneA + kOzPlaneB);
    oz_gemm_kernel<S><<<(unsigned)(a.m_tiles * a.n_tiles), 128, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace

size_t ozaki_workspace_bytes(int slices, int m_tiles, int n_tiles, int K) {
    return OzLayout(slices, (size_t)m_tiles * kTile, (size_t)n_tiles * kTile, (size_t)K).bytes;
}

bool ozaki_supported(int slices, int K) {
    // int32 group sums: at most `slices` products of magnitude <= 2^14 per k
    return (slices >= 4 && slices <= 8) && K % 128 == 0 && K >= 128 && (long)K * slices * 16384 < 2147483647l;
}

cudaError_t configure_ozaki_kernels() {
    cudaError_t e = cudaSuccess;
#define OZ_CFG(S)                                                                                          \
    if (e == cudaSuccess)                                                                                  \
        e = cudaFuncSetAttribute(oz_gemm_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                 oz_stages(S) * S * (kOzPlaneA + kOzPlaneB));
    OZ_CFG(4) OZ_CFG(5) OZ_CFG(6) OZ_CFG(7) OZ_CFG(8)
#undef OZ_CFG
    return e;
}

// C (-)= A S B^T through int8 slices; single frame, no peers.  Column-list mode as in gemm_nt_kernel: tile column t of
// the grid is tile column n_list[t] of C and of B (both then point at absolute tile column 0).  `error`: device int, raised by a
// pipeline timeout (the caller checks it at its next synchronisation point).
// reuse_a: the previous call on this workspace sliced the same A (pointer, rows, K, slices): keep its planes.
cudaError_t launch_gemm_ozaki(const GemmArgs& g, int m_tiles, int n_tiles, int slices, void* ws, size_t ws_bytes,
                              int* error, cudaStream_t stream, bool reuse_a) {
    if (!ozaki_supported(slices, g.K) || g.n_peers) return cudaErrorInvalidValue;
    switch (slices) {
        case 4: return oz_launch_t<4>(g, m_tiles, n_tiles, ws, ws_bytes, error, reuse_a, stream);
        case 5: return oz_launch_t<5>(g, m_tiles, n_tiles, ws, ws_bytes, error, reuse_a, stream);
        case 6: return oz_launch_t<6>(g, m_tiles, n_tiles, ws, ws_bytes, error, reuse_a, stream);
        case 7: return oz_launch_t<7>(g, m_tiles, n_tiles, ws, ws_bytes, error, reuse_a, stream);
        case 8: return oz_launch_t<8>(g, m_tiles, n_tiles, ws, ws_bytes, error, reuse_a, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace cheq
