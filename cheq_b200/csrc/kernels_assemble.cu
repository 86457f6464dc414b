// Assembly kernels: shielded-Coulomb matrix tiles, external potential, right-hand-side rows.
//
// Replaces the rayon pair loop of build_invariant_system (/root/reference/src/solver/implementation.rs:491-536),
// shielding::{gto,sto}::calculate_integral (src/shielding/gto.rs:28-90, sto.rs:27-48) and
// compute_external_potential (implementation.rs:369-447).
//
// Layout: atoms are in the engine's internal order (non-hydrogen block, then hydrogen block, each sorted by
// element type and padded to a multiple of 128 with identity rows), coordinates SoA.  The matrix is
// column-major with leading dimension ld = NP + 128; only tiles on or below the diagonal are produced
// (8 * N (N + 1) / 2 algorithmic bytes): every warp stores 512 contiguous bytes per column.
#include <cuda_runtime.h>

#include "device_types.h"
#include "kernels.h"

namespace cheq {

namespace {

constexpr double kKe = kHartreeToEv * kBohrToAngstrom;  // eV * Angstrom: J_eV = kKe * (1 - F) / R_Angstrom

__device__ __forceinline__ PairTableView rebase(PairTableView v, const char* from, const char* to) {
    auto mv = [&](const void* p) { return (const void*)(to + ((const char*)p - from)); };
    v.rcut2 = (const double*)mv(v.rcut2);
    v.rzero2 = (const double*)mv(v.rzero2);
    v.j0 = (const double*)mv(v.j0);
    v.beta = (const double*)mv(v.beta);
    v.term_begin = (const int*)mv(v.term_begin);
    v.term_count = (const int*)mv(v.term_count);
    v.term_c = (const double*)mv(v.term_c);
    v.coef_begin = (const int*)mv(v.coef_begin);
    v.coef_count = (const int*)mv(v.coef_count);
    v.coef = (const double*)mv(v.coef);
    return v;
}

// 1/sqrt(r2): CUDA's rsqrt (documented max error 1 ulp, not correctly rounded) plus ONE Newton step, after which
// the result is within 0.51 ulp of the exact value -- as good as the reference's correctly rounded sqrt followed by a
// divide (src/solver/implementation.rs:498-502; 0.5 + 0.5 ulp).  Three FMAs per pair; tests/test_gpu_parity.py
// (test_pair_integrals_ulp_bound in tests/test_gpu_headline.py) pins |J_gpu - J_oracle| to a few ulp for both bases.
__device__ __forceinline__ double rinv_refined(double r2) {
    const double y = rsqrt(r2);
    const double h = fma(-r2 * y, y, 1.0);       // 1 - r2 y^2
    return fma(0.5 * y, h, y);
}

// J_eV for one pair of element types at squared distance r2 [Angstrom^2].
template <int BASIS>
__device__ __forceinline__ double pair_value(const PairTableView& tb, int ti, int tj, double r2) {
    const int p = ti * tb.num_types + tj;
    if (r2 < tb.rzero2[p]) return tb.j0[p];
    // 1/R and R from one refined rsqrt instead of sqrt + divide: the far-pair path is then store-bound
    const double rinv = rinv_refined(r2);
    const double R = r2 * rinv;
    if (BASIS == 0) {
        // gto.rs:82-90: erf(beta R) / R
        return kKe * rinv * erf(tb.beta[p] * R);
    } else {
        double F = 0.0;
        if (r2 < tb.rcut2[p]) {
            const int t0 = tb.term_begin[p], t1 = t0 + tb.term_count[p];
            for (int t = t0; t < t1; ++t) {
                const int c0 = tb.coef_begin[t];
                double acc = 0.0;
                for (int k = tb.coef_count[t] - 1; k >= 0; --k) acc = fma(acc, R, tb.coef[c0 + k]);
                F = fma(exp(-tb.term_c[t] * R), acc, F);
            }
        }
        return kKe * rinv * (1.0 - F);
    }
}

// Stages the table blob in shared memory when it fits; returns the view to use.
__device__ __forceinline__ PairTableView stage_tables(const PairTableView& view, const char* blob,
                                                      int blob_bytes, char* smem, int smem_bytes) {
    if (blob_bytes > smem_bytes) return view;
    for (int i = threadIdx.x * 8; i < blob_bytes; i += blockDim.x * 8)
        *(double*)(smem + i) = *(const double*)(blob + i);
    __syncthreads();
    return rebase(view, blob, smem);
}

constexpr int kAsmRows = 128, kAsmCols = 64, kAsmThreads = 256;

template <int BASIS>
__global__ void __launch_bounds__(kAsmThreads)
assemble_lower_kernel(AssembleArgs a, PairTableView view, const char* blob, int blob_bytes) {
    extern __shared__ __align__(16) char smem_raw[];
    const int bi = blockIdx.x, bj = blockIdx.y, frame = blockIdx.z;
    if (bi * kAsmRows + kAsmRows - 1 < bj * kAsmCols) return;  // tile strictly above the diagonal
    if (a.col_mask && !a.col_mask[(bj * kAsmCols) / kTile]) return;   // tile column assembled by another rank
    double* xj = (double*)smem_raw;
    double* yj = xj + kAsmCols;
    double* zj = yj + kAsmCols;
    int* tj = (int*)(zj + kAsmCols);
    char* tab_smem = (char*)(tj + kAsmCols);
    const long pos_off = (long)frame * a.pos_stride;
    const int t = threadIdx.x;
    if (t < kAsmCols) {
        const int j = bj * kAsmCols + t;
        xj[t] = a.x[pos_off + j];
        yj[t] = a.y[pos_off + j];
        zj[t] = a.z[pos_off + j];
        tj[t] = a.type[j];
    }
    PairTableView tb = stage_tables(view, blob, blob_bytes, tab_smem, a.table_smem_bytes);
    __syncthreads();

    const int rp = t & 63, cg = t >> 6;
    const int i0 = bi * kAsmRows + 2 * rp;
    const double xi0 = a.x[pos_off + i0], yi0 = a.y[pos_off + i0], zi0 = a.z[pos_off + i0];
    const double xi1 = a.x[pos_off + i0 + 1], yi1 = a.y[pos_off + i0 + 1], zi1 = a.z[pos_off + i0 + 1];
    const int ti0 = a.type[i0], ti1 = a.type[i0 + 1];
    const double d0 = a.diag[i0], d1 = a.diag[i0 + 1];
    double* out = a.A + (long)frame * a.a_stride;

#pragma unroll 4
    for (int c = 0; c < 16; ++c) {
        const int jl = cg * 16 + c;
        const int j = bj * kAsmCols + jl;
        const int tjj = tj[jl];
        double v0 = 0.0, v1 = 0.0;
        if (tjj != kPadType) {
            const double dx0 = xi0 - xj[jl], dy0 = yi0 - yj[jl], dz0 = zi0 - zj[jl];
            const double dx1 = xi1 - xj[jl], dy1 = yi1 - yj[jl], dz1 = zi1 - zj[jl];
            const double r20 = dx0 * dx0 + dy0 * dy0 + dz0 * dz0, r21 = dx1 * dx1 + dy1 * dy1 + dz1 * dz1;
            // beyond every pair type's cutoff F(R) < 1e-19: J = kKe / R without touching the tables (with the
            // space-filling-curve order most warps take this branch as a whole)
            if (BASIS == 1 && r20 >= a.rcut2_max && r21 >= a.rcut2_max) {
                v0 = (ti0 != kPadType) ? kKe * rinv_refined(r20) : 0.0;
                v1 = (ti1 != kPadType) ? kKe * rinv_refined(r21) : 0.0;
            } else {
                if (ti0 != kPadType) v0 = pair_value<BASIS>(tb, ti0, tjj, r20);
                if (ti1 != kPadType) v1 = pair_value<BASIS>(tb, ti1, tjj, r21);
            }
        }
        if (j == i0) v0 = d0;
        if (j == i0 + 1) v1 = d1;
        *(double2*)(out + (long)i0 + (long)j * a.ld) = make_double2(v0, v1);
    }
}

// Rows NP .. NP+127 of every column: row NP = c_j = -(chi_j + v_j), row NP+1 = 1 (0 for padding), rest 0.
__global__ void rhs_rows_kernel(double* A, long ld, long a_stride, int NP, const double* chi,
                                const double* vext, long vec_stride, const uint8_t* type) {
    const int j = blockIdx.x;
    const int frame = blockIdx.z;
    const int r = threadIdx.x;  // 0..127
    double v = 0.0;
    const bool real = type[j] != kPadType;
    if (real) {
        if (r == 0) v = -(chi[j] + (vext ? vext[(long)frame * vec_stride + j] : 0.0));
        if (r == 1) v = 1.0;
    }
    A[(long)frame * a_stride + (long)NP + r + (long)j * ld] = v;
}

// One warp per atom: v_i = sum_k q_k J_eV(|r_i - r_k|) - E . r_i   (implementation.rs:402-440)
template <int BASIS>
__global__ void __launch_bounds__(256)
vext_kernel(VextArgs a, PairTableView view, const char* blob, int blob_bytes) {
    extern __shared__ __align__(16) char smem_raw[];
    PairTableView tb = stage_tables(view, blob, blob_bytes, smem_raw, a.table_smem_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + warp;
    if (i >= a.NP) return;
    const int ti = a.type[i];
    double v = 0.0;
    if (ti != kPadType) {
        const double xi = a.x[i], yi = a.y[i], zi = a.z[i];
        for (long k = lane; k < a.m; k += 32) {
            const double dx = xi - a.pc_xyz[3 * k], dy = yi - a.pc_xyz[3 * k + 1], dz = zi - a.pc_xyz[3 * k + 2];
            const double jev = pair_value<BASIS>(tb, ti, a.pc_type[k], dx * dx + dy * dy + dz * dz);
            v = fma(a.pc_charge[k], jev, v);
        }
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        v -= a.field[0] * xi + a.field[1] * yi + a.field[2] * zi;
    }
    if (lane == 0) a.v_out[i] = v;
}

// Independent pairs (test entry point): same device function as the assembly.
template <int BASIS>
__global__ void pair_list_kernel(long count, const double* dist, const int* ti, const int* tj,
                                 PairTableView tb, double* out) {
    const long k = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (k >= count) return;
    out[k] = pair_value<BASIS>(tb, ti[k], tj[k], dist[k] * dist[k]);
}

// internal order <- input order (AoS xyz -> SoA, parameters), padding filled with neutral values
__global__ void gather_atoms_kernel(GatherArgs g) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int frame = blockIdx.z;
    if (i >= g.NP) return;
    const int src = g.perm[i];
    const long po = (long)frame * g.pos_stride;
    if (src < 0) {
        g.x[po + i] = 0.0;
        g.y[po + i] = 0.0;
        g.z[po + i] = 0.0;
        if (frame == 0) {
            g.diag[i] = 1.0;
            g.chi[i] = 0.0;
            g.hard[i] = 0.0;
        }
        return;
    }
    const double* p = g.xyz_in + ((long)frame * g.n + src) * 3;
    g.x[po + i] = p[0];
    g.y[po + i] = p[1];
    g.z[po + i] = p[2];
    if (frame == 0) {
        g.diag[i] = g.hard_in[src];
        g.hard[i] = g.hard_in[src];
        g.chi[i] = g.chi_in[src];
    }
}

__global__ void scatter_charges_kernel(int NP, const int* perm, const double* q, long q_stride, double* out,
                                       long n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int frame = blockIdx.z;
    if (i >= NP) return;
    const int dst = perm[i];
    if (dst >= 0) out[(long)frame * n + dst] = q[(long)frame * q_stride + i];
}

// Mirrors the lower triangle into the upper one and scatters to input order (test entry point).
__global__ void export_matrix_kernel(int NP, long ld, const double* A, const int* perm, long n, double* out) {
    const int i = blockIdx.y * blockDim.x + threadIdx.x;
    const int j = blockIdx.x;
    if (i >= NP) return;
    const int oi = perm[i], oj = perm[j];
    if (oi < 0 || oj < 0) return;
    const double v = (i >= j) ? A[(long)i + (long)j * ld] : A[(long)j + (long)i * ld];
    out[(long)oi + (long)oj * n] = v;
}


// ---------------------------------------------------------------------------------------------------------------
// Matrix-free path (large N: the (N + 1)^2 matrix is never formed; SURVEY.md section 8(e) last row, App. E.4).
// y_i = sum_{j != i} J(i, j) v_j for the rows [row0, row0 + 128 gridDim.x) with the J tiles recomputed on the fly
// from the same device function as the assembly (pair_value).  CTA = 128 rows x one slice of the columns
// (gridDim.y slices, partial sums reduced by matfree_reduce_kernel in a fixed order); columns are staged in tiles of
// 256 (coordinates, types, up to two input vectors); thread (rp, cg) owns rows 2 rp, 2 rp + 1 and the 64 columns
// cg * 64 .. + 63 of every staged tile.  FP64-ALU-bound: no matrix traffic at all.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMfRows = 128, kMfCols = 256, kMfThreads = 256;

template <int BASIS, int NV>
__global__ void __launch_bounds__(kMfThreads)
matfree_matvec_kernel(MatfreeArgs a, PairTableView view, const char* blob, int blob_bytes) {
    extern __shared__ __align__(16) char smem_raw[];
    double* xj = (double*)smem_raw;
    double* yj = xj + kMfCols;
    double* zj = yj + kMfCols;
    double* vj = zj + kMfCols;                    // [NV][kMfCols]
    int* tj = (int*)(vj + NV * kMfCols);
    char* tab_smem = (char*)(tj + kMfCols);
    PairTableView tb = stage_tables(view, blob, blob_bytes, tab_smem, a.table_smem_bytes);
    const int t = threadIdx.x;
    const int rp = t & 63, cg = t >> 6;
    const int i0 = a.row0 + blockIdx.x * kMfRows + 2 * rp;
    const double xi0 = a.x[i0], yi0 = a.y[i0], zi0 = a.z[i0];
    const double xi1 = a.x[i0 + 1], yi1 = a.y[i0 + 1], zi1 = a.z[i0 + 1];
    const int ti0 = a.type[i0], ti1 = a.type[i0 + 1];
    double acc0[NV], acc1[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc0[v] = acc1[v] = 0.0;
    // this CTA's slice of the column tiles
    const int tiles = a.NP / kMfCols;
    const int per = (tiles + gridDim.y - 1) / gridDim.y;
    const int tile_begin = blockIdx.y * per, tile_end = min(tiles, tile_begin + per);
    for (int ct = tile_begin; ct < tile_end; ++ct) {
        __syncthreads();
        {
            const int j = ct * kMfCols + t;
            xj[t] = a.x[j];
            yj[t] = a.y[j];
            zj[t] = a.z[j];
            tj[t] = a.type[j];
#pragma unroll
            for (int v = 0; v < NV; ++v) vj[v * kMfCols + t] = a.vin[(long)v * a.vstride + j];
        }
        __syncthreads();
#pragma unroll 4
        for (int c = 0; c < 64; ++c) {
            const int jl = cg * 64 + c;
            const int j = ct * kMfCols + jl;
            const int tjj = tj[jl];
            if (tjj == kPadType) continue;
            const double dx0 = xi0 - xj[jl], dy0 = yi0 - yj[jl], dz0 = zi0 - zj[jl];
            const double dx1 = xi1 - xj[jl], dy1 = yi1 - yj[jl], dz1 = zi1 - zj[jl];
            const double r20 = dx0 * dx0 + dy0 * dy0 + dz0 * dz0, r21 = dx1 * dx1 + dy1 * dy1 + dz1 * dz1;
            double v0 = 0.0, v1 = 0.0;
            if (BASIS == 1 && r20 >= a.rcut2_max && r21 >= a.rcut2_max) {
                v0 = kKe * rinv_refined(r20);
                v1 = kKe * rinv_refined(r21);
            } else {
                if (j != i0 && ti0 != kPadType) v0 = pair_value<BASIS>(tb, ti0, tjj, r20);
                if (j != i0 + 1 && ti1 != kPadType) v1 = pair_value<BASIS>(tb, ti1, tjj, r21);
            }
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const double xv = vj[v * kMfCols + jl];
                acc0[v] = fma(v0, xv, acc0[v]);
                acc1[v] = fma(v1, xv, acc1[v]);
            }
        }
    }
    // reduce the four column groups (fixed order), one partial per (slice, row)
    __syncthreads();
    double* red = (double*)smem_raw;                 // [4][128][NV]
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        red[(cg * kMfRows + 2 * rp) * NV + v] = (ti0 != kPadType) ? acc0[v] : 0.0;
        red[(cg * kMfRows + 2 * rp + 1) * NV + v] = (ti1 != kPadType) ? acc1[v] : 0.0;
    }
    __syncthreads();
    if (t < kMfRows) {
        const long row = (long)blockIdx.x * kMfRows + t;            // row inside this rank's range
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const double s = red[(0 * kMfRows + t) * NV + v] + red[(1 * kMfRows + t) * NV + v] +
                             red[(2 * kMfRows + t) * NV + v] + red[(3 * kMfRows + t) * NV + v];
            a.part[((long)blockIdx.y * NV + v) * a.rows + row] = s;
        }
    }
}

// y_i = sum_slices part + (diag_i + shift_i) v_i + nu e_i   for this rank's rows; e_i = 1 on real atoms, padding
// rows are identity rows.  Written at absolute row positions of `out` (the ranks' pieces are all-gathered after).
__global__ void __launch_bounds__(256)
matfree_reduce_kernel(MatfreeArgs a, int nv, int slices) {
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= a.rows) return;
    const int i = a.row0 + r;
    const bool pad = a.type[i] == kPadType;
    for (int v = 0; v < nv; ++v) {
        double s = 0.0;
        for (int k = 0; k < slices; ++k) s += a.part[((long)k * nv + v) * a.rows + r];
        const double xi = a.vin[(long)v * a.vstride + i];
        const double nu = a.vin[(long)v * a.vstride + a.NP];
        s = pad ? xi : s + (a.diag[i] + (a.shift ? a.shift[i] : 0.0)) * xi + nu;
        a.out[(long)v * a.vstride + i] = s;
    }
}

// border entry of K v: e . v over the real atoms (one CTA per vector, fixed order); optionally out = rhs - K v
__global__ void __launch_bounds__(1024)
matfree_border_kernel(MatfreeArgs a, int NP) {
    __shared__ double red[32];
    const int v = blockIdx.x;
    const double* x = a.vin + (long)v * a.vstride;
    double s = 0.0;
    for (int i = threadIdx.x; i < NP; i += 1024)
        if (a.type[i] != kPadType) s += x[i];
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 32; ++w) tot += red[w];
        a.out[(long)v * a.vstride + NP] = tot;
    }
}

// out = rhs - kv (n + 1 entries), used for true residuals
__global__ void __launch_bounds__(256)
axpby_residual_kernel(const double* rhs, const double* kv, double* out, int len) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < len) out[i] = rhs[i] - kv[i];
}

// 128 x 128 diagonal blocks of A = diag(hardness) + J for the block-Jacobi preconditioner: blocks[b] column-major,
// ld 128, both triangles (the factorisation reads the lower one).  One CTA per block.
template <int BASIS>
__global__ void __launch_bounds__(256)
assemble_diag_blocks_kernel(MatfreeArgs a, PairTableView view, const char* blob, int blob_bytes, double* blocks) {
    extern __shared__ __align__(16) char smem_raw[];
    double* xj = (double*)smem_raw;
    double* yj = xj + kTile;
    double* zj = yj + kTile;
    int* tj = (int*)(zj + kTile);
    char* tab_smem = (char*)(tj + kTile);
    const int b = blockIdx.x, t = threadIdx.x;
    if (t < kTile) {
        const int j = b * kTile + t;
        xj[t] = a.x[j];
        yj[t] = a.y[j];
        zj[t] = a.z[j];
        tj[t] = a.type[j];
    }
    PairTableView tb = stage_tables(view, blob, blob_bytes, tab_smem, a.table_smem_bytes);
    __syncthreads();
    double* out = blocks + (long)b * kTile * kTile;
    for (int idx = t; idx < kTile * kTile; idx += 256) {
        const int r = idx & (kTile - 1), c = idx >> 7;
        const int tr = tj[r], tc = tj[c];
        double v;
        if (r == c) {
            v = (tr == kPadType) ? 1.0 : a.diag[b * kTile + r];
        } else if (tr == kPadType || tc == kPadType) {
            v = 0.0;
        } else {
            const double dx = xj[r] - xj[c], dy = yj[r] - yj[c], dz = zj[r] - zj[c];
            v = pair_value<BASIS>(tb, tr, tc, dx * dx + dy * dy + dz * dz);
        }
        out[idx] = v;
    }
}

// u_b = M_b^T S_b M_b r_b  (= A_b^-1 r_b with A_b = L S L^T, M = S L^-1) for every 128-block: M staged in shared
// memory (padded), two triangular products.
constexpr int kBjLd = kTile + 1;
__global__ void __launch_bounds__(256)
block_jacobi_apply_kernel(const double* Minv, const double* sgn, const double* r, double* u, int nv, long vstride) {
    extern __shared__ __align__(16) double Ms[];    // [128][129]
    __shared__ double tv[2][kTile];
    const int b = blockIdx.x, t = threadIdx.x;
    const double* M = Minv + (long)b * kTile * kTile;
    for (int idx = t; idx < kTile * kTile; idx += 256) {
        const int i = idx & (kTile - 1), j = idx >> 7;            // M[i][j], column-major in global memory
        Ms[i * kBjLd + j] = M[idx];
    }
    __syncthreads();
    for (int v = 0; v < nv; ++v) {
        const double* rb = r + (long)v * vstride + (long)b * kTile;
        if (t < kTile) tv[0][t] = rb[t];
        __syncthreads();
        if (t < kTile) {                       // t_i = s_i sum_{j <= i} M[i][j] r_j
            double s = 0.0;
            for (int j = 0; j <= t; ++j) s = fma(Ms[t * kBjLd + j], tv[0][j], s);
            tv[1][t] = sgn[(long)b * kTile + t] * s;
        }
        __syncthreads();
        if (t < kTile) {                       // u_j = sum_{i >= j} M[i][j] t_i
            double s = 0.0;
            for (int i = t; i < kTile; ++i) s = fma(Ms[i * kBjLd + t], tv[1][i], s);
            u[(long)v * vstride + (long)b * kTile + t] = s;
        }
        __syncthreads();
    }
}

}  // namespace

int assemble_smem_bytes(int table_bytes, int* table_smem_bytes) {
    const int base = 3 * kAsmCols * (int)sizeof(double) + kAsmCols * (int)sizeof(int);
    const int budget = 96 * 1024;
    *table_smem_bytes = (table_bytes <= budget) ? ((table_bytes + 15) / 16) * 16 : 0;
    return base + *table_smem_bytes;
}

cudaError_t launch_assemble(const AssembleArgs& a, const PairTableView& view, const char* blob, int blob_bytes,
                            int batch, cudaStream_t stream) {
    AssembleArgs args = a;
    const int smem = assemble_smem_bytes(blob_bytes, &args.table_smem_bytes);
    dim3 grid(a.NP / kAsmRows, a.NP / kAsmCols, batch);
    auto kern = view.basis == 0 ? assemble_lower_kernel<0> : assemble_lower_kernel<1>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kAsmThreads, smem, stream>>>(args, view, blob, blob_bytes);
    return cudaGetLastError();
}

cudaError_t launch_rhs_rows(double* A, long ld, long a_stride, int NP, const double* chi, const double* vext,
                            long vec_stride, const uint8_t* type, int batch, cudaStream_t stream) {
    dim3 grid(NP, 1, batch);
    rhs_rows_kernel<<<grid, kRhsRows, 0, stream>>>(A, ld, a_stride, NP, chi, vext, vec_stride, type);
    return cudaGetLastError();
}

cudaError_t launch_vext(const VextArgs& a, const PairTableView& view, const char* blob, int blob_bytes,
                        cudaStream_t stream) {
    VextArgs args = a;
    const int budget = 96 * 1024;
    args.table_smem_bytes = (blob_bytes <= budget) ? ((blob_bytes + 15) / 16) * 16 : 0;
    auto kern = view.basis == 0 ? vext_kernel<0> : vext_kernel<1>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         args.table_smem_bytes);
    if (e != cudaSuccess) return e;
    kern<<<(a.NP + 7) / 8, 256, args.table_smem_bytes, stream>>>(args, view, blob, blob_bytes);
    return cudaGetLastError();
}

cudaError_t launch_pair_list(long count, const double* dist, const int* ti, const int* tj,
                             const PairTableView& view, double* out, cudaStream_t stream) {
    const int threads = 128;
    const unsigned blocks = (unsigned)((count + threads - 1) / threads);
    if (view.basis == 0)
        pair_list_kernel<0><<<blocks, threads, 0, stream>>>(count, dist, ti, tj, view, out);
    else
        pair_list_kernel<1><<<blocks, threads, 0, stream>>>(count, dist, ti, tj, view, out);
    return cudaGetLastError();
}

cudaError_t launch_gather(const GatherArgs& g, int batch, cudaStream_t stream) {
    dim3 grid((g.NP + 255) / 256, 1, batch);
    gather_atoms_kernel<<<grid, 256, 0, stream>>>(g);
    return cudaGetLastError();
}

cudaError_t launch_scatter_charges(int NP, const int* perm, const double* q, long q_stride, double* out, long n,
                                   int batch, cudaStream_t stream) {
    dim3 grid((NP + 255) / 256, 1, batch);
    scatter_charges_kernel<<<grid, 256, 0, stream>>>(NP, perm, q, q_stride, out, n);
    return cudaGetLastError();
}

cudaError_t launch_export_matrix(int NP, long ld, const double* A, const int* perm, long n, double* out,
                                 cudaStream_t stream) {
    dim3 grid(NP, (NP + 255) / 256, 1);
    export_matrix_kernel<<<grid, 256, 0, stream>>>(NP, ld, A, perm, n, out);
    return cudaGetLastError();
}


cudaError_t launch_matfree_matvec(const MatfreeArgs& a, int nv, const PairTableView& view, const char* blob,
                                  int blob_bytes, cudaStream_t stream) {
    MatfreeArgs args = a;
    const int budget = 96 * 1024;
    args.table_smem_bytes = (blob_bytes <= budget) ? ((blob_bytes + 15) / 16) * 16 : 0;
    const int stage = kMfCols * (3 + nv) * (int)sizeof(double) + kMfCols * (int)sizeof(int);
    const int red = 4 * kMfRows * nv * (int)sizeof(double);
    const int smem = (stage > red ? stage : red) + args.table_smem_bytes;
    // the table blob sits behind the staging area: make sure the reduction scratch does not reach into it
    const int base = stage > red ? stage : red;
    (void)base;
    dim3 grid(a.rows / kMfRows, a.slices);
    cudaError_t e;
    if (nv == 1) {
        auto kern = view.basis == 0 ? matfree_matvec_kernel<0, 1> : matfree_matvec_kernel<1, 1>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, kMfThreads, smem, stream>>>(args, view, blob, blob_bytes);
    } else {
        auto kern = view.basis == 0 ? matfree_matvec_kernel<0, 2> : matfree_matvec_kernel<1, 2>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, kMfThreads, smem, stream>>>(args, view, blob, blob_bytes);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    matfree_reduce_kernel<<<(a.rows + 255) / 256, 256, 0, stream>>>(args, nv, a.slices);
    return cudaGetLastError();
}

cudaError_t launch_matfree_border(const MatfreeArgs& a, int nv, cudaStream_t stream) {
    matfree_border_kernel<<<nv, 1024, 0, stream>>>(a, a.NP);
    return cudaGetLastError();
}

cudaError_t launch_residual(const double* rhs, const double* kv, double* out, int len, cudaStream_t stream) {
    axpby_residual_kernel<<<(len + 255) / 256, 256, 0, stream>>>(rhs, kv, out, len);
    return cudaGetLastError();
}

cudaError_t launch_assemble_diag_blocks(const MatfreeArgs& a, const PairTableView& view, const char* blob,
                                        int blob_bytes, double* blocks, cudaStream_t stream) {
    MatfreeArgs args = a;
    const int budget = 96 * 1024;
    args.table_smem_bytes = (blob_bytes <= budget) ? ((blob_bytes + 15) / 16) * 16 : 0;
    const int smem = kTile * 3 * (int)sizeof(double) + kTile * (int)sizeof(int) + args.table_smem_bytes;
    auto kern = view.basis == 0 ? assemble_diag_blocks_kernel<0> : assemble_diag_blocks_kernel<1>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    kern<<<a.NP / kTile, 256, smem, stream>>>(args, view, blob, blob_bytes, blocks);
    return cudaGetLastError();
}

cudaError_t launch_block_jacobi_apply(const double* Minv, const double* sgn, const double* r, double* u, int NP,
                                      int nv, long vstride, cudaStream_t stream) {
    const int smem = kTile * kBjLd * (int)sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(block_jacobi_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    block_jacobi_apply_kernel<<<NP / kTile, 256, smem, stream>>>(Minv, sgn, r, u, nv, vstride);
    return cudaGetLastError();
}

}  // namespace cheq
