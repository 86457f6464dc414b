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

}  // namespace cheq
