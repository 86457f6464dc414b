// Kernel launchers (implemented in kernels_assemble.cu and kernels_dense.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "device_types.h"

namespace cheq {

struct AssembleArgs {
    const double *x, *y, *z;   // [batch][NP] (pos_stride apart)
    const uint8_t* type;       // [NP]
    const double* diag;        // [NP] hardness (1 for padding)
    double* A;                 // [batch] column-major, ld
    long ld, a_stride, pos_stride;
    int NP;
    int table_smem_bytes;      // filled by the launcher
    double rcut2_max;          // STO: max over pair types of the squared cutoff [Angstrom^2] (inf disables)
    const uint8_t* col_mask;   // nullable, one flag per 128-column tile: 0 = another rank assembles this tile column
};

struct VextArgs {
    const double *x, *y, *z;
    const uint8_t* type;
    int NP;
    long m;
    const double* pc_xyz;      // 3 m AoS
    const double* pc_charge;
    const int* pc_type;
    double field[3];
    double* v_out;
    int table_smem_bytes;
};

struct GatherArgs {
    int NP;
    long n, pos_stride;
    const int* perm;           // internal -> input index, -1 for padding
    const double *xyz_in, *chi_in, *hard_in;
    double *x, *y, *z, *diag, *chi, *hard;
};

constexpr int kMaxPeers = 7;   // one node: up to 8 GPUs

struct GemmArgs {
    const double* A;           // m x K, column-major
    const double* B;           // n x K, column-major
    double* C;                 // m x n
    long lda, ldb, ldc;
    long strideA, strideB, strideC;  // per frame
    int K;
    int lower;                 // 1: skip tiles strictly above the diagonal of C
    int subtract;              // 1: C -= A S B^T ; 0: C = A S B^T (C may alias A when the grid has one tile column)
    // nullable: negative pivots of the K range, per 128-column tile of K (first tile of the range at index 0):
    // neg_cnt[tile] entries of neg_idx[tile * 128 ...] are the in-tile column indices with pivot sign -1
    const int* neg_cnt;
    const int* neg_idx;
    long strideNeg;            // per frame, in tiles
    int m_tiles, n_tiles;      // filled by the launcher
    const ScfState* state;     // nullable; frames with active == 0 are skipped
    // Column-list mode (distributed / look-ahead updates): tile column t of the grid is tile column n_list[t] of
    // C and of B; B and C then point at absolute tile column 0 and row0 is the absolute row of C's first row
    // (used by the `lower` test).  nullptr: identity.
    const int* n_list;
    int row0;
    // Peer push (multi-GPU panel factorisation): every C tile is ALSO stored at the same offset of each peer's
    // trapezoid (peer-mapped pointers over NVLink), so the finished panel reaches the other ranks tile by tile
    // while the owner is still factoring -- no pack / broadcast / unpack afterwards.
    int n_peers;
    double* peerC[kMaxPeers];   // peer bases corresponding to C
};

// Peer destinations of one diagonal tile's outputs (pointers already offset to the tile)
struct PotrfPeers {
    int n;
    double* A[kMaxPeers];
    double* Minv[kMaxPeers];
    double* sgn[kMaxPeers];
    int* neg_cnt[kMaxPeers];
    int* neg_idx[kMaxPeers];
};

int assemble_smem_bytes(int table_bytes, int* table_smem_bytes);
cudaError_t launch_assemble(const AssembleArgs& a, const PairTableView& view, const char* blob, int blob_bytes,
                            int batch, cudaStream_t stream);
cudaError_t launch_rhs_rows(double* A, long ld, long a_stride, int NP, const double* chi, const double* vext,
                            long vec_stride, const uint8_t* type, int batch, cudaStream_t stream);
cudaError_t launch_vext(const VextArgs& a, const PairTableView& view, const char* blob, int blob_bytes,
                        cudaStream_t stream);
cudaError_t launch_pair_list(long count, const double* dist, const int* ti, const int* tj,
                             const PairTableView& view, double* out, cudaStream_t stream);
cudaError_t launch_gather(const GatherArgs& g, int batch, cudaStream_t stream);
cudaError_t launch_scatter_charges(int NP, const int* perm, const double* q, long q_stride, double* out, long n,
                                   int batch, cudaStream_t stream);
cudaError_t launch_export_matrix(int NP, long ld, const double* A, const int* perm, long n, double* out,
                                 cudaStream_t stream);

// ---- matrix-free path (kernels_assemble.cu) -------------------------------------------------------------------
struct MatfreeArgs {
    const double *x, *y, *z;   // [NP] internal (Morton) order
    const uint8_t* type;       // [NP], 255 = padding
    const double* diag;        // [NP] hardness
    const double* shift;       // [NP] nullable: charge-dependent diagonal shift of the n == 1 atoms
    int NP;                    // padded atom count, multiple of 256
    int row0, rows;            // this rank's row range (multiples of 128)
    int slices;                // column slices per row tile (partials reduced in order)
    const double* vin;         // [nv][vstride] input vectors: NP entries + the border unknown at index NP
    double* out;               // [nv][vstride]
    long vstride;
    double* part;              // [slices][nv][rows]
    double rcut2_max;
    int table_smem_bytes;      // filled by the launcher
};
// out[rows of this rank] = (K v)[rows],  K = [[diag(hardness + shift) + J, e], [e^T, 0]]; nv = 1 or 2
cudaError_t launch_matfree_matvec(const MatfreeArgs& a, int nv, const PairTableView& view, const char* blob,
                                  int blob_bytes, cudaStream_t stream);
cudaError_t launch_matfree_border(const MatfreeArgs& a, int nv, cudaStream_t stream);
cudaError_t launch_residual(const double* rhs, const double* kv, double* out, int len, cudaStream_t stream);
cudaError_t launch_assemble_diag_blocks(const MatfreeArgs& a, const PairTableView& view, const char* blob,
                                        int blob_bytes, double* blocks, cudaStream_t stream);
cudaError_t launch_block_jacobi_apply(const double* Minv, const double* sgn, const double* r, double* u, int NP,
                                      int nv, long vstride, cudaStream_t stream);

// ---- dense (kernels_dense.cu) ------------------------------------------------------------------------------
cudaError_t launch_gemm_nt(const GemmArgs& g, int m_tiles, int n_tiles, int batch, cudaStream_t stream);
// ---- FP64-equivalent GEMM on the tcgen05 tensor cores (kernels_ozaki.cu) ------------------------------------------
// C -= A S B^T with both operands split into `slices` int8 planes per entry (row-scaled, error-free), the slice
// products on tcgen05.mma.kind::i8 with int32 accumulators in TMEM, recombined in fp64.  Single frame, no column
// list, no peers; K a multiple of 128 and at most kOzakiMaxK (int32 group sums); `error`: device flag raised if the
// copy / MMA pipeline times out (C is then left untouched by the remaining CTAs).
constexpr int kOzakiMaxK = 8192;
size_t ozaki_workspace_bytes(int slices, int m_tiles, int n_tiles, int K);
bool ozaki_supported(int slices, int K);
cudaError_t configure_ozaki_kernels();
cudaError_t launch_gemm_ozaki(const GemmArgs& g, int m_tiles, int n_tiles, int slices, void* ws, size_t ws_bytes,
                              int* error, cudaStream_t stream, bool reuse_a = false);
cudaError_t launch_dmma_rate(double* out, int blocks, int threads, int iters, cudaStream_t stream);
cudaError_t launch_latency_probe(double* out, long long* cycles, int iters, cudaStream_t stream);
cudaError_t launch_dfma_rate(double* out, int blocks, int iters, cudaStream_t stream);
void set_gemm_warps(int warps);   // tuning knob: 8 or 16 warps per GEMM CTA
// A = L S L^T of one 128 x 128 diagonal block in place (lower) + M = S L^-1 (dense 128 x 128, ld 128) + the signs.
// |pivot| <= pivot_tol sets state.info (the caller then falls back to the pivoted LU).
cudaError_t launch_potrf_tile(double* A, long lda, long strideA, double* Linv, long strideLinv, double* sgn,
                              long strideSgn, int* neg_cnt, int* neg_idx, long strideNeg, ScfState* state,
                              double pivot_tol, int batch, cudaStream_t stream, const PotrfPeers* peers = nullptr);
// release-store `value` into each peer's signal word (after a system-scope fence)
cudaError_t launch_signal_peers(unsigned* const* peer_words, int n_peers, unsigned value, cudaStream_t stream);
// copies the tiles on/below the diagonal of an (m_tiles x n_tiles) tile region
// col_mask (nullable): one flag per tile column of the region, 0 = skip (columns owned by another rank)
cudaError_t launch_copy_lower_tiles(const double* src, long lds, long stride_src, double* dst, long ldd,
                                    long stride_dst, int m_tiles, int n_tiles, const ScfState* state, int batch,
                                    cudaStream_t stream, const uint8_t* col_mask = nullptr);
// A_ii = S0_ii + hardness_i * kHChargeFactor * clamp(q_i, -0.95, 0.95) for the real atoms of the hydrogen block
cudaError_t launch_set_h_diagonal(double* A, long lda, long strideA, const double* S0, long ld0, long stride0,
                                  const double* hard_h, const double* q_h, long q_stride, const uint8_t* type_h,
                                  int nhp, const ScfState* state, int batch, cudaStream_t stream,
                                  const uint8_t* col_mask = nullptr);
// mu = (f1.fc - Q) / (f1.f1) from rows NP, NP+1; v = fc - mu f1
cudaError_t launch_mu(const double* A, long lda, long strideA, int NP, double total_charge, double* v,
                      long v_stride, const double* sgn, long strideSgn, ScfState* state, int batch,
                      cudaStream_t stream);
// one block-row step (block kb) of the backward substitution L^T x = v (right-looking): x_kb is written to x,
// the right-hand sides of the earlier blocks are updated in v
cudaError_t launch_backsolve_step(const double* A, long lda, long strideA, const double* Linv, long strideLinv,
                                  const double* sgn, long strideSgn, int NP, int kb, double* v, double* x,
                                  long v_stride, const ScfState* state, int batch, cudaStream_t stream);
int backsolve_max_ctas();
// the whole backward substitution as one kernel (chain of CTAs linked by flags); work: batch * (NP / 128 + 1)
// unsigned (tickets + flags, cleared by the launcher); v is not modified
// nb_solve >= 0: only blocks [0, nb_solve) are solved; x of the blocks from nb_solve on is already in place
cudaError_t launch_backsolve_chain(const double* A, long lda, long strideA, const double* Linv, long strideLinv,
                                   const double* sgn, long strideSgn, int NP, const double* v, double* x,
                                   long v_stride, unsigned* work, const ScfState* state, int batch,
                                   cudaStream_t stream, int nb_solve = -1);
// per-device opt-in to large dynamic shared memory for every dense kernel (call after cudaSetDevice)
cudaError_t configure_dense_kernels();
// debug: accumulated cycle counters of the blocked diagonal-tile kernel {load, A, B, C, store, launches, -, -}
cudaError_t potrf_debug_cycles(unsigned long long* out, int reset);
// delta = max |x - q| over real atoms; q <- (1 - d) q + d x
cudaError_t launch_scf_mix(const double* x, double* q, long stride, const uint8_t* type, int NP, ScfState* state,
                           int batch, cudaStream_t stream);
// convergence test + damping schedule (implementation.rs:316-338); damping_kind as in cheq_b200.h
cudaError_t launch_scf_decide(ScfState* state, double tolerance, int damping_kind, int batch, int* active_count,
                              cudaStream_t stream);
cudaError_t launch_scf_init(ScfState* state, double damping, int batch, cudaStream_t stream);

// ---- stale-factor SCF iterations (kernels_krylov.cu) ------------------------------------------------------------
// Device-resident scalars of the reduced bordered hydrogen system (see kernels_krylov.cu)
constexpr int kGmresMaxBasis = 256;   // largest restart length of the GMRES driver (engine.cu)
struct KrylovScalars {
    double g, h;                    // 1_x^T A_xx^-1 1_x,  1_x^T A_xx^-1 c_x
    double den;                     // t1 . (S_s^-1 t1) + g of the frozen factor
    unsigned long long dmax_bits;   // running max |d_new - d_ref| (positive doubles compare as integers)
    double dots[kGmresMaxBasis + 8];   // Gram-Schmidt coefficients of the current step; [kGmresMaxBasis + 1]: w . w
    double hsum[kGmresMaxBasis + 8];   // accumulated coefficients of both Gram-Schmidt passes
    double y[kGmresMaxBasis + 8];      // least-squares solution uploaded by the host
};

struct TileMatvecArgs {
    const void* M;     // nt x nt tiles of 128, column-major; doubles, or floats when f32 != 0
    int f32;
    long ld;
    int nt;
    int upper;         // 0: tiles i >= j hold the data (lower), 1: tiles i <= j (upper)
    int sym;           // symmetric matrix stored as its lower triangle (N and T products in one pass)
    const double* xN;  // N product:  pN[tile rows] = T x[tile columns]
    const double* xT;  // T product:  pT[tile columns] = T^T x[tile rows]
    const double* scaleN;   // nullable: x is scaled entry-wise (pivot signs)
    const double* scaleT;
    double* partN;     // [nt * nt][128] partial results per tile
    double* partT;
    const uint8_t* col_mask;   // nullable: tile columns owned by this rank
    int row_begin, row_end;    // tile rows held by this rank (row_end == 0: all)
    const long long* col_base; // nullable: tile-major storage, element offset of the first stored tile of each column
};

struct TileReduceArgs {
    int nt, upper, op;
    int row_begin, row_end;    // as in TileMatvecArgs
    int border_owner;          // op 1: 1 = this rank contributes the border entry (exactly one rank of a group)
    const double* partN;   // nullable
    const double* partT;   // nullable
    const uint8_t* col_mask;
    const double* v;       // op 1, 2: the vector that was multiplied (n + 1 entries)
    const double* d;       // op 1, 2: diagonal shift D_k
    const double* t1;      // op 1, 2
    const double* rhs;     // op 2
    const KrylovScalars* sc;
    double* out;
};

// All-reduce of a short vector over peer memory (see kernels_krylov.cu).  inbox[d] / flags[d]: rank d's inbox base
// (for the buffer of this round's parity) and flag words as mapped into THIS process; d == rank is local memory.
struct P2PAllreduce {
    int world, rank;
    long slot;                 // doubles per rank slot
    unsigned seq;              // sequence number of this round (the same on every rank)
    double* inbox[kMaxPeers + 1];
    unsigned* flags[kMaxPeers + 1];
};
cudaError_t launch_p2p_allreduce(const double* src, double* out, int len, const P2PAllreduce& pa, unsigned* done_counter,
                                 int* error, cudaStream_t stream);
cudaError_t launch_tile_matvec(const TileMatvecArgs& a, bool do_n, bool do_t, cudaStream_t stream);
cudaError_t launch_tile_matvec_reduce(const TileReduceArgs& a, cudaStream_t stream);
// column-major (src, ld; src at global tile row 0) -> tile-major dst (float if to_f32), stored tiles only
cudaError_t launch_pack_tiles(const double* src, long ld, void* dst, int to_f32, const long long* col_base, int nt,
                              int upper, int row_begin, int row_end, const uint8_t* col_mask, cudaStream_t stream);
cudaError_t launch_multi_dot(const double* V, long stride, const double* w, int len, double* out, int m,
                             bool with_norm, cudaStream_t stream);
cudaError_t launch_gs_update(const double* V, long stride, double* w, int len, const double* h, double* hsum, int m,
                             bool first, cudaStream_t stream);
cudaError_t launch_scale_by_rnorm(const double* src, double* dst, int len, const double* norm2, cudaStream_t stream);
cudaError_t launch_combine(const double* V, long stride, double* x, int len, const double* y, int m,
                           cudaStream_t stream);
cudaError_t launch_precond_border(const double* u0, const double* rin, const double* t1, const double* ws,
                                  const KrylovScalars* sc, double* out, int n, cudaStream_t stream);
cudaError_t launch_precond_den(const double* t1, const double* ws, KrylovScalars* sc, int n, cudaStream_t stream);
// col_mask (nullable): tile columns of the hydrogen block owned by this rank; the entries of t_c, t_1 (and the
// border entry of t_c) that belong to other ranks are written as 0 so that a sum over the ranks completes them
cudaError_t launch_krylov_setup(const double* A, long lda, int NP, int nxp, const double* sgn, const double* S0,
                                long ld0, int nhp, double total_charge, KrylovScalars* sc, double* tc, double* t1,
                                const uint8_t* col_mask, int border_owner, cudaStream_t stream);
cudaError_t launch_h_shift(const double* hard_h, const double* q_h, const uint8_t* type_h, int nhp, double* d_new,
                           const double* d_ref, KrylovScalars* sc, cudaStream_t stream);
cudaError_t launch_set_identity_tiles(double* R, long ld, int nt, int row_begin, int row_end, cudaStream_t stream);
// wx (nullable): W^T x_h, already subtracted from the right-hand side of the non-hydrogen back substitution
cudaError_t launch_krylov_finish(const double* A, long lda, int NP, int nxp, const double* sol, int nhp, double* v,
                                 double* x, ScfState* state, const double* wx, cudaStream_t stream);
// out[nt_cols * 128] = sum over the tile rows [row_begin, row_end) of M(i, j)^T x_i; M: nt_rows x nt_cols tiles, ld
cudaError_t launch_rect_matvec_t(const double* M, long ld, int nt_rows, int nt_cols, int row_begin, int row_end,
                                 const double* x, double* part, double* out, cudaStream_t stream);
cudaError_t launch_krylov_capture(const double* x, int nxp, int nhp, const ScfState* state, double* sol,
                                  cudaStream_t stream);

cudaError_t launch_matfree_rhs(const double* chi, const double* vext, const uint8_t* type, int NP, double total_charge,
                               double* rhs, double* e, cudaStream_t stream);
cudaError_t launch_matfree_shift(const double* hard, const double* q, const uint8_t* hflag, int NP, double* d,
                                 cudaStream_t stream);
cudaError_t launch_matfree_unpack(const double* sol, int NP, double* x, ScfState* state, cudaStream_t stream);

// ---- LU fallback (kernels_lu.cu) ------------------------------------------------------------------------------
int lu_padded_size(int NP);
size_t lu_matrix_bytes(int NP);      // F: (NP + 128) x (NP + 256) doubles
size_t lu_workspace_bytes(int NP);   // Ut: transposed U12 of one outer block
// Builds the bordered matrix from the assembled lower triangle (A is left untouched), factors it with partial
// pivoting (32-column panels inside 128-column outer blocks, the update right of an outer block on the FP64 tensor
// cores), solves, and writes x (internal order) to v and the border unknown to state[0].mu.
cudaError_t launch_lu_solve(const double* A, long ld, int NP, int nxp, int scf, double* F, double* Ut, const double* q,
                            const double* hard, const uint8_t* type, double total_charge, int* ipiv, double* v,
                            ScfState* state, unsigned long long* launches, cudaStream_t stream);

}  // namespace cheq
