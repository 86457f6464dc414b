/*
 * cheq_b200.h -- C ABI of the B200-native QEq engine (drop-in for the hot path of caltechmsc/cheq).
 *
 * The reference is a pure-Rust crate with no FFI layer; its boundary is the Rust public API
 * (/root/reference/src/lib.rs:98-104).  This header is what a `cheq-b200-sys` crate binds so that
 * `QEqSolver::solve` / `solve_in_field` (src/solver/implementation.rs:174-188, 239-266) can hand the
 * whole numerical path -- `build_invariant_system` (:454-554), `compute_external_potential` (:369-447),
 * `run_scf_iterations` (:273-345) and `run_single_solve` (:557-603) -- to the GPU.  The seam sits after
 * `fetch_element_data` (:348-362): the caller resolves atomic numbers to per-atom parameters
 * (ParameterNotFound stays on the host side) and passes plain arrays.
 *
 * Conventions
 *   - All pointers are caller-owned.  `cheq_solve*` takes HOST pointers; `cheq_solve_device` takes DEVICE
 *     pointers for the per-atom arrays (inputs already resident in HBM) and host pointers for scalars.
 *   - Positions and radii in Angstrom, energies in eV, charges in e (cheq's units, constants.rs:13,22).
 *   - Every function returns a status code and never unwinds.  `cheq_last_error` gives the text.
 *   - A context owns one CUDA stream and a grow-only device workspace; calls on one context are serialised
 *     by an internal mutex, distinct contexts are independent (QEqSolver is Send + Sync in the reference).
 *   - There is no CPU fallback: without a CUDA device `cheq_ctx_create` fails with CHEQ_ERR_CUDA.
 */
#ifndef CHEQ_B200_H
#define CHEQ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes.  1..4 mirror CheqError (src/error.rs:18-68). */
#define CHEQ_OK 0
#define CHEQ_ERR_NO_ATOMS 1            /* CheqError::NoAtoms           (implementation.rs:180-182, 250-252) */
#define CHEQ_ERR_PARAMETER_NOT_FOUND 2 /* CheqError::ParameterNotFound (raised by the host shim, :348-362, :378-387) */
#define CHEQ_ERR_NOT_CONVERGED 3       /* CheqError::NotConverged      (implementation.rs:341-344) */
#define CHEQ_ERR_LINALG 4              /* CheqError::LinalgError       (implementation.rs:578-585) */
#define CHEQ_ERR_CUDA 5                /* CUDA runtime failure (no reference equivalent) */
#define CHEQ_ERR_INVALID_ARGUMENT 6
#define CHEQ_ERR_OUT_OF_MEMORY 7

#define CHEQ_BASIS_GTO 0 /* BasisType::Gto (options.rs:9-22) */
#define CHEQ_BASIS_STO 1 /* BasisType::Sto, the default */

#define CHEQ_DAMPING_NONE 0  /* DampingStrategy::None        (options.rs:26-45) */
#define CHEQ_DAMPING_FIXED 1 /* DampingStrategy::Fixed(d)  */
#define CHEQ_DAMPING_AUTO 2  /* DampingStrategy::Auto { initial } */

typedef struct cheq_ctx cheq_ctx;

/* SolverOptions (options.rs:53-100), passed by value semantics. */
typedef struct cheq_options {
    double tolerance;        /* max-abs charge change accepted as converged; default 1e-6 */
    double lambda_scale;     /* Slater exponent scale lambda; default 0.5 */
    double damping_value;    /* Fixed(d) or Auto{initial}; ignored for None; default 0.4 */
    uint32_t max_iterations; /* default 2000 (the reference CLI uses 100) */
    uint8_t hydrogen_scf;    /* default 1 */
    uint8_t basis;           /* CHEQ_BASIS_*; default STO */
    uint8_t damping_kind;    /* CHEQ_DAMPING_*; default AUTO */
    uint8_t reserved;
} cheq_options;

/* ExternalPotential (types.rs:84-97, 136-148) with point-charge elements already resolved. */
typedef struct cheq_external {
    uint64_t num_point_charges;
    const double* xyz;     /* 3 * m, AoS */
    const double* charge;  /* m */
    const double* radius;  /* m, covalent radius of each point charge's element */
    const uint8_t* n;      /* m, principal quantum number of each point charge's element */
    double field[3];       /* uniform field, V/Angstrom */
} cheq_external;

/* Per-call measurements (device times from CUDA events on the context's stream). */
typedef struct cheq_stats {
    double ms_total;         /* whole call, host clock */
    double ms_host_prepare;  /* typing, ordering, table construction */
    double ms_h2d;           /* input upload (0 for cheq_solve_device) */
    double ms_assemble;      /* v_ext + matrix assembly kernels */
    double ms_factor_once;   /* geometry-invariant part of the factorisation */
    double ms_scf;           /* all SCF iterations (refactorisation of the hydrogen block + solves) */
    double ms_d2h;           /* result download */
    uint64_t kernel_launches;
    uint64_t n_atoms, n_hydrogen, n_padded;
    uint32_t iterations;
    uint32_t pair_types;
    double flops_factor;     /* LAPACK-count flops of the factorisation work actually performed */
    double bytes_assemble;   /* algorithmic bytes of the assembly: 8 * N (N + 1) / 2 */
    uint64_t lu_fallback_frames; /* frames re-solved with the pivoted-LU path (unpivoted factorisation broke down) */
    /* filled only when cheq_set_profiling(ctx, 1): every launch of the tensor-core GEMM bracketed by CUDA events */
    double gemm_ms;          /* sum of the launch durations */
    double gemm_flops;       /* algorithmic flops of those launches (2 m n k; n^2 k for the symmetric part) */
    uint64_t gemm_launches;
    double potrf_ms;         /* same, diagonal-tile factorisation launches */
    double backsolve_ms;     /* same, whole back substitutions */
    /* stale-factor SCF iterations: iterations solved by preconditioned GMRES on a frozen factor instead of a
     * refactorisation of the hydrogen block, and the operator applications (one pass over S0 + two over R) they took */
    uint64_t krylov_iterations;
    uint64_t krylov_steps;
    double krylov_ms;        /* device time of those iterations (CUDA events), incl. building R = L^-T S */
    double krylov_bytes;     /* algorithmic bytes streamed by the tiled matrix-vector kernels in them */
    /* matrix-free path (cheq_solve_iterative): operator applications and the pair integrals they recomputed */
    uint64_t matfree_matvecs;
    double matfree_pairs;
    /* updates executed on the tcgen05 tensor cores as int8 slice products (FP64-equivalent, see cheq_set_ozaki) */
    uint64_t tc_gemm_launches;
    double tc_gemm_flops;    /* algorithmic FP64 flops of those launches */
    double tc_gemm_ops;      /* int8 tensor operations they stand for: flops x slice products per entry, S (S + 1) / 2 */
    double tc_gemm_ms;       /* their summed duration incl. the slicing passes (only with cheq_set_profiling) */
} cheq_stats;

int32_t cheq_ctx_create(int32_t device, cheq_ctx** out);
void cheq_ctx_destroy(cheq_ctx* ctx);
const char* cheq_last_error(const cheq_ctx* ctx);
void cheq_options_default(cheq_options* opt);
int32_t cheq_get_stats(const cheq_ctx* ctx, cheq_stats* out);
/* The context's CUDA stream as a cudaStream_t cast to void* (for callers that time with events). */
void* cheq_ctx_stream(const cheq_ctx* ctx);
/* Per-launch CUDA-event timing of the dominant kernel (off by default; adds two event records per launch). */
int32_t cheq_set_profiling(cheq_ctx* ctx, int32_t enabled);

/* SCF trace of the last single-frame solve on this context: per iteration the max-abs charge change
 * (implementation.rs:589-594), how the linear system was solved (0: hydrogen block refactorised, 1: frozen factor +
 * GMRES) and the GMRES operator applications.  Arrays of `capacity` entries (each nullable); returns the number of
 * iterations recorded in *count_out.  Used by the parity tests to compare the whole trajectory with the CPU path. */
int32_t cheq_get_scf_trace(const cheq_ctx* ctx, uint32_t capacity, double* delta_out, int32_t* mode_out,
                           int32_t* steps_out, uint32_t* count_out);
/* Stale-factor policy.  enabled = 0 refactors the hydrogen block in every iteration (the
 * round-1 behaviour); min_hydrogen: smallest hydrogen block that uses the frozen factor; freeze_ev: freeze once the
 * largest change of the hydrogen diagonal between two iterations falls below this many eV; force_iteration: freeze at
 * this SCF iteration at the latest.  enabled < 0, min_hydrogen <= 0, freeze_ev == 0 and force_iteration <= 0 keep the
 * current setting; freeze_ev < 0 selects the built-in cost model (the default: 3.5 .. 12 eV depending on the size of the
 * hydrogen block and the number of GPUs). */
int32_t cheq_set_krylov(cheq_ctx* ctx, int32_t enabled, int32_t min_hydrogen, double freeze_ev, int32_t force_iteration);
/* FP64-equivalent updates on the tcgen05 tensor cores.  tcgen05.mma has no f64 kind; deep
 * single-frame updates C -= A S B^T of the factorisation are therefore run as exact int8 slice products (Ozaki
 * split: `slices` base-256 digits per entry relative to the largest magnitude of its row, int32 accumulators in TMEM,
 * fp64 recombination).  slices: 0 = off (every update on the FP64 DMMA path), 4..8 (7, the default, carries 54 bits);
 * slices_r: the count used while building the frozen factor's R = L^-T S, which only preconditions GMRES (0 = same as
 * slices; default 5); min_k: updates shallower than this stay on DMMA (default 512).  Negative / zero arguments keep
 * the current setting (slices = 0 switches the path off).  Env: CHEQ_OZAKI, CHEQ_OZAKI_R, CHEQ_OZAKI_MIN_K. */
int32_t cheq_set_ozaki(cheq_ctx* ctx, int32_t slices, int32_t slices_r, int32_t min_k);

/* ---- multi-GPU: one dense system across the GPUs of a node ----------------------------
 * The reference is single-node CPU code (rayon, implementation.rs:394, 491) and has no equivalent; this is the
 * "distributed factorisation with NCCL panel broadcasts" row of the scope table.  One process per GPU.  Rank 0
 * obtains an id with cheq_dist_unique_id, the host runtime hands the 128 bytes to every rank (torch.distributed,
 * MPI, a file ...), and every rank calls cheq_ctx_dist_init on its own context.  From then on cheq_solve /
 * cheq_solve_device on that context are COLLECTIVE for single-frame calls: all ranks must call them with
 * identical inputs; the matrix columns are assembled and updated by their owners (1-D block-cyclic panels), each
 * factored panel is broadcast over NVLink, and every rank returns the same charges.  cheq_solve_batch stays
 * local (frames are independent: shard them across ranks instead). */
#define CHEQ_DIST_ID_BYTES 128
int32_t cheq_dist_unique_id(uint8_t* id_out /* CHEQ_DIST_ID_BYTES */);
int32_t cheq_ctx_dist_init(cheq_ctx* ctx, int32_t rank, int32_t world, const uint8_t* id);
int32_t cheq_ctx_dist_info(const cheq_ctx* ctx, int32_t* rank_out, int32_t* world_out);
/* Factorisation schedule: mode 0 = automatic (right-looking panels with look-ahead when the context is part of a
 * group, the recursive tall-panel sequence on one GPU), 1 = always panels, 2 = never panels on one GPU.
 * tiles_per_panel: panel width in 128-column tiles (0 keeps the current value; default 4). */
int32_t cheq_set_factor_mode(cheq_ctx* ctx, int32_t mode, int32_t tiles_per_panel);
/* Host-only: the owner rank and panel index of every 128-column tile for a layout with x_tiles non-hydrogen and
 * h_tiles hydrogen tile columns (arrays of x_tiles + h_tiles entries, either may be NULL). */
int32_t cheq_dist_plan(int32_t x_tiles, int32_t h_tiles, int32_t tiles_per_panel, int32_t world,
                       int32_t* tile_owner_out, int32_t* tile_panel_out);

/*
 * QEqSolver::solve / solve_in_field after element lookup.  `external` may be NULL or empty
 * (types.rs:281-286) for a vacuum solve.  On CHEQ_OK and on CHEQ_ERR_NOT_CONVERGED the outputs hold the
 * last iterate (the reference returns only max_iterations and delta on NotConverged).
 *   charges_out[n]      CalculationResult::charges, input order
 *   potential_out       CalculationResult::equilibrated_potential (the bordered system's last unknown)
 *   iterations_out      CalculationResult::iterations (1 when the SCF is skipped)
 *   delta_out           last max-abs charge change (NotConverged::delta)
 */
int32_t cheq_solve(cheq_ctx* ctx, uint64_t n_atoms, const double* xyz, const double* chi,
                   const double* hardness, const double* radius, const uint8_t* n_quantum,
                   double total_charge, const cheq_options* options, const cheq_external* external,
                   double* charges_out, double* potential_out, uint32_t* iterations_out,
                   double* delta_out);

/*
 * The same solve WITHOUT ever forming the (N + 1)^2 matrix (the reference keeps three copies of it,
 * implementation.rs:464, 282, 575): restarted GMRES on the bordered system in every SCF iteration, the shielded-
 * Coulomb tiles recomputed on the fly in each operator application, block-Jacobi preconditioner over 128-atom
 * clusters.  Memory O(N) per GPU; on a joined context the rows of the operator are sharded over the ranks (one
 * ncclAllGather per application).  cheq_solve switches to this path by itself when the dense matrix cannot fit the
 * device; this entry point forces it (parity tests, large-N runs).  Same arguments, outputs and status codes as
 * cheq_solve; CHEQ_ERR_LINALG if GMRES does not reach its tolerance.
 */
int32_t cheq_solve_iterative(cheq_ctx* ctx, uint64_t n_atoms, const double* xyz, const double* chi,
                             const double* hardness, const double* radius, const uint8_t* n_quantum,
                             double total_charge, const cheq_options* options, const cheq_external* external,
                             double* charges_out, double* potential_out, uint32_t* iterations_out,
                             double* delta_out);

/* Same, with xyz/chi/hardness/radius/n_quantum/charges_out (and the external arrays) in DEVICE memory. */
int32_t cheq_solve_device(cheq_ctx* ctx, uint64_t n_atoms, const double* d_xyz, const double* d_chi,
                          const double* d_hardness, const double* d_radius, const uint8_t* d_n_quantum,
                          double total_charge, const cheq_options* options, const cheq_external* d_external,
                          double* d_charges_out, double* potential_out, uint32_t* iterations_out,
                          double* delta_out);

/*
 * Batched frames of ONE topology (an MD trajectory): per-atom parameters are shared, coordinates differ.
 *   xyz            n_frames * n_atoms * 3
 *   charges_out    n_frames * n_atoms
 *   potential_out, iterations_out, delta_out, status_out   n_frames each (status per frame, CHEQ_* codes)
 * Returns CHEQ_OK if the batch ran; per-frame convergence failures are reported in status_out.
 */
int32_t cheq_solve_batch(cheq_ctx* ctx, uint64_t n_frames, uint64_t n_atoms, const double* xyz,
                         const double* chi, const double* hardness, const double* radius,
                         const uint8_t* n_quantum, double total_charge, const cheq_options* options,
                         double* charges_out, double* potential_out, uint32_t* iterations_out,
                         double* delta_out, int32_t* status_out);

/* ---- component entry points (used by the parity tests; same kernels as cheq_solve) -------------- */

/* shielding::{sto,gto}::calculate_integral (sto.rs:27-39, gto.rs:28-43) * HARTREE_TO_EV for `count` independent
 * pairs: distance in Angstrom, radii in Angstrom; out_ev[count].  Host pointers. */
int32_t cheq_pair_integrals(cheq_ctx* ctx, uint64_t count, const double* distance, const uint8_t* n1,
                            const double* radius1, const uint8_t* n2, const double* radius2,
                            double lambda_scale, uint8_t basis, double* out_ev);

/* build_invariant_system's matrix without the border (implementation.rs:468-536): n x n, column-major,
 * diagonal = hardness, off-diagonal = J_eV, both triangles filled, input atom order.  Host pointers. */
int32_t cheq_assemble_matrix(cheq_ctx* ctx, uint64_t n_atoms, const double* xyz, const double* hardness,
                             const double* radius, const uint8_t* n_quantum, double lambda_scale,
                             uint8_t basis, double* matrix_out);

/* compute_external_potential (implementation.rs:369-447): v_out[n].  Host pointers. */
int32_t cheq_external_potential(cheq_ctx* ctx, uint64_t n_atoms, const double* xyz, const double* radius,
                                const uint8_t* n_quantum, const cheq_external* external,
                                double lambda_scale, uint8_t basis, double* v_out);

/* Dense kernels, exposed for unit tests.  All matrices column-major in HOST memory.
 *   cheq_dbg_cholesky_solve: solves A x = b for SPD A (n x n, lower triangle read) with the blocked
 *     tensor-core factorisation used by cheq_solve; nrhs <= 2.  Returns CHEQ_ERR_LINALG if not SPD.
 *   cheq_dbg_gemm_nt: C (m x n) = beta * C - ... : C <- C - A B^T when subtract != 0, else C <- A B^T;
 *     m, n, k multiples of 128. */
int32_t cheq_dbg_cholesky_solve(cheq_ctx* ctx, uint64_t n, const double* a, uint64_t nrhs, const double* b,
                                double* x_out, double* l_out /* nullable, n x n */);
int32_t cheq_dbg_gemm_nt(cheq_ctx* ctx, uint64_t m, uint64_t n, uint64_t k, const double* a,
                         const double* b, double* c, int32_t subtract);
/* The same product through the tcgen05 int8 slice path (cheq_set_ozaki): C <- C - A S B^T (subtract != 0) or
 * A S B^T, S = diag(signs) (nullable: identity; entries +-1), `lower` != 0 skips the 128 x 64 tiles entirely above the
 * diagonal; slices 4..8, k <= 8192. */
int32_t cheq_dbg_gemm_ozaki(cheq_ctx* ctx, uint64_t m, uint64_t n, uint64_t k, const double* a, const double* b,
                            double* c, int32_t subtract, int32_t lower, int32_t slices, const double* signs);
/* Times `reps` launches of the same GEMM on device-resident operands (CUDA events); ms_out = ms per launch.
 * lower != 0 skips the tiles above the diagonal (the symmetric trailing update). */
int32_t cheq_dbg_gemm_bench(cheq_ctx* ctx, uint64_t m, uint64_t n, uint64_t k, int32_t lower, int32_t subtract,
                            int32_t reps, double* ms_out);

/* Debug: SM-cycle counters of the diagonal-tile factorisation kernel accumulated since the last reset:
 * {load, phase A (diagonal 16 x 16 blocks), phase B (block inverse applied), phase C (rank-16 updates), store,
 * launches, 0, 0}. */
int32_t cheq_dbg_potrf_cycles(cheq_ctx* ctx, uint64_t* out8, int32_t reset);

/* Plain FP64 FMA (non-tensor) throughput of the device in TFLOP/s: the ALU roofline of the STO assembly. */
int32_t cheq_dbg_dfma_rate(cheq_ctx* ctx, double* tflops_out);

/* Host-only: evaluates the STO pair table (built exactly as for the device) in fp64 on the CPU.  Used to
 * test the table construction without a GPU; not a solve path.  Distance in bohr, result in Hartree. */
double cheq_host_sto_table_eval(int32_t n1, double zeta1, int32_t n2, double zeta2, double distance_bohr,
                                int32_t* kind_out, int32_t* terms_out, int32_t* coefs_out,
                                double* max_abs_err_out);

#ifdef __cplusplus
}
#endif
#endif /* CHEQ_B200_H */
