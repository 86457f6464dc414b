//! Raw bindings to `include/cheq_b200.h`.  NOT compiled in the build container (no Rust toolchain there);
//! kept mechanical so that it can be checked against the header by eye.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

pub const CHEQ_OK: i32 = 0;
pub const CHEQ_ERR_NO_ATOMS: i32 = 1;
pub const CHEQ_ERR_PARAMETER_NOT_FOUND: i32 = 2;
pub const CHEQ_ERR_NOT_CONVERGED: i32 = 3;
pub const CHEQ_ERR_LINALG: i32 = 4;
pub const CHEQ_ERR_CUDA: i32 = 5;
pub const CHEQ_ERR_INVALID_ARGUMENT: i32 = 6;
pub const CHEQ_ERR_OUT_OF_MEMORY: i32 = 7;

#[repr(C)]
pub struct cheq_ctx {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct cheq_options {
    pub tolerance: f64,
    pub lambda_scale: f64,
    pub damping_value: f64,
    pub max_iterations: u32,
    pub hydrogen_scf: u8,
    pub basis: u8,        // 0 GTO, 1 STO
    pub damping_kind: u8, // 0 None, 1 Fixed, 2 Auto
    pub reserved: u8,
}

#[repr(C)]
pub struct cheq_external {
    pub num_point_charges: u64,
    pub xyz: *const f64,
    pub charge: *const f64,
    pub radius: *const f64,
    pub n: *const u8,
    pub field: [f64; 3],
}

extern "C" {
    pub fn cheq_ctx_create(device: i32, out: *mut *mut cheq_ctx) -> i32;
    pub fn cheq_ctx_destroy(ctx: *mut cheq_ctx);
    pub fn cheq_last_error(ctx: *const cheq_ctx) -> *const c_char;
    pub fn cheq_options_default(opt: *mut cheq_options);
    pub fn cheq_ctx_stream(ctx: *const cheq_ctx) -> *mut c_void;
    pub fn cheq_solve(
        ctx: *mut cheq_ctx, n_atoms: u64, xyz: *const f64, chi: *const f64, hardness: *const f64,
        radius: *const f64, n_quantum: *const u8, total_charge: f64, options: *const cheq_options,
        external: *const cheq_external, charges_out: *mut f64, potential_out: *mut f64,
        iterations_out: *mut u32, delta_out: *mut f64,
    ) -> i32;
    /// Same solve without ever forming the (N + 1)^2 matrix (restarted GMRES, tiles recomputed on the fly).
    pub fn cheq_solve_iterative(
        ctx: *mut cheq_ctx, n_atoms: u64, xyz: *const f64, chi: *const f64, hardness: *const f64,
        radius: *const f64, n_quantum: *const u8, total_charge: f64, options: *const cheq_options,
        external: *const cheq_external, charges_out: *mut f64, potential_out: *mut f64,
        iterations_out: *mut u32, delta_out: *mut f64,
    ) -> i32;
    /// Per-iteration trace of the last single-frame solve (max-abs charge change, solve mode, GMRES steps).
    pub fn cheq_get_scf_trace(
        ctx: *const cheq_ctx, capacity: u32, delta_out: *mut f64, mode_out: *mut i32, steps_out: *mut i32,
        count_out: *mut u32,
    ) -> i32;
    /// Stale-factor policy of the hydrogen SCF iterations (0 = refactor every iteration).
    pub fn cheq_set_krylov(
        ctx: *mut cheq_ctx, enabled: i32, min_hydrogen: i32, freeze_ev: f64, force_iteration: i32,
    ) -> i32;
    /// FP64-equivalent updates on the tcgen05 tensor cores (int8 slice products): 0 = off, 4..8 slices (7 = fp64 grade).
    pub fn cheq_set_ozaki(ctx: *mut cheq_ctx, slices: i32, slices_r: i32, min_k: i32) -> i32;
    pub fn cheq_solve_batch(
        ctx: *mut cheq_ctx, n_frames: u64, n_atoms: u64, xyz: *const f64, chi: *const f64,
        hardness: *const f64, radius: *const f64, n_quantum: *const u8, total_charge: f64,
        options: *const cheq_options, charges_out: *mut f64, potential_out: *mut f64,
        iterations_out: *mut u32, delta_out: *mut f64, status_out: *mut i32,
    ) -> i32;

    // ---- one dense system over the GPUs of a node (one process per GPU; include/cheq_b200.h) ----
    /// Rank 0: fills 128 bytes that every rank must receive (any transport) before `cheq_ctx_dist_init`.
    pub fn cheq_dist_unique_id(id_out: *mut u8) -> i32;
    /// Joins the group; afterwards single-frame solves on `ctx` are collective (same inputs on every rank).
    pub fn cheq_ctx_dist_init(ctx: *mut cheq_ctx, rank: i32, world: i32, id: *const u8) -> i32;
    pub fn cheq_ctx_dist_info(ctx: *const cheq_ctx, rank_out: *mut i32, world_out: *mut i32) -> i32;
    /// 0 = automatic schedule, 1 = right-looking panels with look-ahead also on one GPU.
    pub fn cheq_set_factor_mode(ctx: *mut cheq_ctx, mode: i32, tiles_per_panel: i32) -> i32;
    /// Host-only: owner rank and panel index of every 128-column tile.
    pub fn cheq_dist_plan(
        x_tiles: i32, h_tiles: i32, tiles_per_panel: i32, world: i32, tile_owner_out: *mut i32,
        tile_panel_out: *mut i32,
    ) -> i32;
}

pub const CHEQ_DIST_ID_BYTES: usize = 128;
