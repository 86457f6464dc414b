// Shared host/device declarations for the B200 QEq engine.
#pragma once
#include <cstdint>

namespace cheq {

// src/shielding/constants.rs:13,22 and src/solver/implementation.rs:24
constexpr double kBohrToAngstrom = 0.529177210903;
constexpr double kHartreeToEv = 27.211386245988;
constexpr double kHChargeFactor = 0.93475415965;
constexpr double kDistanceThresholdBohr = 1e-12;

constexpr int kTile = 128;          // tile edge of every dense kernel; all blocks are padded to it
constexpr int kRhsRows = 128;       // extra rows below the matrix that carry the two right-hand sides
constexpr uint8_t kPadType = 255;   // element-type id of padding atoms

// Per-frame SCF state, resident on the device (mirrors the locals of run_scf_iterations,
// implementation.rs:294-339).
struct ScfState {
    double damping;
    double prev_delta;
    double delta;        // max |x_i - q_i| of the last solve
    double mu;           // equilibrated potential of the last solve
    unsigned long long delta_bits;  // running max accumulator (positive doubles compare as integers)
    int iteration;
    int converged;
    int info;            // != 0: factorisation met a non-positive pivot
    int active;          // 0: frame finished (converged or failed); kernels skip it
};

// Pair-type tables in Angstrom units, one flat blob.  Offsets in units of the respective element type.
struct PairTableView {
    int num_types;
    int basis;             // 0 GTO, 1 STO
    const double* rcut2;   // [T*T] squared distance beyond which F = 0              (STO)
    const double* rzero2;  // [T*T] squared distance below which J = j0
    const double* j0;      // [T*T] eV
    const double* beta;    // [T*T] 1/Angstrom                                        (GTO)
    const int* term_begin; // [T*T]
    const int* term_count; // [T*T]
    const double* term_c;  // [nterms] 1/Angstrom
    const int* coef_begin; // [nterms]
    const int* coef_count; // [nterms]
    const double* coef;    // [ncoef], coefficient of R_Angstrom^k
};

}  // namespace cheq
