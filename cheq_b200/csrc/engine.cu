// Host orchestration and C ABI of the B200 QEq engine (include/cheq_b200.h).
//
// One call = QEqSolver::solve / solve_in_field after element lookup
// (/root/reference/src/solver/implementation.rs:174-188, 239-266):
//   typing + ordering (host, O(N)) -> pair tables (host, per distinct element pair, cached)
//   -> v_ext kernel (:369-447) -> matrix assembly kernel (:454-554)
//   -> blocked Cholesky of the non-hydrogen block and the hydrogen Schur complement, once
//   -> SCF loop (:273-345): only the hydrogen block is refactorised per iteration; damping, clamp and
//      convergence logic run on the device, the host reads back one small record per iteration.
// There is no CPU fallback: every numerical step is a kernel of kernels_*.cu.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/cheq_b200.h"
#include "device_types.h"
#include "kernels.h"
#include "sto_tables.h"

using namespace cheq;

namespace {

struct DeviceBuffer {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 16;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return (T*)p; }
};

using Clock = std::chrono::steady_clock;
double ms_since(Clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
}

struct TableKey {
    int n1, n2;
    uint64_t z1, z2;
    bool operator<(const TableKey& o) const { return std::tie(n1, n2, z1, z2) < std::tie(o.n1, o.n2, o.z1, o.z2); }
};
uint64_t bits(double x) {
    uint64_t u;
    std::memcpy(&u, &x, 8);
    return u;
}

struct ElementType {
    uint8_t n;
    double radius;
};

// Internal ordering of one topology.
struct Layout {
    long n = 0;
    int nx = 0, nh = 0, nxp = 0, nhp = 0, NP = 0, MT = 0;
    long ld = 0;
    bool scf = false;
    std::vector<int> perm;            // NP, internal -> input, -1 padding
    std::vector<uint8_t> type;        // NP
    std::vector<ElementType> types;   // distinct (n, radius), atoms first, then external charges
};

int round_up_tile(int v) { return ((v + kTile - 1) / kTile) * kTile; }

// Where to split each panel of the recursive factorisation.  Halving is flop-optimal but the trailing update of a
// panel is a grid of 128 x 128 tiles run in waves of 148 CTAs; a split one tile to the left or right can save a
// whole wave of K = 128 n1 work.  cost() models a launch as waves x (K/16 x t_k + t_fixed) with the constants
// measured on B200 (tools/gemm_bench.py), and a dynamic programme over (first tile, tile count) picks the splits.
struct SplitPlan {
    int mt_total = 0;   // tile rows of the trapezoid (NP / 128 + 1)
    int batch = 1;
    int nmax = 0;
    std::vector<short> split;   // [j0t * (nmax + 1) + n] = tiles in the left part
    std::vector<float> cost;
    int oz_slices = 0, oz_min_k = 0;   // > 0: single-frame updates at least this deep run as int8 slice products
    // tcgen05 path (kernels_ozaki.cu), microseconds, from tools/ozaki_probe on B200: 128 x 64 tiles in waves of 148,
    // 2.6 us per 128-deep K tile at 7 slices (28 slice products) + ~4 us of prologue / epilogue per CTA, plus the
    // slicing passes (0.13 us per 128 x 128 block of an operand and five short launches)
    static double gemm_cost_oz(long mt, long nt, long k_tiles, bool lower, int S) {
        const long tiles = lower ? nt * (nt + 1) / 2 + (mt - nt) * nt : mt * nt;
        if (tiles <= 0) return 0.0;
        const double per_k = 2.6 * (S * (S + 1) / 2.0) / 28.0;
        return std::ceil(2.0 * tiles / 148.0) * (k_tiles * per_k + 4.0) + (double)(mt + nt) * k_tiles * 0.13 + 12.0;
    }
    double gemm_cost(long mt, long nt, long k_tiles, bool lower, int batch) const {
        if (oz_slices > 0 && batch == 1 && k_tiles * kTile >= oz_min_k) return gemm_cost_oz(mt, nt, k_tiles, lower, oz_slices);
        long tiles = lower ? nt * (nt + 1) / 2 + (mt - nt) * nt : mt * nt;
        if (tiles <= 0) return 0.0;
        tiles *= batch;
        const double ktiles16 = k_tiles * 8.0;
        if (tiles <= 74) return std::ceil(2.0 * tiles / 148.0) * (ktiles16 * 1.25 + 8.0);
        return std::ceil(tiles / 148.0) * (ktiles16 * 2.44 + 9.5);
    }
    void build(int mt_total_, int j_begin, int j_end, int batch_, int oz_slices_ = 0, int oz_min_k_ = 0) {
        // plan for panels inside tile columns [j_begin, j_end)
        mt_total = mt_total_;
        batch = batch_;
        oz_slices = oz_slices_;
        oz_min_k = oz_min_k_;
        nmax = j_end;
        split.assign((size_t)(nmax + 1) * (nmax + 1), 0);
        cost.assign((size_t)(nmax + 1) * (nmax + 1), 0.0f);
        const double leaf = 67.0;
        for (int n = 1; n <= j_end - j_begin; ++n)
            for (int j0 = j_begin; j0 + n <= j_end; ++j0) {
                const size_t id = (size_t)j0 * (nmax + 1) + n;
                if (n == 1) {
                    cost[id] = (float)(leaf + gemm_cost(mt_total - j0 - 1, 1, 1, false, batch));
                    continue;
                }
                double best = 1e30;
                int arg = n / 2;
                for (int n1 = 1; n1 < n; ++n1) {
                    const double c = cost[(size_t)j0 * (nmax + 1) + n1] +
                                     gemm_cost(mt_total - (j0 + n1), n - n1, n1, true, batch) +
                                     cost[(size_t)(j0 + n1) * (nmax + 1) + (n - n1)];
                    if (c < best - 1e-9) {
                        best = c;
                        arg = n1;
                    }
                }
                cost[id] = (float)best;
                split[id] = (short)arg;
            }
    }
    int left_tiles(int j0t, int n) const {
        if (j0t + n > nmax || split.empty()) return n / 2;
        const int v = split[(size_t)j0t * (nmax + 1) + n];
        return v > 0 ? v : n / 2;
    }
};

}  // namespace

// Ownership of the column panels of one layout (distributed / look-ahead factorisation).
// The X block (non-hydrogen) and the H block (n == 1 atoms) are cut separately into panels of `tpp` tile columns
// (the last panel of a block may be narrower); panel p belongs to rank p % world (1-D block-cyclic).
struct PanelPlan {
    int nxt = -1, nht = -1, tpp = -1, world = -1, rank = -1;
    std::vector<int> begin, end, owner;   // per panel, tile units
    int x_panels = 0;                     // panels [0, x_panels) cover the X block
    std::vector<int> own_tiles;           // tile columns owned by `rank`, ascending
    std::vector<uint8_t> mask;            // per tile column: 1 = owned by `rank`
    void build(int nxt_, int nht_, int tpp_, int world_, int rank_) {
        nxt = nxt_; nht = nht_; tpp = tpp_; world = world_; rank = rank_;
        begin.clear(); end.clear(); owner.clear(); own_tiles.clear();
        mask.assign((size_t)nxt + nht, 0);
        auto cut = [&](int t0, int t1) {
            for (int t = t0; t < t1; t += tpp) {
                begin.push_back(t);
                end.push_back(std::min(t + tpp, t1));
                owner.push_back((int)(owner.size() % (size_t)world));
            }
        };
        cut(0, nxt);
        x_panels = (int)begin.size();
        cut(nxt, nxt + nht);
        for (size_t p = 0; p < begin.size(); ++p)
            if (owner[p] == rank)
                for (int t = begin[p]; t < end[p]; ++t) {
                    own_tiles.push_back(t);
                    mask[t] = 1;
                }
    }
};

// NCCL entry points, resolved at run time (libnccl.so.2 is the one torch already loaded when the host process
// initialised torch.distributed; the engine itself has no link-time dependency on it).
struct NcclId { char internal[128]; };
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
constexpr int kNcclChar = 0, kNcclInt = 2, kNcclDouble = 8, kNcclMax = 2;

struct MatfreeBuffers {
    DeviceBuffer vec, part, blocks, minv, sgn, hflag, basis;
    void release() {
        for (DeviceBuffer* b : {&vec, &part, &blocks, &minv, &sgn, &hflag, &basis}) b->release();
    }
};

struct cheq_ctx {
    int device = 0;
    MatfreeBuffers* mf = nullptr;          // workspace of the matrix-free path (allocated on first use)
    cudaStream_t stream = nullptr;
    cudaStream_t cur = nullptr;            // stream the factorisation kernels are launched on (stream or panel_stream)
    // distributed / look-ahead factorisation
    int rank = 0, world = 1;
    void* comm = nullptr;                  // ncclComm_t
    cudaStream_t panel_stream = nullptr;   // high priority: panel factorisation + broadcasts
    cudaEvent_t ev_ready = nullptr, ev_bcast = nullptr;
    int factor_mode = 0;                   // 0: auto (recursive on one GPU, blocked when world > 1), 1: blocked
    int tiles_per_panel = 4;
    bool tiles_per_panel_set = false;      // chosen by the caller (API / CHEQ_TILES_PER_PANEL): no automatic widening
    int tpp_now = 4;                       // panel width of the current plan
    PanelPlan pplan;
    DeviceBuffer stage, own_list, col_mask, chain;
    // peer-memory data plane (multi-GPU): the owner's kernels store finished panel tiles straight into the peers'
    // trapezoids over NVLink; `arena` holds what must be peer-visible besides the trapezoid (signal words, the
    // diagonal-tile inverses, pivot signs, negative-pivot lists)
    struct PeerLinks {
        bool tried = false, enabled = false;
        int n = 0;                          // world - 1
        int rank_of[kMaxPeers] = {};
        void* opened[kMaxPeers][2] = {};    // mappings returned by cudaIpcOpenMemHandle (trapezoid, arena)
        char* A[kMaxPeers] = {};            // peer trapezoid / arena bases (mapping + offset inside the allocation)
        char* arena[kMaxPeers] = {};
        void* local[2] = {nullptr, nullptr};
        unsigned seq = 0;                   // panels signalled so far (identical on every rank)
    } links;
    DeviceBuffer arena, ipc_stage;
    int p2p_mode = 1;                       // CHEQ_DIST_P2P: 0 NCCL broadcasts only, 1 automatic, 2 always push
    bool push_now = false;                  // decision for the current solve
    bool push_dma = false;                  // push with the copy engines (cudaMemcpy2DAsync to the peers) instead of
                                            // from the kernels' epilogues
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_leaf = nullptr, ev_copy = nullptr;
    cudaStream_t bulk_stream = nullptr;     // below-the-panel work of the tile-level look-ahead
    cudaEvent_t ev_c[64] = {}, ev_b[64] = {};
    int panel_flat = 1;                     // CHEQ_PANEL_FLAT=0: recursive kernel sequence inside a panel
    int (*wait_value32)(cudaStream_t, unsigned long long, unsigned, unsigned) = nullptr;
    int (*mem_address_range)(unsigned long long*, size_t*, unsigned long long) = nullptr;
    double ms_comm_host = 0.0;
    std::mutex mu;
    std::string err;
    std::map<TableKey, StoPairTable> sto_cache;
    cheq_stats stats{};
    // device workspace (grow-only)
    DeviceBuffer in_xyz, in_chi, in_hard, in_radius, in_nq, perm, type, x, y, z, diag, chi, hard, vext, A, S0,
        Linv, v, q, scratch, counters, state, active, blob, pc_xyz, pc_q, pc_type, out_q, misc0, misc1, misc2,
        lu_f, lu_ut, lu_piv, sgn, sgn_flags, xsol;
    // stale-factor SCF iterations (kernels_krylov.cu): explicit R = L_hh^-T S of the frozen hydrogen factor, partial
    // sums of the tiled matrix-vector products, the small vectors of GMRES, device / pinned-host scalars
    DeviceBuffer kr_R, kr_Rf, kr_S0t, kr_tb, kr_part, kr_vec, kr_sc, kr_wx, kx_local;
    unsigned kx_seq = 0;              // rounds of the peer-memory all-reduce so far (identical on every rank)
    int krylov_p2p = 0;               // CHEQ_KRYLOV_P2P: 1 peer-memory all-reduce, 0 ncclAllReduce (default), -1 peer memory
                                      // from 4 ranks on.  Measured S20k: 0.221 vs 0.216 s at 2 ranks, 0.1733 vs 0.1732 s at 8: the three
                                      // all-reduces of a step are not what limits it, so NCCL stays the default)
    int krylov_f32 = 1;               // CHEQ_KRYLOV_F32=0: apply the frozen factor from the fp64 R
    KrylovScalars* h_ksc = nullptr;   // pinned
    int morton = 1;                   // CHEQ_MORTON=0: atoms of one element stay in input order (test knob: a
                                      // different elimination order for the unpivoted factorisation)
    int krylov_mode = 1;              // CHEQ_KRYLOV=0: refactor the hydrogen block in every SCF iteration
    int krylov_min_nhp = 4096;        // CHEQ_KRYLOV_MIN_NH: smallest padded hydrogen block that uses the stale factor (B3k,
                                      // nhp = 2048: 33.6 ms with it vs 23.7 ms without -- its GMRES steps are launch-latency-bound)
    double krylov_freeze_ev = 0.0;    // CHEQ_KRYLOV_FREEZE_EV: freeze once max |D_k - D_(k-1)| falls below this many eV;
                                      // 0 = from the cost model in factor_and_scf (3.5 .. 12 eV)
    int krylov_force_it = 5;          // ... or at this SCF iteration at the latest
    // FP64-equivalent updates on the tcgen05 tensor cores (kernels_ozaki.cu): int8 slices per operand, 0 = DMMA only
    int ozaki_slices = 7;             // CHEQ_OZAKI: 0 = off, 4..8 (7 = 54 bits per entry relative to its row maximum: fp64
                                      // grade; measured on S20k: charges 9e-12 e from the CPU oracle, DMMA 1.4e-13, 8 slices 4e-14)
    int ozaki_slices_r = 5;           // CHEQ_OZAKI_R: slices while building the frozen factor's R (a preconditioner, kept in
                                      // fp32 afterwards: 38 bits are plenty); 0 = as CHEQ_OZAKI
    int ozaki_min_k = 512;            // CHEQ_OZAKI_MIN_K: shallower updates stay on DMMA (slicing + epilogue do not amortise;
                                      // S20k: 0.1957 / 0.1909 / 0.1961 s with 1024 / 512 / 256)
    int force_lu = 0;                 // CHEQ_FORCE_LU=1: single-frame solves take the pivoted-LU path (timing / test knob)
    int ozaki_now = 0;                // slices in force for the launches being issued
    bool oz_same_a = false;           // set by a caller whose next update multiplies the SAME left operand as its previous one
    struct { const double* A = nullptr; long lda = 0; int mt = 0, K = 0, S = 0; } oz_last;   // A operand whose planes are in oz_ws
    DeviceBuffer oz_ws, oz_err;
    int* h_oz_err = nullptr;          // pinned
    int matfree_restart = 160;        // CHEQ_MATFREE_RESTART: GMRES restart length of the matrix-free path
    int matfree_max_steps = 4000;     // CHEQ_MATFREE_MAX_STEPS: operator applications per SCF iteration
    double matfree_tol = 1e-12;       // CHEQ_MATFREE_TOL: |P (b - K x)| <= tol |x|
    struct TraceEntry { double delta; int mode; int steps; };   // mode 0: refactored, 1: stale factor + GMRES, 2: matrix-free GMRES
    std::vector<TraceEntry> trace;    // per SCF iteration of the last single-frame solve
    ScfState* h_state = nullptr;   // pinned
    int* h_active = nullptr;       // pinned
    size_t h_state_cap = 0;
    cudaEvent_t ev[8]{};
    uint64_t launches = 0;
    // optional per-launch timing of the dominant kernel (gemm_nt_kernel) with CUDA events on the stream
    std::map<std::tuple<int, int, int, int, int>, SplitPlan> plans;   // (mt_total, j_begin, j_end, batch, tcgen05 setting)
    bool profile = false;
    std::vector<cudaEvent_t> prof_events;   // pairs (start, stop)
    std::vector<int> prof_cat;              // category of each pair: 0 GEMM, 1 diagonal-tile factorisation, 2 back substitution
    struct ProfShape { int K, mt, nt; double flops; };
    std::vector<ProfShape> prof_shape;      // GEMM shapes, for the optional launch log (env CHEQ_GEMM_LOG)
    size_t prof_used = 0;
    double prof_flops = 0.0;
};

namespace {

#define CU(call)                                                                         \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);              \
            cudaGetLastError();                                                          \
            return (e__ == cudaErrorMemoryAllocation) ? CHEQ_ERR_OUT_OF_MEMORY : CHEQ_ERR_CUDA; \
        }                                                                                \
    } while (0)

#define LAUNCH(call)      \
    do {                  \
        CU(call);         \
        ctx->launches++;  \
    } while (0)

int fail(cheq_ctx* ctx, int code, const std::string& msg) {
    ctx->err = msg;
    return code;
}

double slater_exponent(unsigned n, double r_cov_bohr, double lambda) {
    // sto.rs:43-48 / gto.rs:47-52
    if (r_cov_bohr < kDistanceThresholdBohr) return 0.0;
    return lambda * (2.0 * (double)n + 1.0) / (2.0 * r_cov_bohr);
}
double gto_fitting_coefficient(unsigned n) {
    // gto.rs:69-78
    switch (n) {
        case 1: return 0.270917;
        case 2: return 0.098800;
        case 3: return 0.055600;
        case 4: return 0.039100;
        case 5: return 0.029600;
        default: return 0.27 * std::pow((double)n, -1.35);
    }
}
double gto_gaussian_exponent(unsigned n, double zeta) {
    // gto.rs:56-62
    if (n == 0) return 0.0;
    return gto_fitting_coefficient(n) * zeta * zeta / (double)n;
}

int find_or_add_type(std::vector<ElementType>& types, std::map<std::pair<uint8_t, uint64_t>, int>& index,
                     uint8_t n, double radius) {
    auto key = std::make_pair(n, bits(radius));
    auto it = index.find(key);
    if (it != index.end()) return it->second;
    int id = (int)types.size();
    types.push_back({n, radius});
    index.emplace(key, id);
    return id;
}

// Typing and ordering: non-hydrogen block first, then the n == 1 block (implementation.rs:475-477), each
// sorted by element type so that warps of the assembly kernel see one pair type.
// 30-bit Morton code of a position inside the bounding box: atoms of one element type are laid out along a
// space-filling curve, so a 128 x 64 assembly tile pairs two compact groups of atoms and whole warps fall on
// the same side of the integral's cutoff (far pairs skip the exponential).
uint32_t morton30(double fx, double fy, double fz) {
    auto spread = [](uint32_t v) {
        v &= 0x3ff;
        v = (v | (v << 16)) & 0x030000ff;
        v = (v | (v << 8)) & 0x0300f00f;
        v = (v | (v << 4)) & 0x030c30c3;
        v = (v | (v << 2)) & 0x09249249;
        return v;
    };
    auto q = [](double f) { return (uint32_t)std::min(1023.0, std::max(0.0, f * 1024.0)); };
    return spread(q(fx)) | (spread(q(fy)) << 1) | (spread(q(fz)) << 2);
}

int make_layout(cheq_ctx* ctx, long n, const double* radius, const uint8_t* nq, bool hydrogen_scf,
                const double* pc_radius, const uint8_t* pc_n, long m, const double* xyz, Layout& L,
                std::vector<int>& pc_type, int spatial_pad = 0) {
    L.n = n;
    std::map<std::pair<uint8_t, uint64_t>, int> index;
    std::vector<int> atom_type(n);
    for (long i = 0; i < n; ++i) atom_type[i] = find_or_add_type(L.types, index, nq[i], radius[i]);
    pc_type.resize(m);
    for (long k = 0; k < m; ++k) pc_type[k] = find_or_add_type(L.types, index, pc_n[k], pc_radius[k]);
    if (L.types.size() >= kPadType) return fail(ctx, CHEQ_ERR_INVALID_ARGUMENT, "more than 254 distinct element types");
    bool any_h = false;
    for (long i = 0; i < n; ++i) any_h |= (nq[i] == 1);
    L.scf = hydrogen_scf && any_h;   // implementation.rs:279-280
    std::vector<int> xs, hs;
    xs.reserve(n);
    for (long i = 0; i < n; ++i) {
        if (L.scf && nq[i] == 1 && !spatial_pad) hs.push_back((int)i);
        else xs.push_back((int)i);
    }
    std::vector<uint32_t> code(n, 0);
    if (xyz && (spatial_pad || (n > 256 && ctx->morton))) {
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (long i = 0; i < n; ++i)
            for (int d = 0; d < 3; ++d) {
                lo[d] = std::min(lo[d], xyz[3 * i + d]);
                hi[d] = std::max(hi[d], xyz[3 * i + d]);
            }
        double ext = 1e-9;
        for (int d = 0; d < 3; ++d) ext = std::max(ext, hi[d] - lo[d]);
        for (long i = 0; i < n; ++i)
            code[i] = morton30((xyz[3 * i] - lo[0]) / ext, (xyz[3 * i + 1] - lo[1]) / ext, (xyz[3 * i + 2] - lo[2]) / ext);
    }
    auto by_type = [&](int a, int b) {
        // matrix-free path (spatial_pad): one block, pure space-filling-curve order, so that 128 consecutive atoms
        // are a compact cluster of whole molecules (the block-Jacobi preconditioner's blocks)
        if (!spatial_pad && atom_type[a] != atom_type[b]) return atom_type[a] < atom_type[b];
        if (code[a] != code[b]) return code[a] < code[b];
        return a < b;
    };
    std::sort(xs.begin(), xs.end(), by_type);
    std::sort(hs.begin(), hs.end(), by_type);
    L.nx = (int)xs.size();
    L.nh = (int)hs.size();
    L.nxp = round_up_tile(L.nx);
    L.nhp = round_up_tile(L.nh);
    if (spatial_pad) L.nxp = (int)(((long)L.nx + spatial_pad - 1) / spatial_pad * spatial_pad);
    L.NP = L.nxp + L.nhp;
    L.MT = L.NP + kRhsRows;
    L.ld = L.MT;
    L.perm.assign(L.NP, -1);
    L.type.assign(L.NP, kPadType);
    for (int i = 0; i < L.nx; ++i) {
        L.perm[i] = xs[i];
        L.type[i] = (uint8_t)atom_type[xs[i]];
    }
    for (int i = 0; i < L.nh; ++i) {
        L.perm[L.nxp + i] = hs[i];
        L.type[L.nxp + i] = (uint8_t)atom_type[hs[i]];
    }
    return CHEQ_OK;
}

struct PackedTables {
    std::vector<char> blob;
    PairTableView view{};   // pointers are OFFSETS into blob until rebased
    int pair_types = 0;
    double rcut2_max = 0.0;  // STO: largest squared cutoff over all pair types [Angstrom^2]
};

// Builds the pair tables for all type pairs, in Angstrom units.
int pack_tables(cheq_ctx* ctx, const std::vector<ElementType>& types, double lambda, int basis, PackedTables& out) {
    const int T = (int)types.size();
    std::vector<double> rcut2(T * T, 0.0), rzero2(T * T, 0.0), j0(T * T, 0.0), beta(T * T, 0.0), term_c, coef;
    std::vector<int> term_begin(T * T, 0), term_count(T * T, 0), coef_begin, coef_count;
    std::vector<double> zeta(T);
    for (int i = 0; i < T; ++i) zeta[i] = slater_exponent(types[i].n, types[i].radius / kBohrToAngstrom, lambda);
    const double B = kBohrToAngstrom;
    for (int i = 0; i < T; ++i)
        for (int j = i; j < T; ++j) {
            const int p = i * T + j, pt = j * T + i;
            if (basis == CHEQ_BASIS_GTO) {
                // gto.rs:36-42, 82-90
                const double a1 = gto_gaussian_exponent(types[i].n, zeta[i]);
                const double a2 = gto_gaussian_exponent(types[j].n, zeta[j]);
                const double b = std::sqrt(2.0 * a1 * a2 / (a1 + a2));
                beta[p] = beta[pt] = b / B;
                // `r > DISTANCE_THRESHOLD_BOHR` else 2 beta / sqrt(pi): r2 < rzero2 takes the else branch
                const double rz = kDistanceThresholdBohr * B;
                rzero2[p] = rzero2[pt] = std::nextafter(rz * rz, 1.0);
                j0[p] = j0[pt] = 2.0 * b / std::sqrt(M_PI) * kHartreeToEv;
                continue;
            }
            TableKey key{types[i].n, types[j].n, bits(zeta[i]), bits(zeta[j])};
            auto it = ctx->sto_cache.find(key);
            if (it == ctx->sto_cache.end())
                it = ctx->sto_cache.emplace(key, build_sto_pair_table(types[i].n, zeta[i], types[j].n, zeta[j])).first;
            const StoPairTable& tab = it->second;
            if (tab.kind == STO_KIND_INVALID) {
                // zeta == 0 or n == 0: undefined in the reference's dependency; propagate NaN
                rzero2[p] = rzero2[pt] = 1e300;
                j0[p] = j0[pt] = std::nan("");
                out.rcut2_max = INFINITY;   // no fast path: the NaN must reach every pair of this type
                continue;
            }
            rcut2[p] = rcut2[pt] = (tab.r_cut * B) * (tab.r_cut * B);
            out.rcut2_max = std::max(out.rcut2_max, rcut2[p]);
            rzero2[p] = rzero2[pt] = (tab.r_zero * B) * (tab.r_zero * B);
            j0[p] = j0[pt] = tab.j0 * kHartreeToEv;
            term_begin[p] = term_begin[pt] = (int)term_c.size();
            term_count[p] = term_count[pt] = (int)tab.terms.size();
            for (const StoTerm& t : tab.terms) {
                term_c.push_back(t.c / B);
                coef_begin.push_back((int)coef.size());
                coef_count.push_back((int)t.coef.size());
                double scale = 1.0;
                for (double c : t.coef) {
                    coef.push_back(c * scale);
                    scale /= B;
                }
            }
            out.pair_types++;
        }
    if (basis == CHEQ_BASIS_GTO) out.pair_types = T * (T + 1) / 2;
    // layout: doubles, then ints
    const size_t nd = 4 * (size_t)T * T + term_c.size() + coef.size();
    const size_t ni = 2 * (size_t)T * T + coef_begin.size() + coef_count.size();
    size_t bytes = nd * 8 + ni * 4;
    bytes = (bytes + 15) / 16 * 16;
    out.blob.assign(bytes, 0);
    char* base = out.blob.data();
    size_t off = 0;
    auto put_d = [&](const std::vector<double>& v) {
        const double* p = (const double*)(uintptr_t)off;
        if (!v.empty()) std::memcpy(base + off, v.data(), v.size() * 8);
        off += v.size() * 8;
        return p;
    };
    auto put_i = [&](const std::vector<int>& v) {
        const int* p = (const int*)(uintptr_t)off;
        if (!v.empty()) std::memcpy(base + off, v.data(), v.size() * 4);
        off += v.size() * 4;
        return p;
    };
    out.view.num_types = T;
    out.view.basis = basis;
    out.view.rcut2 = put_d(rcut2);
    out.view.rzero2 = put_d(rzero2);
    out.view.j0 = put_d(j0);
    out.view.beta = put_d(beta);
    out.view.term_c = put_d(term_c);
    out.view.coef = put_d(coef);
    out.view.term_begin = put_i(term_begin);
    out.view.term_count = put_i(term_count);
    out.view.coef_begin = put_i(coef_begin);
    out.view.coef_count = put_i(coef_count);
    return CHEQ_OK;
}

PairTableView rebase_view(PairTableView v, const char* base) {
    auto mv = [&](const void* p) { return (const void*)(base + (uintptr_t)p); };
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

int upload_tables(cheq_ctx* ctx, const PackedTables& pk, PairTableView& dev_view) {
    CU(ctx->blob.ensure(pk.blob.size()));
    CU(cudaMemcpyAsync(ctx->blob.p, pk.blob.data(), pk.blob.size(), cudaMemcpyHostToDevice, ctx->stream));
    dev_view = rebase_view(pk.view, ctx->blob.as<char>());
    return CHEQ_OK;
}

// |pivot| below this is treated as a breakdown of the unpivoted factorisation -> pivoted-LU path.  The test is
// RELATIVE to the matrix scale: kPivotRel times the largest hardness of the system (the diagonal of the
// matrix, 4..30 eV for the default table), so accepted pivots bound the element growth by ~1e5 (error ~1e-11) whatever
// the units of the parameter set; kPivotTol is the absolute value used where no hardness is at hand (the test hooks).
constexpr double kPivotTol = 1.0e-4;
constexpr double kPivotRel = 1.0e-5;

// ---- factorisation driver ------------------------------------------------------------------------------------
struct Dense {
    cheq_ctx* ctx;
    double* A;
    long ld, strideA;
    double* Linv;
    long strideLinv;
    int MT;
    ScfState* state;
    int batch;
    double* sgn;        // [batch][NP] pivot signs of A = L S L^T
    long strideSgn;
    int* neg_cnt;       // [batch][NP / 128] number of negative pivots per tile
    int* neg_idx;       // [batch][NP / 128][128] their in-tile indices
    long strideNeg;
    double flops = 0.0;
    double pivot_tol = kPivotTol;             // absolute threshold for this solve (relative test resolved on the host)
    const struct SplitPlan* plan = nullptr;   // wave-aware recursion splits (nullptr: halve)
    bool push = false;                        // panel outputs are also stored into the peers' memory (epilogues)
    bool push_dma = false;                    // each finished tile column is copied to the peers by the copy engines
};



int prof_begin(cheq_ctx* ctx, int category) {
    if (!ctx->profile) return CHEQ_OK;
    if (ctx->prof_used + 2 > ctx->prof_events.size()) {
        const size_t old = ctx->prof_events.size();
        ctx->prof_events.resize(old + 4096, nullptr);
        for (size_t i = old; i < ctx->prof_events.size(); ++i) CU(cudaEventCreate(&ctx->prof_events[i]));
    }
    if (ctx->prof_cat.size() < ctx->prof_events.size() / 2) ctx->prof_cat.resize(ctx->prof_events.size() / 2);
    ctx->prof_cat[ctx->prof_used / 2] = category;
    CU(cudaEventRecord(ctx->prof_events[ctx->prof_used], ctx->cur));
    return CHEQ_OK;
}
int prof_end(cheq_ctx* ctx) {
    if (!ctx->profile) return CHEQ_OK;
    CU(cudaEventRecord(ctx->prof_events[ctx->prof_used + 1], ctx->cur));
    ctx->prof_used += 2;
    return CHEQ_OK;
}

int launch_gemm_timed(cheq_ctx* ctx, const GemmArgs& g, int mt, int nt, int batch, double flops) {
    // deep single-frame updates go to the tcgen05 tensor cores as int8 slice products (kernels_ozaki.cu)
    const int S = ctx->ozaki_now;
    bool oz = S > 0 && batch == 1 && g.subtract && g.n_peers == 0 && g.K >= ctx->ozaki_min_k &&
              ctx->cur == ctx->stream && ozaki_supported(S, std::min(g.K, kOzakiMaxK));
    const bool same_a = ctx->oz_same_a;
    ctx->oz_same_a = false;
    if (oz) {
        const size_t need = ozaki_workspace_bytes(S, mt, nt, std::min(g.K, kOzakiMaxK));
        if (need > ctx->oz_ws.cap) {
            CU(cudaStreamSynchronize(ctx->cur));   // earlier launches may still read the old workspace
            ctx->oz_last.A = nullptr;
            if (ctx->oz_ws.ensure(need) != cudaSuccess) {
                cudaGetLastError();                // no room for the slices: this solve stays on DMMA
                ctx->ozaki_now = 0;
                oz = false;
            }
        }
    }
    int rc = prof_begin(ctx, oz ? 3 : 0);
    if (rc) return rc;
    if (oz) {
        // K is cut into chunks whose int32 group sums cannot overflow; each chunk has its own row scales
        for (int k0 = 0; k0 < g.K; k0 += kOzakiMaxK) {
            GemmArgs c = g;
            c.K = std::min(kOzakiMaxK, g.K - k0);
            c.A = g.A + (long)k0 * g.lda;
            c.B = g.B + (long)k0 * g.ldb;
            if (g.neg_cnt) {
                c.neg_cnt = g.neg_cnt + k0 / kTile;
                c.neg_idx = g.neg_idx + k0;
            }
            const bool reuse = same_a && g.K <= kOzakiMaxK && ctx->oz_last.A == c.A && ctx->oz_last.lda == c.lda &&
                               ctx->oz_last.mt == mt && ctx->oz_last.K == c.K && ctx->oz_last.S == S;
            LAUNCH(launch_gemm_ozaki(c, mt, nt, S, ctx->oz_ws.p, ctx->oz_ws.cap, ctx->oz_err.as<int>(), ctx->cur, reuse));
            ctx->oz_last.A = c.A;
            ctx->oz_last.lda = c.lda;
            ctx->oz_last.mt = mt;
            ctx->oz_last.K = c.K;
            ctx->oz_last.S = S;
            ctx->launches += 4;   // two row-maximum and two slicing kernels ride along
            ctx->stats.tc_gemm_launches++;
        }
        ctx->stats.tc_gemm_flops += flops;
        ctx->stats.tc_gemm_ops += flops * (S * (S + 1) / 2);
    } else {
        LAUNCH(launch_gemm_nt(g, mt, nt, batch, ctx->cur));
    }
    if (ctx->profile) {
        ctx->prof_flops += flops * batch;
        if (ctx->prof_shape.size() < ctx->prof_cat.size()) ctx->prof_shape.resize(ctx->prof_cat.size());
        ctx->prof_shape[ctx->prof_used / 2] = {g.K, mt, nt, flops * batch};
    }
    return prof_end(ctx);
}

const SplitPlan* get_plan(cheq_ctx* ctx, int MT, int j_begin, int j_end, int batch) {
    if (j_end - j_begin < 2) return nullptr;
    const bool oz = ctx->ozaki_now > 0 && batch == 1 && ctx->cur == ctx->stream;
    auto key = std::make_tuple(MT / kTile, j_begin / kTile, j_end / kTile, batch, oz ? ctx->ozaki_now * 100000 + ctx->ozaki_min_k : 0);
    auto it = ctx->plans.find(key);
    if (it == ctx->plans.end()) {
        if (ctx->plans.size() > 64) ctx->plans.clear();
        it = ctx->plans.emplace(key, SplitPlan()).first;
        it->second.build(MT / kTile, j_begin / kTile, j_end / kTile, batch, oz ? ctx->ozaki_now : 0, ctx->ozaki_min_k);
    }
    return &it->second;
}

// Recursive tall-panel Cholesky of columns [j0, j0 + n) with all rows down to MT (the rows below the panel are
// overwritten by A L^-T, i.e. the triangular solve is part of every leaf).
int factor_panel(Dense& D, int j0, int n) {
    cheq_ctx* ctx = D.ctx;
    if (n <= 0) return CHEQ_OK;
    if (n == kTile) {
        double* Ajj = D.A + (long)j0 + (long)j0 * D.ld;
        double* Lk = D.Linv + (long)(j0 / kTile) * kTile * kTile;
        if (int rcp = prof_begin(ctx, 1)) return rcp;
        PotrfPeers pp{};
        if (D.push) {
            const cheq_ctx::PeerLinks& lk = ctx->links;
            pp.n = lk.n;
            for (int i = 0; i < lk.n; ++i) {
                // the arena-backed arrays of the peers sit at the same offsets as ours
                pp.A[i] = (double*)lk.A[i] + (Ajj - D.A);
                pp.Minv[i] = (double*)(lk.arena[i] + ((char*)Lk - ctx->arena.as<char>()));
                pp.sgn[i] = (double*)(lk.arena[i] + ((char*)(D.sgn + j0) - ctx->arena.as<char>()));
                pp.neg_cnt[i] = (int*)(lk.arena[i] + ((char*)(D.neg_cnt + j0 / kTile) - ctx->arena.as<char>()));
                pp.neg_idx[i] = (int*)(lk.arena[i] + ((char*)(D.neg_idx + (long)j0) - ctx->arena.as<char>()));
            }
        }
        LAUNCH(launch_potrf_tile(Ajj, D.ld, D.strideA, Lk, D.strideLinv, D.sgn + j0, D.strideSgn, D.neg_cnt + j0 / kTile,
                                 D.neg_idx + (long)j0, D.strideNeg, D.state, D.pivot_tol, D.batch, ctx->cur,
                                 D.push ? &pp : nullptr));
        if (int rcp = prof_end(ctx)) return rcp;
        const int mt = (D.MT - j0 - kTile) / kTile;
        if (mt > 0) {
            GemmArgs g{};
            g.A = Ajj + kTile;
            g.lda = D.ld;
            g.strideA = D.strideA;
            g.B = Lk;
            g.ldb = kTile;
            g.strideB = D.strideLinv;
            g.C = Ajj + kTile;
            g.ldc = D.ld;
            g.strideC = D.strideA;
            g.K = kTile;
            g.lower = 0;
            g.subtract = 0;
            g.state = D.state;
            if (D.push) {
                g.n_peers = ctx->links.n;
                for (int i = 0; i < g.n_peers; ++i) g.peerC[i] = (double*)ctx->links.A[i] + (g.C - D.A);
            }
            const int rc2 = launch_gemm_timed(ctx, g, mt, 1, D.batch, 2.0 * mt * kTile * (double)kTile * kTile);
            if (rc2) return rc2;
        }
        if (D.push_dma) {
            // the finished tile column (diagonal tile + everything below, right-hand-side rows included) goes to
            // every peer's trapezoid by DMA while this rank carries on with the panel
            CU(cudaEventRecord(ctx->ev_leaf, ctx->cur));
            CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_leaf, 0));
            const size_t width = (size_t)(D.MT - j0) * 8;
            for (int i = 0; i < ctx->links.n; ++i)
                CU(cudaMemcpy2DAsync((double*)ctx->links.A[i] + (Ajj - D.A), (size_t)D.ld * 8, Ajj, (size_t)D.ld * 8,
                                     width, kTile, cudaMemcpyDeviceToDevice, ctx->copy_stream));
        }
        const double m = (double)(D.MT - j0 - kTile);
        D.flops += (double)kTile * kTile * kTile / 3.0 + m * kTile * kTile;
        return CHEQ_OK;
    }
    const int n1 = (D.plan ? D.plan->left_tiles(j0 / kTile, n / kTile) : (n / kTile) / 2) * kTile;
    int rc = factor_panel(D, j0, n1);
    if (rc) return rc;
    {
        const int r0 = j0 + n1;
        GemmArgs g{};
        g.A = D.A + (long)r0 + (long)j0 * D.ld;
        g.lda = D.ld;
        g.strideA = D.strideA;
        g.B = g.A;
        g.ldb = D.ld;
        g.strideB = D.strideA;
        g.C = D.A + (long)r0 + (long)r0 * D.ld;
        g.ldc = D.ld;
        g.strideC = D.strideA;
        g.K = n1;
        g.lower = 1;
        g.subtract = 1;
        g.neg_cnt = D.neg_cnt + j0 / kTile;
        g.neg_idx = D.neg_idx + (long)j0;
        g.strideNeg = D.strideNeg;
        g.state = D.state;
        const int mt = (D.MT - r0) / kTile, nt = (n - n1) / kTile;
        const double w = (double)(n - n1), m = (double)(D.MT - r0);
        const double fl = (double)n1 * (w * w + 2.0 * (m - w) * w);   // syrk on the square part + gemm below it
        const int rc2 = launch_gemm_timed(ctx, g, mt, nt, D.batch, fl);
        if (rc2) return rc2;
        D.flops += fl;
    }
    return factor_panel(D, j0 + n1, n - n1);
}

// Panel factorisation with TILE-LEVEL look-ahead (blocked schedule).  The dependency chain of a panel of T tile
// columns is  potrf(t) -> in-panel solve (tiles t+1..T-1 of column t) -> in-panel update -> potrf(t+1): these run
// on the panel stream.  Everything below the panel -- the update of column t by the panel's earlier columns and its
// triangular solve -- runs on the bulk stream one tile column behind, concurrently with the next potrf.
int factor_panel_flat(Dense& D, int j0, int n) {
    cheq_ctx* ctx = D.ctx;
    const int T = n / kTile;
    cudaStream_t ps = ctx->panel_stream, bs = ctx->bulk_stream;
    const int rb = j0 + n;                      // first row below the panel
    const int mb = (D.MT - rb) / kTile;         // tile rows below it (>= 1: the right-hand-side rows)
    const double tile3 = (double)kTile * kTile * kTile;
    auto peers_of = [&](GemmArgs& g) {
        if (!D.push) return;
        g.n_peers = ctx->links.n;
        for (int i = 0; i < g.n_peers; ++i) g.peerC[i] = (double*)ctx->links.A[i] + (g.C - D.A);
    };
    for (int t = 0; t < T; ++t) {
        const int jt = j0 + t * kTile;
        double* Ajj = D.A + (long)jt + (long)jt * D.ld;
        double* Lk = D.Linv + (long)(jt / kTile) * kTile * kTile;
        // ---- chain (panel stream) ----
        ctx->cur = ps;
        if (int rcp = prof_begin(ctx, 1)) return rcp;
        PotrfPeers pp{};
        if (D.push) {
            const cheq_ctx::PeerLinks& lk = ctx->links;
            pp.n = lk.n;
            for (int i = 0; i < lk.n; ++i) {
                pp.A[i] = (double*)lk.A[i] + (Ajj - D.A);
                pp.Minv[i] = (double*)(lk.arena[i] + ((char*)Lk - ctx->arena.as<char>()));
                pp.sgn[i] = (double*)(lk.arena[i] + ((char*)(D.sgn + jt) - ctx->arena.as<char>()));
                pp.neg_cnt[i] = (int*)(lk.arena[i] + ((char*)(D.neg_cnt + jt / kTile) - ctx->arena.as<char>()));
                pp.neg_idx[i] = (int*)(lk.arena[i] + ((char*)(D.neg_idx + (long)jt) - ctx->arena.as<char>()));
            }
        }
        LAUNCH(launch_potrf_tile(Ajj, D.ld, D.strideA, Lk, D.strideLinv, D.sgn + jt, D.strideSgn, D.neg_cnt + jt / kTile,
                                 D.neg_idx + (long)jt, D.strideNeg, D.state, D.pivot_tol, D.batch, ps,
                                 D.push ? &pp : nullptr));
        if (int rcp = prof_end(ctx)) return rcp;
        D.flops += tile3 / 3.0;
        const int mt_in = T - 1 - t;            // tiles of this column that lie inside the panel, below the diagonal
        if (mt_in > 0) {
            GemmArgs g{};                       // in-panel solve: L[t+1.., t] = A[t+1.., t] M_t^T
            g.A = Ajj + kTile;
            g.lda = D.ld;
            g.strideA = D.strideA;
            g.B = Lk;
            g.ldb = kTile;
            g.strideB = D.strideLinv;
            g.C = Ajj + kTile;
            g.ldc = D.ld;
            g.strideC = D.strideA;
            g.K = kTile;
            g.state = D.state;
            peers_of(g);
            int rc = launch_gemm_timed(ctx, g, mt_in, 1, D.batch, 2.0 * mt_in * tile3);
            if (rc) return rc;
            GemmArgs u{};                       // in-panel update of the tiles right of column t
            u.A = Ajj + kTile;
            u.lda = D.ld;
            u.strideA = D.strideA;
            u.B = u.A;
            u.ldb = D.ld;
            u.strideB = D.strideA;
            u.C = D.A + (long)(jt + kTile) + (long)(jt + kTile) * D.ld;
            u.ldc = D.ld;
            u.strideC = D.strideA;
            u.K = kTile;
            u.lower = 1;
            u.subtract = 1;
            u.neg_cnt = D.neg_cnt + jt / kTile;
            u.neg_idx = D.neg_idx + (long)jt;
            u.strideNeg = D.strideNeg;
            u.state = D.state;
            const double fu = tile3 * mt_in * (double)mt_in;
            rc = launch_gemm_timed(ctx, u, mt_in, mt_in, D.batch, fu);
            if (rc) return rc;
            D.flops += mt_in * tile3 + fu;
        }
        CU(cudaEventRecord(ctx->ev_c[t], ps));
        // ---- below the panel (bulk stream) ----
        CU(cudaStreamWaitEvent(bs, ctx->ev_c[t], 0));
        ctx->cur = bs;
        if (t > 0) {
            GemmArgs r{};                       // A[below, t] -= sum_{t' < t} L[below, t'] S L[t, t']^T
            r.A = D.A + (long)rb + (long)j0 * D.ld;
            r.lda = D.ld;
            r.strideA = D.strideA;
            r.B = D.A + (long)jt + (long)j0 * D.ld;
            r.ldb = D.ld;
            r.strideB = D.strideA;
            r.C = D.A + (long)rb + (long)jt * D.ld;
            r.ldc = D.ld;
            r.strideC = D.strideA;
            r.K = t * kTile;
            r.subtract = 1;
            r.neg_cnt = D.neg_cnt + j0 / kTile;
            r.neg_idx = D.neg_idx + (long)j0;
            r.strideNeg = D.strideNeg;
            r.state = D.state;
            const double fr = 2.0 * mb * t * tile3;
            const int rc = launch_gemm_timed(ctx, r, mb, 1, D.batch, fr);
            if (rc) return rc;
            D.flops += fr;
        }
        {
            GemmArgs l{};                       // L[below, t] = A[below, t] M_t^T
            l.A = D.A + (long)rb + (long)jt * D.ld;
            l.lda = D.ld;
            l.strideA = D.strideA;
            l.B = Lk;
            l.ldb = kTile;
            l.strideB = D.strideLinv;
            l.C = D.A + (long)rb + (long)jt * D.ld;
            l.ldc = D.ld;
            l.strideC = D.strideA;
            l.K = kTile;
            l.state = D.state;
            peers_of(l);
            const int rc = launch_gemm_timed(ctx, l, mb, 1, D.batch, 2.0 * mb * tile3);
            if (rc) return rc;
            D.flops += mb * tile3;
        }
        CU(cudaEventRecord(ctx->ev_b[t], bs));
        if (D.push_dma) {
            // column t is final in both parts: DMA it to the peers while the panel goes on
            cudaStream_t cs = ctx->copy_stream;
            CU(cudaStreamWaitEvent(cs, ctx->ev_c[t], 0));
            CU(cudaStreamWaitEvent(cs, ctx->ev_b[t], 0));
            const size_t width = (size_t)(D.MT - jt) * 8;
            for (int i = 0; i < ctx->links.n; ++i)
                CU(cudaMemcpy2DAsync((double*)ctx->links.A[i] + (Ajj - D.A), (size_t)D.ld * 8, Ajj, (size_t)D.ld * 8,
                                     width, kTile, cudaMemcpyDeviceToDevice, cs));
        }
    }
    // "panel stream done" must mean that the whole panel is done
    CU(cudaStreamWaitEvent(ps, ctx->ev_b[T - 1], 0));
    ctx->cur = ps;
    return CHEQ_OK;
}

// Schur update of the trailing block by an already factored leading block [0, k): C[k:MT, k:NP] -= W W^T.
int schur_update(Dense& D, int k, int ncols) {
    cheq_ctx* ctx = D.ctx;
    if (k <= 0 || ncols <= 0) return CHEQ_OK;
    GemmArgs g{};
    g.A = D.A + (long)k;
    g.lda = D.ld;
    g.strideA = D.strideA;
    g.B = g.A;
    g.ldb = D.ld;
    g.strideB = D.strideA;
    g.C = D.A + (long)k + (long)k * D.ld;
    g.ldc = D.ld;
    g.strideC = D.strideA;
    g.K = k;
    g.lower = 1;
    g.subtract = 1;
    g.neg_cnt = D.neg_cnt;
    g.neg_idx = D.neg_idx;
    g.strideNeg = D.strideNeg;
    g.state = D.state;
    const double w = (double)ncols, m = (double)(D.MT - k);
    const double fl = (double)k * (w * w + 2.0 * (m - w) * w);
    const int rc2 = launch_gemm_timed(ctx, g, (D.MT - k) / kTile, ncols / kTile, D.batch, fl);
    if (rc2) return rc2;
    D.flops += fl;
    return CHEQ_OK;
}

// x = L^-T v.  v is consumed (its blocks become the updated right-hand sides), the solution goes to x.
int backsolve(Dense& D, int NP, double* v, double* x, long v_stride) {
    cheq_ctx* ctx = D.ctx;
    if (int rcp = prof_begin(ctx, 2)) return rcp;
    static const bool steps = [] { const char* e = getenv("CHEQ_BACKSOLVE_STEPS"); return e && e[0] == '1'; }();
    if (!steps) {
        CU(ctx->chain.ensure((size_t)D.batch * (NP / kTile + 1) * sizeof(unsigned)));
        LAUNCH(launch_backsolve_chain(D.A, D.ld, D.strideA, D.Linv, D.strideLinv, D.sgn, D.strideSgn, NP, v, x,
                                      v_stride, ctx->chain.as<unsigned>(), D.state, D.batch, ctx->cur));
        return prof_end(ctx);
    }
    for (int kb = NP / kTile - 1; kb >= 0; --kb)
        LAUNCH(launch_backsolve_step(D.A, D.ld, D.strideA, D.Linv, D.strideLinv, D.sgn, D.strideSgn, NP, kb, v, x,
                                     v_stride, D.state, D.batch, ctx->cur));
    return prof_end(ctx);
}

// ---- NCCL (run-time binding) ---------------------------------------------------------------------------------
NcclApi g_nccl;
std::mutex g_nccl_mu;

bool nccl_load(std::string& err) {
    std::lock_guard<std::mutex> lock(g_nccl_mu);
    if (g_nccl.handle) return true;
    std::vector<std::string> names;
    if (const char* env = getenv("CHEQ_NCCL_LIB")) names.push_back(env);
    names.push_back("libnccl.so.2");
    names.push_back("libnccl.so");
    void* h = nullptr;
    for (const std::string& nm : names) {
        h = dlopen(nm.c_str(), RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        err = std::string("cannot load NCCL (set CHEQ_NCCL_LIB): ") + (dlerror() ? dlerror() : "");
        return false;
    }
    NcclApi api;
    auto sym = [&](const char* name) {
        void* f = dlsym(h, name);
        if (!f) err = std::string("NCCL symbol missing: ") + name;
        return f;
    };
    api.GetUniqueId = (int (*)(void*))sym("ncclGetUniqueId");
    api.CommInitRank = (int (*)(void**, int, NcclId, int))sym("ncclCommInitRank");
    api.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
    api.Broadcast = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))sym("ncclBroadcast");
    api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))sym("ncclAllReduce");
    api.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))sym("ncclAllGather");
    api.GroupStart = (int (*)())sym("ncclGroupStart");
    api.GroupEnd = (int (*)())sym("ncclGroupEnd");
    api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.Broadcast || !api.AllReduce ||
        !api.AllGather || !api.GroupStart || !api.GroupEnd || !api.GetErrorString)
        return false;
    api.handle = h;
    g_nccl = api;
    return true;
}

#define NC(call)                                                                              \
    do {                                                                                      \
        int r__ = (call);                                                                     \
        if (r__ != 0) {                                                                       \
            ctx->err = std::string(#call) + ": " + g_nccl.GetErrorString(r__);                \
            return CHEQ_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

// Panel plan of a layout for this context's (rank, world); uploads the owned-tile list and the ownership mask.
int ensure_panel_plan(cheq_ctx* ctx, const Layout& L) {
    const int nxt = (L.scf ? L.nxp : L.NP) / kTile, nht = (L.scf ? L.nhp : 0) / kTile;
    PanelPlan& pp = ctx->pplan;
    // With the trailing updates on the tcgen05 path, 8-tile panels (K = 1024) feed it better than 4-tile ones (S20k
    // panel schedule on one GPU: 0.199 s with 4, 0.186 s with 8, 0.205 s with 16) -- as long as every rank still owns
    // at least four panels, otherwise the owner-computes updates go out of balance.
    int tpp = ctx->tiles_per_panel;
    if (!ctx->tiles_per_panel_set && ctx->ozaki_slices > 0 && (long)(nxt + nht) >= 32l * ctx->world) tpp = 8;
    ctx->tpp_now = tpp;
    if (pp.nxt == nxt && pp.nht == nht && pp.tpp == tpp && pp.world == ctx->world && pp.rank == ctx->rank)
        return CHEQ_OK;
    pp.build(nxt, nht, tpp, ctx->world, ctx->rank);
    CU(ctx->own_list.ensure(std::max<size_t>(4, pp.own_tiles.size() * sizeof(int))));
    CU(ctx->col_mask.ensure(std::max<size_t>(4, pp.mask.size())));
    CU(cudaStreamSynchronize(ctx->stream));   // earlier kernels may still read the previous plan's arrays
    if (!pp.own_tiles.empty())
        CU(cudaMemcpy(ctx->own_list.p, pp.own_tiles.data(), pp.own_tiles.size() * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->col_mask.p, pp.mask.data(), pp.mask.size(), cudaMemcpyHostToDevice));
    return CHEQ_OK;
}

// ---- peer-memory data plane ----------------------------------------------------------------------------------
// Layout of the arena for a padded size NP: signal words, then Linv (NP x 128), sgn (NP), negative-pivot counts
// (NP / 128) and lists (NP).  Identical on every rank, so a local offset is also the peer's offset.
// Round 2: + the inboxes of the peer-memory all-reduce of the stale-factor iterations (kernels_krylov.cu): flag words
// at byte 1024 of the signal page (2 parities x 16 ranks), data = 2 parities x 8 rank slots x (NP + 8) doubles.
struct ArenaLayout {
    size_t linv, sgn, neg_cnt, neg_idx, kx_data, kx_slot, bytes;
    static constexpr size_t kx_flags = 1024;
    explicit ArenaLayout(long NP) {
        linv = 4096;
        sgn = linv + (size_t)NP * kTile * 8;
        neg_cnt = sgn + (size_t)NP * 8;
        neg_idx = neg_cnt + (((size_t)(NP / kTile) * 4 + 255) / 256) * 256;
        kx_data = ((neg_idx + (size_t)NP * 4 + 255) / 256) * 256;
        kx_slot = (size_t)NP + 8;
        bytes = kx_data + 2 * (size_t)(kMaxPeers + 1) * kx_slot * 8;
        bytes = std::max<size_t>(bytes, (size_t)8 << 20);   // its own allocation, never a sub-allocation
    }
};

void p2p_close(cheq_ctx* ctx) {
    cheq_ctx::PeerLinks& lk = ctx->links;
    for (int i = 0; i < kMaxPeers; ++i)
        for (int k = 0; k < 2; ++k)
            if (lk.opened[i][k]) {
                cudaIpcCloseMemHandle(lk.opened[i][k]);
                lk.opened[i][k] = nullptr;
            }
    lk.enabled = false;
}

struct IpcRecord {
    cudaIpcMemHandle_t handle[2];
    unsigned long long offset[2];
    int ok;
    int pad;
};

// Start-of-solve rendezvous of a group: a barrier (no rank may push into a peer that is still reading its factor
// from the previous solve) that also (re)establishes the peer mappings when any rank's buffers moved.
int p2p_setup(cheq_ctx* ctx) {
    cheq_ctx::PeerLinks& lk = ctx->links;
    cudaStream_t ms = ctx->stream, ps = ctx->panel_stream;
    const int world = ctx->world, rank = ctx->rank;
    if (!lk.tried) {
        if (const char* env = getenv("CHEQ_DIST_P2P")) ctx->p2p_mode = std::max(0, std::min(3, atoi(env)));
        cudaDriverEntryPointQueryResult q;
        void* fn = nullptr;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            ctx->wait_value32 = (int (*)(cudaStream_t, unsigned long long, unsigned, unsigned))fn;
        fn = nullptr;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            ctx->mem_address_range = (int (*)(unsigned long long*, size_t*, unsigned long long))fn;
        cudaGetLastError();
        if (!ctx->wait_value32 || !ctx->mem_address_range || world - 1 > kMaxPeers) ctx->p2p_mode = 0;
    }
    void* local[2] = {ctx->A.p, ctx->arena.p};
    int changed = (!lk.tried || local[0] != lk.local[0] || local[1] != lk.local[1]) ? 1 : 0;
    lk.tried = true;
    CU(ctx->ipc_stage.ensure(4096 + (size_t)world * sizeof(IpcRecord)));
    int* d_flag = ctx->ipc_stage.as<int>();
    CU(cudaEventRecord(ctx->ev_ready, ms));
    CU(cudaStreamWaitEvent(ps, ctx->ev_ready, 0));
    CU(cudaMemcpyAsync(d_flag, &changed, sizeof(int), cudaMemcpyHostToDevice, ps));
    NC(g_nccl.AllReduce(d_flag, d_flag, 1, kNcclInt, kNcclMax, ctx->comm, ps));
    CU(cudaMemcpyAsync(&changed, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ps));
    CU(cudaStreamSynchronize(ps));
    if (!changed || !ctx->p2p_mode) {
        if (!ctx->p2p_mode) lk.enabled = false;
        return CHEQ_OK;
    }
    // some rank's buffers moved: everybody re-exports and re-imports
    p2p_close(ctx);
    IpcRecord mine{};
    mine.ok = 1;
    for (int k = 0; k < 2; ++k) {
        unsigned long long base = 0;
        size_t size = 0;
        if (ctx->mem_address_range(&base, &size, (unsigned long long)local[k]) != 0 ||
            cudaIpcGetMemHandle(&mine.handle[k], (void*)base) != cudaSuccess) {
            cudaGetLastError();
            mine.ok = 0;
            break;
        }
        mine.offset[k] = (unsigned long long)local[k] - base;
    }
    char* d_all = ctx->ipc_stage.as<char>() + 4096;
    CU(cudaMemcpyAsync(d_all + (size_t)rank * sizeof(IpcRecord), &mine, sizeof(IpcRecord), cudaMemcpyHostToDevice, ps));
    NC(g_nccl.AllGather(d_all + (size_t)rank * sizeof(IpcRecord), d_all, sizeof(IpcRecord), kNcclChar, ctx->comm, ps));
    std::vector<IpcRecord> all(world);
    CU(cudaMemcpyAsync(all.data(), d_all, (size_t)world * sizeof(IpcRecord), cudaMemcpyDeviceToHost, ps));
    CU(cudaStreamSynchronize(ps));
    int ok = 1;
    for (int r = 0; r < world; ++r) ok &= all[r].ok;
    lk.n = 0;
    for (int r = 0; r < world && ok; ++r) {
        if (r == rank) continue;
        const int i = lk.n;
        for (int k = 0; k < 2; ++k) {
            void* mapped = nullptr;
            if (cudaIpcOpenMemHandle(&mapped, all[r].handle[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = 0;
                break;
            }
            lk.opened[i][k] = mapped;
            (k == 0 ? lk.A[i] : lk.arena[i]) = (char*)mapped + all[r].offset[k];
        }
        lk.rank_of[i] = r;
        if (ok) lk.n++;
    }
    // every rank must have succeeded, or nobody uses the peer path
    CU(cudaMemcpyAsync(d_flag, &ok, sizeof(int), cudaMemcpyHostToDevice, ps));
    NC(g_nccl.AllReduce(d_flag, d_flag, 1, kNcclInt, 3 /* ncclMin */, ctx->comm, ps));
    CU(cudaMemcpyAsync(&ok, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ps));
    CU(cudaStreamSynchronize(ps));
    if (!ok) {
        p2p_close(ctx);
        ctx->p2p_mode = 0;
    } else {
        lk.enabled = true;
    }
    lk.local[0] = local[0];
    lk.local[1] = local[1];
    return CHEQ_OK;
}

// Trailing update by the factored panel [j0, j0 + n):  C[:, cols] -= W S W[cols]^T for `count` owned tile columns
// (absolute indices in the device list `list`, host copy `hlist`), rows from tile e down to MT.
int update_columns(Dense& D, int j0, int n, int e, const int* list, const int* hlist, int count) {
    if (count <= 0) return CHEQ_OK;
    cheq_ctx* ctx = D.ctx;
    const int MTt = D.MT / kTile;
    GemmArgs g{};
    g.A = D.A + (long)e * kTile + (long)j0 * D.ld;
    g.lda = D.ld;
    g.strideA = D.strideA;
    g.B = D.A + (long)j0 * D.ld;
    g.ldb = D.ld;
    g.strideB = D.strideA;
    g.C = D.A + (long)e * kTile;
    g.ldc = D.ld;
    g.strideC = D.strideA;
    g.K = n;
    g.lower = 1;
    g.subtract = 1;
    g.neg_cnt = D.neg_cnt + j0 / kTile;
    g.neg_idx = D.neg_idx + (long)j0;
    g.strideNeg = D.strideNeg;
    g.state = D.state;
    g.n_list = list;
    g.row0 = e * kTile;
    double fl = 0.0;
    for (int i = 0; i < count; ++i)
        fl += (double)n * kTile * kTile * (2.0 * (MTt - hlist[i]) - 1.0);   // diagonal tile: symmetric half
    const int rc = launch_gemm_timed(ctx, g, MTt - e, count, D.batch, fl);
    if (rc) return rc;
    D.flops += fl;
    return CHEQ_OK;
}

// Right-looking factorisation of panels [p_begin, p_end) of the plan with one-panel look-ahead:
//   panel stream (high priority): the owner factors panel p (recursive tall-panel kernel sequence), the panel, its
//     diagonal-tile inverses and pivot signs are broadcast to all ranks (NCCL over NVLink), non-owners store it;
//   main stream: every rank updates its own tile columns to the right of the panel; the columns of panel p + 1
//     go first, so that its owner can start factoring while the rest of the update is still running.
// Afterwards every rank holds the complete factor L, the forward-substituted right-hand-side rows and (for the
// columns it owns) the updated trailing matrix.  One GPU (world == 1): the same schedule without broadcasts.
int factor_blocked(Dense& D, int p_begin, int p_end) {
    cheq_ctx* ctx = D.ctx;
    const PanelPlan& pp = ctx->pplan;
    cudaStream_t ms = ctx->stream, ps = ctx->panel_stream;
    const int* d_list = ctx->own_list.as<int>();
    // nothing on the panel stream (factorisation, received panels) may overtake what the main stream has queued
    CU(cudaEventRecord(ctx->ev_ready, ms));
    CU(cudaStreamWaitEvent(ps, ctx->ev_ready, 0));
    for (int p = p_begin; p < p_end; ++p) {
        const int j0 = pp.begin[p] * kTile, n = (pp.end[p] - pp.begin[p]) * kTile;
        const long rows = D.MT - j0;
        const bool mine = pp.owner[p] == ctx->rank;
        const bool push = ctx->world > 1 && ctx->push_now;
        if (mine) {
            CU(cudaStreamWaitEvent(ps, ctx->ev_ready, 0));
            const SplitPlan* saved = D.plan;
            D.plan = nullptr;   // panels are a few tiles wide: plain halving
            ctx->cur = ps;
            D.push = push && !ctx->push_dma;
            D.push_dma = push && ctx->push_dma;
            // the bulk stream starts from what the panel stream sees now (all earlier updates of these columns)
            if (ctx->panel_flat && D.batch == 1 && n / kTile <= 64) {
                CU(cudaEventRecord(ctx->ev_leaf, ps));
                CU(cudaStreamWaitEvent(ctx->bulk_stream, ctx->ev_leaf, 0));
            }
            const int rc = (ctx->panel_flat && D.batch == 1 && n / kTile <= 64) ? factor_panel_flat(D, j0, n)
                                                                                  : factor_panel(D, j0, n);
            D.push = false;
            D.push_dma = false;
            ctx->cur = ms;
            D.plan = saved;
            if (rc) return rc;
        }
        if (push) {
            // data plane = peer stores issued by the owner's kernels; what is left is one flag per peer
            const unsigned seq = ++ctx->links.seq;
            if (mine) {
                unsigned* words[kMaxPeers];
                for (int i = 0; i < ctx->links.n; ++i) words[i] = (unsigned*)ctx->links.arena[i] + ctx->rank;
                CU(cudaEventRecord(ctx->ev_bcast, ps));
                CU(cudaStreamWaitEvent(ms, ctx->ev_bcast, 0));   // this rank's update needs only its own kernels
                if (ctx->push_dma) {
                    // the panel's small outputs, then the flag, behind the tile-column copies on the copy stream
                    cudaStream_t cs = ctx->copy_stream;
                    CU(cudaStreamWaitEvent(cs, ctx->ev_bcast, 0));
                    const char* ar = ctx->arena.as<char>();
                    const double* Lk = D.Linv + (long)(j0 / kTile) * kTile * kTile;
                    for (int i = 0; i < ctx->links.n; ++i) {
                        char* pa = ctx->links.arena[i];
                        CU(cudaMemcpyAsync(pa + ((const char*)Lk - ar), Lk, (size_t)n * kTile * 8,
                                           cudaMemcpyDeviceToDevice, cs));
                        CU(cudaMemcpyAsync(pa + ((const char*)(D.sgn + j0) - ar), D.sgn + j0, (size_t)n * 8,
                                           cudaMemcpyDeviceToDevice, cs));
                        CU(cudaMemcpyAsync(pa + ((const char*)(D.neg_cnt + j0 / kTile) - ar), D.neg_cnt + j0 / kTile,
                                           (size_t)(n / kTile) * 4, cudaMemcpyDeviceToDevice, cs));
                        CU(cudaMemcpyAsync(pa + ((const char*)(D.neg_idx + j0) - ar), D.neg_idx + j0, (size_t)n * 4,
                                           cudaMemcpyDeviceToDevice, cs));
                    }
                    LAUNCH(launch_signal_peers(words, ctx->links.n, seq, cs));
                } else {
                    LAUNCH(launch_signal_peers(words, ctx->links.n, seq, ps));
                }
            } else {
                const unsigned long long addr = (unsigned long long)(ctx->arena.as<unsigned>() + pp.owner[p]);
                if (ctx->wait_value32(ms, addr, seq, 0 /* CU_STREAM_WAIT_VALUE_GEQ */) != 0)
                    return fail(ctx, CHEQ_ERR_CUDA, "cuStreamWaitValue32 failed");
            }
        } else {
            if (ctx->world > 1) {
                auto t0 = Clock::now();
                double* stage = ctx->stage.as<double>();
                double* Ap = D.A + (long)j0 + (long)j0 * D.ld;
                if (mine)
                    CU(cudaMemcpy2DAsync(stage, rows * 8, Ap, D.ld * 8, rows * 8, n, cudaMemcpyDeviceToDevice, ps));
                double* Lk = D.Linv + (long)(j0 / kTile) * kTile * kTile;
                const int root = pp.owner[p];
                NC(g_nccl.GroupStart());
                NC(g_nccl.Broadcast(stage, stage, (size_t)rows * n, kNcclDouble, root, ctx->comm, ps));
                NC(g_nccl.Broadcast(Lk, Lk, (size_t)n * kTile, kNcclDouble, root, ctx->comm, ps));
                NC(g_nccl.Broadcast(D.sgn + j0, D.sgn + j0, (size_t)n, kNcclDouble, root, ctx->comm, ps));
                NC(g_nccl.Broadcast(D.neg_cnt + j0 / kTile, D.neg_cnt + j0 / kTile, (size_t)n / kTile, kNcclInt, root,
                                    ctx->comm, ps));
                NC(g_nccl.Broadcast(D.neg_idx + j0, D.neg_idx + j0, (size_t)n, kNcclInt, root, ctx->comm, ps));
                NC(g_nccl.GroupEnd());
                if (!mine)
                    CU(cudaMemcpy2DAsync(Ap, D.ld * 8, stage, rows * 8, rows * 8, n, cudaMemcpyDeviceToDevice, ps));
                ctx->launches += 1;
                ctx->ms_comm_host += ms_since(t0);
            }
            CU(cudaEventRecord(ctx->ev_bcast, ps));
            CU(cudaStreamWaitEvent(ms, ctx->ev_bcast, 0));
        }
        // trailing update of the owned tile columns right of the panel
        const int e = pp.end[p];
        const int pos = (int)(std::lower_bound(pp.own_tiles.begin(), pp.own_tiles.end(), e) - pp.own_tiles.begin());
        const int cnt = (int)pp.own_tiles.size() - pos;
        const bool next_mine = (p + 1 < p_end) && pp.owner[p + 1] == ctx->rank;
        int head = 0;
        if (next_mine) {
            head = pp.end[p + 1] - pp.begin[p + 1];
            const int rc = update_columns(D, j0, n, e, d_list + pos, pp.own_tiles.data() + pos, head);
            if (rc) return rc;
            CU(cudaEventRecord(ctx->ev_ready, ms));
        }
        ctx->oz_same_a = head > 0;   // the panel's rows were sliced for the head update already
        const int rc = update_columns(D, j0, n, e, d_list + pos + head, pp.own_tiles.data() + pos + head, cnt - head);
        ctx->oz_same_a = false;
        if (rc) return rc;
    }
    if (ctx->world > 1 && ctx->push_now && ctx->push_dma) {
        // the copies read this rank's finished panels: nothing later on the main stream may overwrite them first
        CU(cudaEventRecord(ctx->ev_copy, ctx->copy_stream));
        CU(cudaStreamWaitEvent(ms, ctx->ev_copy, 0));
    }
    if (ctx->world > 1) {
        // a pivot breakdown seen by one owner must stop every rank
        int* info = (int*)((char*)D.state + offsetof(ScfState, info));
        NC(g_nccl.AllReduce(info, info, 1, kNcclInt, kNcclMax, ctx->comm, ps));
        CU(cudaEventRecord(ctx->ev_bcast, ps));
        CU(cudaStreamWaitEvent(ms, ctx->ev_bcast, 0));
    }
    return CHEQ_OK;
}


// ---- stale-factor SCF iterations: host side (kernels in kernels_krylov.cu) ---------------------------------------
// Everything below works on the hydrogen block of ONE frame: n = nhp unknowns plus the border unknown mu.
struct Krylov {
    static constexpr int kMaxBasis = 48;        // GMRES restart length
    static constexpr int kMaxSteps = 96;        // operator applications per SCF iteration before giving up
    int nhp = 0, nt = 0, len = 0;               // len = nhp + 1 entries per vector
    long vs = 0;                                // vector stride (doubles)
    double *R = nullptr, *partN = nullptr, *partT = nullptr;
    float* Rf = nullptr;                        // fp32, TILE-MAJOR copy of the stored tiles of R used by the preconditioner
    double* S0t = nullptr;                      // tile-major copy of this rank's tile columns of S0 (lower triangle)
    const long long *baseR = nullptr, *baseS = nullptr;   // device: first-tile offsets per tile column (tile_matvec_kernel)
    bool s0t_ready = false;
    double *tc = nullptr, *t1 = nullptr, *ws = nullptr, *d_cur = nullptr, *d_prev = nullptr, *d_frozen = nullptr,
           *sol = nullptr, *r = nullptr, *z = nullptr, *tmp1 = nullptr, *tmp2 = nullptr, *w = nullptr, *V = nullptr;
    KrylovScalars* sc = nullptr;
    const double* S0 = nullptr;
    long ld0 = 0;
    const double* sgn_h = nullptr;
    bool ready = false;                         // buffers allocated, per-geometry scalars set
    bool frozen = false;                        // R, ws, den describe the factor currently in the trapezoid
    double bytes = 0.0;
    // group of GPUs: S0 is held by tile columns (col_mask: the columns this rank owns), R by a contiguous chunk of
    // tile rows [ra, rb) sized for equal work in its construction; partial products are summed with ncclAllReduce
    bool dist = false;
    const uint8_t* col_mask = nullptr;
    int ra = 0, rb = 0;
    bool p2p = false;                           // all-reduces go through the peer-mapped arenas
    size_t kx_data = 0, kx_slot = 0;            // ArenaLayout of the current solve
};

// sums a vector over the ranks of the group (identical result on every rank): through the peer-mapped arenas when
// the NVLink mappings are up (one push + one combine kernel, ~10 us), else ncclAllReduce (~55 us on 8 GPUs)
int krylov_allreduce_n(cheq_ctx* ctx, const Krylov& K, double* v, int len) {
    if (!K.dist) return CHEQ_OK;
    if (K.p2p) {
        const cheq_ctx::PeerLinks& lk = ctx->links;
        P2PAllreduce pa{};
        pa.world = ctx->world;
        pa.rank = ctx->rank;
        pa.slot = (long)K.kx_slot;
        pa.seq = ++ctx->kx_seq;
        const size_t par = pa.seq & 1u;
        for (int d = 0; d < ctx->world; ++d) {
            char* base = nullptr;
            if (d == ctx->rank) base = ctx->arena.as<char>();
            else
                for (int i = 0; i < lk.n; ++i)
                    if (lk.rank_of[i] == d) base = lk.arena[i];
            if (!base) return fail(ctx, CHEQ_ERR_CUDA, "peer arena of a rank is not mapped");
            pa.inbox[d] = (double*)(base + K.kx_data) + par * (size_t)(kMaxPeers + 1) * K.kx_slot;
            pa.flags[d] = (unsigned*)(base + ArenaLayout::kx_flags) + par * 16;
        }
        LAUNCH(launch_p2p_allreduce(v, v, len, pa, ctx->kx_local.as<unsigned>(), ctx->kx_local.as<int>() + 1, ctx->stream));
        ctx->launches++;
        return CHEQ_OK;
    }
    NC(g_nccl.AllReduce(v, v, (size_t)len, kNcclDouble, 0 /* ncclSum */, ctx->comm, ctx->stream));
    return CHEQ_OK;
}
int krylov_allreduce(cheq_ctx* ctx, const Krylov& K, double* v) { return krylov_allreduce_n(ctx, K, v, K.len); }

int krylov_alloc(cheq_ctx* ctx, Krylov& K, int nhp, bool dist, const uint8_t* col_mask_host) {
    K.nhp = nhp;
    K.nt = nhp / kTile;
    K.len = nhp + 1;
    K.vs = nhp + 8;
    K.dist = dist;
    K.ra = 0;
    K.rb = K.nt;
    if (dist) {
        // row tile i of R costs ~ (nt - i)^2 in the construction: chunk boundaries at equal cumulative cost
        auto bound = [&](int k) {
            const double f = 1.0 - std::cbrt(1.0 - (double)k / ctx->world);
            return std::min(K.nt, std::max(0, (int)std::lround(K.nt * f)));
        };
        K.ra = bound(ctx->rank);
        K.rb = (ctx->rank == ctx->world - 1) ? K.nt : bound(ctx->rank + 1);
        if (K.rb <= K.ra) K.ra = K.rb = K.nt + 1;   // no rows here: a range that matches no tile (0, 0 would mean "all")
    }
    const size_t r_rows = (size_t)std::max(1, K.rb - K.ra) * kTile;
    CU(ctx->kr_R.ensure(r_rows * nhp * 8));
    // tile-major copies: offsets of the first stored tile of every tile column (R: rows [ra, rb) on/above the
    // diagonal, fp32; S0: this rank's columns on/below the diagonal, fp64)
    K.Rf = nullptr;
    K.S0t = nullptr;
    K.s0t_ready = false;
    if (ctx->krylov_f32) {
        std::vector<long long> base(2 * (size_t)K.nt, -1);
        long long tr = 0, ts = 0;
        for (int j = 0; j < K.nt; ++j) {
            const int cnt_r = std::min(j, K.rb - 1) - K.ra + 1;
            if (cnt_r > 0) {
                base[j] = tr * (long long)kTile * kTile;
                tr += cnt_r;
            }
            if (!col_mask_host || col_mask_host[j]) {
                base[K.nt + j] = ts * (long long)kTile * kTile;
                ts += K.nt - j;
            }
        }
        CU(ctx->kr_Rf.ensure(std::max<size_t>(16, (size_t)tr * kTile * kTile * 4)));
        CU(ctx->kr_S0t.ensure(std::max<size_t>(16, (size_t)ts * kTile * kTile * 8)));
        CU(ctx->kr_tb.ensure(base.size() * sizeof(long long)));
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaMemcpy(ctx->kr_tb.p, base.data(), base.size() * sizeof(long long), cudaMemcpyHostToDevice));
        K.Rf = ctx->kr_Rf.as<float>();
        K.S0t = ctx->kr_S0t.as<double>();
        K.baseR = ctx->kr_tb.as<long long>();
        K.baseS = K.baseR + K.nt;
    }
    CU(ctx->kr_part.ensure((size_t)2 * K.nt * K.nt * kTile * 8));
    const size_t nvec = 12 + Krylov::kMaxBasis + 1;
    CU(ctx->kr_vec.ensure(nvec * K.vs * 8));
    CU(ctx->kr_sc.ensure(sizeof(KrylovScalars)));
    if (!ctx->h_ksc) CU(cudaMallocHost((void**)&ctx->h_ksc, sizeof(KrylovScalars)));
    // R is addressed with GLOBAL tile-row indices: the base is shifted up by the rows this rank does not hold
    K.R = ctx->kr_R.as<double>() - (long)K.ra * kTile;
    K.partN = ctx->kr_part.as<double>();
    K.partT = K.partN + (size_t)K.nt * K.nt * kTile;
    double* v = ctx->kr_vec.as<double>();
    double** slots[] = {&K.tc, &K.t1, &K.ws, &K.d_cur, &K.d_prev, &K.d_frozen, &K.sol, &K.r, &K.z, &K.tmp1, &K.tmp2, &K.w};
    for (double** sl : slots) {
        *sl = v;
        v += K.vs;
    }
    K.V = v;
    K.sc = (KrylovScalars*)ctx->kr_sc.p;
    CU(cudaMemsetAsync(ctx->kr_vec.p, 0, nvec * K.vs * 8, ctx->stream));
    CU(cudaMemsetAsync(ctx->kr_sc.p, 0, sizeof(KrylovScalars), ctx->stream));
    return CHEQ_OK;
}

// out = S_s^-1 in = R S R^T in   (n entries; two passes over the upper triangle of R)
int krylov_apply_sinv(cheq_ctx* ctx, Krylov& K, const double* in, double* out) {
    cudaStream_t st = ctx->stream;
    TileMatvecArgs a{};
    a.M = K.Rf ? (const void*)K.Rf : (const void*)K.R;
    a.f32 = K.Rf ? 1 : 0;
    a.col_base = K.Rf ? K.baseR : nullptr;
    a.ld = (long)(K.rb - K.ra) * kTile;
    a.nt = K.nt;
    a.upper = 1;
    a.xT = in;
    a.partN = K.partN;
    a.partT = K.partT;
    a.row_begin = K.ra;
    a.row_end = K.rb;
    LAUNCH(launch_tile_matvec(a, false, true, st));
    TileReduceArgs r{};
    r.nt = K.nt;
    r.upper = 1;
    r.op = 0;
    r.row_begin = K.ra;
    r.row_end = K.rb;
    r.partT = K.partT;
    r.out = K.tmp1;
    LAUNCH(launch_tile_matvec_reduce(r, st));
    if (K.dist) {
        CU(cudaMemsetAsync(K.tmp1 + K.nhp, 0, sizeof(double), st));
        if (int rc = krylov_allreduce(ctx, K, K.tmp1)) return rc;
    }
    a.xT = nullptr;
    a.xN = K.tmp1;
    a.scaleN = K.sgn_h;
    LAUNCH(launch_tile_matvec(a, true, false, st));
    r.partT = nullptr;
    r.partN = K.partN;
    r.out = out;
    LAUNCH(launch_tile_matvec_reduce(r, st));
    if (K.dist) {
        CU(cudaMemsetAsync(out + K.nhp, 0, sizeof(double), st));
        if (int rc = krylov_allreduce(ctx, K, out)) return rc;
    }
    K.bytes += (K.Rf ? 4.0 : 8.0) * (double)K.nhp * (K.nhp + 1.0) / (K.dist ? ctx->world : 1);
    return CHEQ_OK;
}

// out = K_k v (op 1) or rhs - K_k v (op 2), K_k = [[S0 + D_k, t1], [t1^T, -g]]   (one pass over the lower triangle of S0)
int krylov_apply_K(cheq_ctx* ctx, Krylov& K, const double* v, double* out, int op) {
    cudaStream_t st = ctx->stream;
    TileMatvecArgs a{};
    a.M = K.s0t_ready ? K.S0t : K.S0;
    a.col_base = K.s0t_ready ? K.baseS : nullptr;
    a.ld = K.ld0;
    a.nt = K.nt;
    a.upper = 0;
    a.sym = 1;
    a.xN = v;
    a.xT = v;
    a.partN = K.partN;
    a.partT = K.partT;
    a.col_mask = K.col_mask;
    LAUNCH(launch_tile_matvec(a, true, true, st));
    TileReduceArgs r{};
    r.nt = K.nt;
    r.upper = 0;
    r.op = op;
    r.partN = K.partN;
    r.partT = K.partT;
    r.col_mask = K.col_mask;
    r.border_owner = (!K.dist || ctx->rank == 0) ? 1 : 0;
    r.v = v;
    r.d = K.d_cur;
    r.t1 = K.t1;
    r.rhs = K.tc;
    r.sc = K.sc;
    r.out = out;
    LAUNCH(launch_tile_matvec_reduce(r, st));
    if (int rc = krylov_allreduce(ctx, K, out)) return rc;
    K.bytes += 4.0 * (double)K.nhp * (K.nhp + 1.0) / (K.dist ? ctx->world : 1);
    return CHEQ_OK;
}

// out = K_s^-1 in  (frozen factor + border elimination)
int krylov_apply_precond(cheq_ctx* ctx, Krylov& K, const double* in, double* out) {
    int rc = krylov_apply_sinv(ctx, K, in, K.tmp2);
    if (rc) return rc;
    LAUNCH(launch_precond_border(K.tmp2, in, K.t1, K.ws, K.sc, out, K.nhp, ctx->stream));
    return CHEQ_OK;
}

// R = L_hh^-T S for the hydrogen factor currently in the trapezoid: right-hand triangular solve applied to the rows
// of an identity, with the same kernel sequence as the rows below a panel in factor_panel (leaf: multiply by
// M^T = (S L_jj^-1)^T; update: C -= A S B^T), restricted to the rows that can be non-zero.  n_h^3 / 3 flops.
int krylov_build_R(Dense& D, Krylov& K, int h0, int j0, int n) {
    cheq_ctx* ctx = D.ctx;
    const long ldr = (long)(K.rb - K.ra) * kTile;      // this rank's rows [ra, rb) of R (global row indices, shifted base)
    const long r0 = (long)K.ra * kTile;
    if (n == kTile) {
        const int mt = std::min(K.rb, (j0 + kTile) / kTile) - K.ra;
        if (mt <= 0) return CHEQ_OK;
        GemmArgs g{};
        g.A = K.R + r0 + (long)j0 * ldr;
        g.lda = ldr;
        g.B = D.Linv + (long)((h0 + j0) / kTile) * kTile * kTile;
        g.ldb = kTile;
        g.C = K.R + r0 + (long)j0 * ldr;
        g.ldc = ldr;
        g.K = kTile;
        const double fl = 2.0 * mt * (double)kTile * kTile * kTile;
        D.flops += fl;
        return launch_gemm_timed(ctx, g, mt, 1, 1, fl);
    }
    const int n1 = ((n / kTile) / 2) * kTile;
    int rc = krylov_build_R(D, K, h0, j0, n1);
    if (rc) return rc;
    const int mt_u = std::min(K.rb, (j0 + n1) / kTile) - K.ra;
    if (mt_u > 0) {
        GemmArgs g{};
        g.A = K.R + r0 + (long)j0 * ldr;
        g.lda = ldr;
        g.B = D.A + (long)(h0 + j0 + n1) + (long)(h0 + j0) * D.ld;
        g.ldb = D.ld;
        g.C = K.R + r0 + (long)(j0 + n1) * ldr;
        g.ldc = ldr;
        g.K = n1;
        g.subtract = 1;
        g.neg_cnt = D.neg_cnt + (h0 + j0) / kTile;
        g.neg_idx = D.neg_idx + (long)(h0 + j0);
        g.strideNeg = D.strideNeg;
        const int nt = (n - n1) / kTile;
        const double fl = 2.0 * (double)mt_u * kTile * (double)(n - n1) * (double)n1;
        D.flops += fl;
        rc = launch_gemm_timed(ctx, g, mt_u, nt, 1, fl);
        if (rc) return rc;
    }
    return krylov_build_R(D, K, h0, j0 + n1, n - n1);
}

// Freeze the factor that is in the trapezoid (it belongs to the diagonal shift `d_of_factor`).
int krylov_freeze(cheq_ctx* ctx, Dense& D, Krylov& K, int nxp, const double* d_of_factor) {
    cudaStream_t st = ctx->stream;
    LAUNCH(launch_set_identity_tiles(K.R + (long)K.ra * kTile, (long)(K.rb - K.ra) * kTile, K.nt, K.ra, K.rb, st));
    if (ctx->ozaki_slices > 0 && ctx->ozaki_slices_r > 0) ctx->ozaki_now = ctx->ozaki_slices_r;
    int rc = krylov_build_R(D, K, nxp, 0, K.nhp);
    ctx->ozaki_now = ctx->ozaki_slices;
    if (rc) return rc;
    CU(cudaMemcpyAsync(K.d_frozen, d_of_factor, (size_t)K.nhp * 8, cudaMemcpyDeviceToDevice, st));
    if (K.Rf) {
        LAUNCH(launch_pack_tiles(K.R, (long)(K.rb - K.ra) * kTile, K.Rf, 1, K.baseR, K.nt, 1, K.ra, K.rb, nullptr, st));
        if (!K.s0t_ready) {
            LAUNCH(launch_pack_tiles(K.S0, K.ld0, K.S0t, 0, K.baseS, K.nt, 0, 0, K.nt, K.col_mask, st));
            K.s0t_ready = true;
        }
    }
    rc = krylov_apply_sinv(ctx, K, K.t1, K.ws);
    if (rc) return rc;
    LAUNCH(launch_precond_den(K.t1, K.ws, K.sc, K.nhp, st));
    K.frozen = true;
    return CHEQ_OK;
}

// Restarted GMRES with iterative refinement, host-driven: the device does the operator / preconditioner applications
// and the Gram-Schmidt arithmetic (classical, twice; fixed summation order), the host keeps the small Hessenberg
// least-squares problem and reads one column of coefficients per step.  Left-preconditioned: the monitored quantity
// |P (b - K x)| estimates the error itself when P ~ K^-1.  Every restart begins from the TRUE residual, and the solve
// only ends when that (not the recurrence's estimate) is below tol * |sol_0|.
struct GmresWork {
    int len = 0;                 // entries per vector (unknowns + border)
    long vs = 0;                 // vector stride in doubles
    int m = 48;                  // restart length (<= kGmresMaxBasis)
    int max_steps = 96;          // operator applications before giving up
    double *r = nullptr, *z = nullptr, *w = nullptr, *V = nullptr;   // device: residual, P r, work vector, basis [m][vs]
    KrylovScalars* sc = nullptr; // device
    double stall_accept = 1e3;   // a stalled restart is accepted if |P r| <= stall_accept * tol * scale
    // Strong preconditioners (frozen exact factor): a cycle that ends because the recurrence's residual estimate fell
    // below trust_margin * tol * scale is accepted without recomputing the true residual (the Arnoldi relation holds to
    // rounding with twice-applied Gram-Schmidt; one operator application saved per SCF iteration).  0: always verify.
    double trust_margin = 0.0;
};

template <class ApplyK, class ApplyP>
int gmres_solve(cheq_ctx* ctx, const GmresWork& G, ApplyK&& apply_K, ApplyP&& apply_P, const double* rhs, double* sol,
                double tol, int* steps, bool* ok) {
    cudaStream_t st = ctx->stream;
    KrylovScalars* hs = ctx->h_ksc;
    const int m = G.m;
    constexpr int kNormSlot = kGmresMaxBasis + 1;
    std::vector<double> H((size_t)(m + 1) * m), cs(m), sn(m), gv(m + 1), y(m);
    *steps = 0;
    *ok = false;
    static const bool debug = [] { const char* e = getenv("CHEQ_KRYLOV_DEBUG"); return e && e[0] == '1'; }();
    const auto t_begin = Clock::now();
    double scale = 0.0, beta_prev = 0.0;
    for (int cycle = 0; cycle < 64; ++cycle) {
        // true residual, preconditioned: z = P (b - K sol)
        int rc = apply_K(sol, G.r);
        if (rc) return rc;
        LAUNCH(launch_residual(rhs, G.r, G.r, G.len, st));
        rc = apply_P(G.r, G.z);
        if (rc) return rc;
        *steps += 1;
        LAUNCH(launch_multi_dot(nullptr, 0, G.z, G.len, G.sc->dots, 0, true, st));
        if (cycle == 0) LAUNCH(launch_multi_dot(nullptr, 0, sol, G.len, G.sc->dots + 1, 0, true, st));
        CU(cudaMemcpyAsync(hs->dots, G.sc->dots, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        const double beta = std::sqrt(hs->dots[0]);
        if (cycle == 0) {
            scale = std::sqrt(hs->dots[1]);
            if (!(scale > 0.0)) scale = std::max(beta, 1e-300);   // zero initial guess: measure against the first correction
        }
        if (debug && ctx->rank == 0)
            fprintf(stderr, "[gmres %.3f ms] cycle %d steps %d  |P r| = %.3e  scale = %.3e\n", ms_since(t_begin), cycle, *steps, beta, scale);
        if (!std::isfinite(beta)) return CHEQ_OK;                        // not ok
        if (beta <= tol * scale) {
            *ok = true;
            return CHEQ_OK;
        }
        if (cycle > 0 && beta > 0.25 * beta_prev) {
            // rounding floor of this system (or a stalled restart): accept only if it is far below the parity budget
            *ok = beta <= G.stall_accept * tol * scale;
            return CHEQ_OK;
        }
        if (*steps >= G.max_steps) return CHEQ_OK;
        beta_prev = beta;
        LAUNCH(launch_scale_by_rnorm(G.z, G.V, G.len, G.sc->dots, st));
        gv[0] = beta;
        int k = 0;
        for (int j = 0; j < m; ++j) {
            const double* vj = G.V + (long)j * G.vs;
            rc = apply_K(vj, G.r);
            if (rc) return rc;
            rc = apply_P(G.r, G.w);
            if (rc) return rc;
            *steps += 1;
            // classical Gram-Schmidt, twice
            LAUNCH(launch_multi_dot(G.V, G.vs, G.w, G.len, G.sc->dots, j + 1, false, st));
            LAUNCH(launch_gs_update(G.V, G.vs, G.w, G.len, G.sc->dots, G.sc->hsum, j + 1, true, st));
            LAUNCH(launch_multi_dot(G.V, G.vs, G.w, G.len, G.sc->dots, j + 1, false, st));
            LAUNCH(launch_gs_update(G.V, G.vs, G.w, G.len, G.sc->dots, G.sc->hsum, j + 1, false, st));
            LAUNCH(launch_multi_dot(nullptr, 0, G.w, G.len, G.sc->dots + kNormSlot, 0, true, st));
            CU(cudaMemcpyAsync(hs->hsum, G.sc->hsum, (j + 1) * sizeof(double), cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(hs->dots + kNormSlot, G.sc->dots + kNormSlot, sizeof(double), cudaMemcpyDeviceToHost, st));
            if (j + 1 < m)
                LAUNCH(launch_scale_by_rnorm(G.w, G.V + (long)(j + 1) * G.vs, G.len, G.sc->dots + kNormSlot, st));
            CU(cudaStreamSynchronize(st));
            for (int i = 0; i <= j; ++i) H[(size_t)i * m + j] = hs->hsum[i];
            const double hn = std::sqrt(std::max(0.0, hs->dots[kNormSlot]));
            for (int i = 0; i < j; ++i) {
                const double a = cs[i] * H[(size_t)i * m + j] + sn[i] * H[(size_t)(i + 1) * m + j];
                H[(size_t)(i + 1) * m + j] = -sn[i] * H[(size_t)i * m + j] + cs[i] * H[(size_t)(i + 1) * m + j];
                H[(size_t)i * m + j] = a;
            }
            const double rr = std::hypot(H[(size_t)j * m + j], hn);
            if (!(rr > 0.0) || !std::isfinite(rr)) {
                k = j;
                break;
            }
            cs[j] = H[(size_t)j * m + j] / rr;
            sn[j] = hn / rr;
            H[(size_t)j * m + j] = rr;
            gv[j + 1] = -sn[j] * gv[j];
            gv[j] = cs[j] * gv[j];
            k = j + 1;
            if (debug && ctx->rank == 0)
                fprintf(stderr, "[gmres %.3f ms]   step %d  estimate %.3e\n", ms_since(t_begin), j, std::fabs(gv[j + 1]));
            if (std::fabs(gv[j + 1]) <= (G.trust_margin > 0.0 ? G.trust_margin : 0.3) * tol * scale || *steps >= G.max_steps)
                break;
        }
        if (k == 0) return CHEQ_OK;
        for (int i = k - 1; i >= 0; --i) {
            double sacc = gv[i];
            for (int c = i + 1; c < k; ++c) sacc -= H[(size_t)i * m + c] * y[c];
            y[i] = sacc / H[(size_t)i * m + i];
        }
        for (int i = 0; i < k; ++i) hs->y[i] = y[i];
        CU(cudaMemcpyAsync(G.sc->y, hs->y, k * sizeof(double), cudaMemcpyHostToDevice, st));
        LAUNCH(launch_combine(G.V, G.vs, sol, G.len, G.sc->y, k, st));
        if (G.trust_margin > 0.0 && std::fabs(gv[k]) <= G.trust_margin * tol * scale) {
            *ok = true;
            return CHEQ_OK;
        }
    }
    return CHEQ_OK;
}

// Solves K_k sol = [t_c; Q - h] with GMRES preconditioned by the frozen K_s^-1.  `sol` holds the previous
// iteration's solution on entry.  Returns the operator applications in *steps; *ok = false if it did not get there
// (the caller then refactors).
int krylov_solve(cheq_ctx* ctx, Krylov& K, double tol, int* steps, bool* ok) {
    GmresWork G;
    G.len = K.len;
    G.vs = K.vs;
    G.m = Krylov::kMaxBasis;
    G.max_steps = Krylov::kMaxSteps;
    G.r = K.r;
    G.z = K.z;
    G.w = K.w;
    G.V = K.V;
    G.sc = K.sc;
    G.trust_margin = 0.1;
    return gmres_solve(
        ctx, G, [&](const double* v, double* out) { return krylov_apply_K(ctx, K, v, out, 1); },
        [&](const double* in, double* out) { return krylov_apply_precond(ctx, K, in, out); }, K.tc, K.sol, tol, steps, ok);
}

int ensure_host_state(cheq_ctx* ctx, size_t batch) {
    if (batch <= ctx->h_state_cap) return CHEQ_OK;
    if (ctx->h_state) cudaFreeHost(ctx->h_state);
    ctx->h_state = nullptr;
    CU(cudaMallocHost((void**)&ctx->h_state, batch * sizeof(ScfState)));
    ctx->h_state_cap = batch;
    return CHEQ_OK;
}

struct SolveIO {
    bool device_ptrs = false;
    long n = 0;
    long frames = 1;
    const double *xyz = nullptr, *chi = nullptr, *hard = nullptr, *radius = nullptr;
    const uint8_t* nq = nullptr;
    double total_charge = 0.0;
    const cheq_options* opt = nullptr;
    const cheq_external* ext = nullptr;
    double* charges_out = nullptr;
    double* potential_out = nullptr;
    uint32_t* iterations_out = nullptr;
    double* delta_out = nullptr;
    int32_t* status_out = nullptr;   // batch only
    // true only for cheq_solve / cheq_solve_device: on a context that joined a group these calls are COLLECTIVE
    // (every rank calls them with identical inputs).  cheq_solve_batch and the component entry points are always
    // local, whatever the chunk size: a one-frame chunk of a batch must never enter the group's factorisation.
    bool collective = false;
    int path = 0;   // 0: dense direct solve, matrix-free only if the matrix cannot fit; 1: dense; 2: matrix-free
};

int validate_options(cheq_ctx* ctx, const cheq_options* o) {
    if (!o) return fail(ctx, CHEQ_ERR_INVALID_ARGUMENT, "options is NULL");
    if (o->basis > 1) return fail(ctx, CHEQ_ERR_INVALID_ARGUMENT, "unknown basis");
    if (o->damping_kind > 2) return fail(ctx, CHEQ_ERR_INVALID_ARGUMENT, "unknown damping kind");
    return CHEQ_OK;
}

bool external_is_empty(const cheq_external* e) {
    // types.rs:281-286
    return !e || (e->num_point_charges == 0 && e->field[0] == 0.0 && e->field[1] == 0.0 && e->field[2] == 0.0);
}

// Uploads inputs, builds layout + tables, gathers to internal order, runs v_ext and the assembly.
// On return the matrix and right-hand-side rows of every frame of the chunk are in ctx->A.
struct Prepared {
    Layout L;
    PairTableView view{};
    int blob_bytes = 0;
    double rcut2_max = INFINITY;
    long strideA = 0;
    int batch = 1;
    double max_hardness = 0.0;   // scale of the matrix diagonal (host pointers only; 0: unknown)
};

// Which factorisation schedule a solve uses.  Right-looking panels with look-ahead are required for a group.  On
// one GPU they were measured 1.3 % faster than the recursive schedule for S20k (0.754 vs 0.764 s, profiles/
// r1_bench_v3*.json) but run the dominant GEMM at K = 512 (21 instead of 28 TFLOP/s) and lose beyond ~40k rows,
// so the recursive kernel sequence stays the single-GPU default; mode 1 selects panels explicitly.
// CHEQ_FACTOR_MODE / cheq_set_factor_mode: 0 automatic, 1 always panels, 2 never panels on one GPU.
bool use_blocked(const cheq_ctx* ctx, int batch, long NP) {
    (void)NP;
    if (batch != 1) return false;
    // ctx->world is 1 for the duration of a non-collective call on a joined context (LocalCallGuard)
    return ctx->world > 1 || ctx->factor_mode == 1;
}

// A context that joined a group runs its non-collective calls (batches, component entry points) as a single GPU:
// rank / world are masked for the duration of the call (the caller holds ctx->mu).
struct LocalCallGuard {
    cheq_ctx* ctx;
    int rank, world;
    LocalCallGuard(cheq_ctx* c, bool collective) : ctx(c), rank(c->rank), world(c->world) {
        if (!collective) {
            ctx->rank = 0;
            ctx->world = 1;
        }
    }
    ~LocalCallGuard() {
        ctx->rank = rank;
        ctx->world = world;
    }
};

int prepare_and_assemble(cheq_ctx* ctx, const SolveIO& io, long frame0, int batch, bool hydrogen_scf,
                         Prepared& P, bool topology_ready, bool all_columns = false) {
    const long n = io.n;
    const cheq_options& o = *io.opt;
    cudaStream_t st = ctx->stream;
    const bool has_ext = !external_is_empty(io.ext);
    const long m = has_ext ? (long)io.ext->num_point_charges : 0;
    auto t_host = Clock::now();

    if (!topology_ready) {
        std::vector<double> h_radius, h_pc_radius;
        std::vector<uint8_t> h_nq, h_pc_n;
        const double* radius = io.radius;
        const uint8_t* nq = io.nq;
        const double* pc_radius = has_ext ? io.ext->radius : nullptr;
        const uint8_t* pc_n = has_ext ? io.ext->n : nullptr;
        if (io.device_ptrs) {
            h_radius.resize(n);
            h_nq.resize(n);
            CU(cudaMemcpy(h_radius.data(), io.radius, n * 8, cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(h_nq.data(), io.nq, n, cudaMemcpyDeviceToHost));
            radius = h_radius.data();
            nq = h_nq.data();
            if (m > 0) {
                h_pc_radius.resize(m);
                h_pc_n.resize(m);
                CU(cudaMemcpy(h_pc_radius.data(), io.ext->radius, m * 8, cudaMemcpyDeviceToHost));
                CU(cudaMemcpy(h_pc_n.data(), io.ext->n, m, cudaMemcpyDeviceToHost));
                pc_radius = h_pc_radius.data();
                pc_n = h_pc_n.data();
            }
        }
        std::vector<int> pc_type;
        std::vector<double> h_xyz;
        const double* xyz0 = io.xyz + frame0 * n * 3;   // first frame's geometry orders the atoms in space
        if (io.device_ptrs) {
            h_xyz.resize(3 * n);
            CU(cudaMemcpy(h_xyz.data(), xyz0, (size_t)n * 24, cudaMemcpyDeviceToHost));
            xyz0 = h_xyz.data();
        }
        int rc = make_layout(ctx, n, radius, nq, hydrogen_scf, pc_radius, pc_n, m, xyz0, P.L, pc_type);
        if (rc) return rc;
        PackedTables pk;
        rc = pack_tables(ctx, P.L.types, o.lambda_scale, o.basis, pk);
        if (rc) return rc;
        ctx->stats.pair_types = pk.pair_types;
        rc = upload_tables(ctx, pk, P.view);
        if (rc) return rc;
        P.blob_bytes = (int)pk.blob.size();
        P.rcut2_max = (o.basis == CHEQ_BASIS_STO) ? pk.rcut2_max : INFINITY;
        CU(ctx->perm.ensure(P.L.NP * sizeof(int)));
        CU(ctx->type.ensure(P.L.NP));
        CU(cudaMemcpyAsync(ctx->perm.p, P.L.perm.data(), P.L.NP * sizeof(int), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->type.p, P.L.type.data(), P.L.NP, cudaMemcpyHostToDevice, st));
        if (m > 0) {
            CU(ctx->pc_type.ensure(m * sizeof(int)));
            CU(cudaMemcpyAsync(ctx->pc_type.p, pc_type.data(), m * sizeof(int), cudaMemcpyHostToDevice, st));
        }
        // the async copies above read pageable host vectors that die with this scope
        CU(cudaStreamSynchronize(st));
    }
    const Layout& L = P.L;
    P.batch = batch;
    P.strideA = (long)L.ld * L.NP;
    if (!io.device_ptrs && P.max_hardness == 0.0)
        for (long i = 0; i < n; ++i) P.max_hardness = std::max(P.max_hardness, std::fabs(io.hard[i]));
    ctx->stats.n_atoms = n;
    ctx->stats.n_hydrogen = L.scf ? L.nh : 0;
    ctx->stats.n_padded = L.NP;
    ctx->stats.ms_host_prepare += ms_since(t_host);

    // memory budget
    {
        size_t need = (size_t)P.strideA * 8 * batch;
        if (L.scf) need += (size_t)(L.MT - L.nxp) * L.nhp * 8 * batch;
        need += (size_t)L.NP * kTile * 8 * batch;
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        const size_t have = free_b + ctx->A.cap + ctx->S0.cap + ctx->Linv.cap;
        if (need > have)
            return fail(ctx, CHEQ_ERR_OUT_OF_MEMORY, "workspace of " + std::to_string(need >> 20) +
                                                         " MiB exceeds free device memory");
    }
    CU(ctx->A.ensure((size_t)P.strideA * 8 * batch));
    if (L.scf) CU(ctx->S0.ensure((size_t)(L.MT - L.nxp) * L.nhp * 8 * batch));
    CU(ctx->Linv.ensure((size_t)L.NP * kTile * 8 * batch));
    const size_t vec = (size_t)L.NP * 8 * batch;
    CU(ctx->x.ensure(vec));
    CU(ctx->y.ensure(vec));
    CU(ctx->z.ensure(vec));
    CU(ctx->v.ensure(vec));
    CU(ctx->xsol.ensure(vec));
    CU(ctx->q.ensure(vec));
    CU(ctx->diag.ensure(L.NP * 8));
    CU(ctx->chi.ensure(L.NP * 8));
    CU(ctx->hard.ensure(L.NP * 8));
    CU(ctx->scratch.ensure((size_t)backsolve_max_ctas() * kTile * 8 * batch));
    CU(ctx->counters.ensure(sizeof(unsigned) * batch));
    CU(ctx->sgn.ensure(vec));
    CU(ctx->sgn_flags.ensure((size_t)(L.NP / kTile) * (kTile + 1) * sizeof(int) * batch));   // counts, then lists
    CU(ctx->state.ensure(sizeof(ScfState) * batch));
    CU(ctx->active.ensure(sizeof(int)));
    int rc = ensure_host_state(ctx, batch);
    if (rc) return rc;
    if (!ctx->h_active) CU(cudaMallocHost((void**)&ctx->h_active, sizeof(int)));
    CU(cudaMemsetAsync(ctx->counters.p, 0, sizeof(unsigned) * batch, st));
    CU(cudaMemsetAsync(ctx->q.p, 0, vec, st));

    // inputs -> device
    CU(cudaEventRecord(ctx->ev[0], st));
    const double *d_xyz, *d_chi, *d_hard;
    if (io.device_ptrs) {
        d_xyz = io.xyz + frame0 * n * 3;
        d_chi = io.chi;
        d_hard = io.hard;
    } else {
        CU(ctx->in_xyz.ensure((size_t)n * 24 * batch));
        CU(ctx->in_chi.ensure(n * 8));
        CU(ctx->in_hard.ensure(n * 8));
        CU(cudaMemcpyAsync(ctx->in_xyz.p, io.xyz + frame0 * n * 3, (size_t)n * 24 * batch, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->in_chi.p, io.chi, n * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->in_hard.p, io.hard, n * 8, cudaMemcpyHostToDevice, st));
        d_xyz = ctx->in_xyz.as<double>();
        d_chi = ctx->in_chi.as<double>();
        d_hard = ctx->in_hard.as<double>();
    }
    const double *d_pc_xyz = nullptr, *d_pc_q = nullptr;
    if (m > 0) {
        if (io.device_ptrs) {
            d_pc_xyz = io.ext->xyz;
            d_pc_q = io.ext->charge;
        } else {
            CU(ctx->pc_xyz.ensure(m * 24));
            CU(ctx->pc_q.ensure(m * 8));
            CU(cudaMemcpyAsync(ctx->pc_xyz.p, io.ext->xyz, m * 24, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(ctx->pc_q.p, io.ext->charge, m * 8, cudaMemcpyHostToDevice, st));
            d_pc_xyz = ctx->pc_xyz.as<double>();
            d_pc_q = ctx->pc_q.as<double>();
        }
    }
    CU(cudaEventRecord(ctx->ev[1], st));

    GatherArgs ga{};
    ga.NP = L.NP;
    ga.n = n;
    ga.pos_stride = L.NP;
    ga.perm = ctx->perm.as<int>();
    ga.xyz_in = d_xyz;
    ga.chi_in = d_chi;
    ga.hard_in = d_hard;
    ga.x = ctx->x.as<double>();
    ga.y = ctx->y.as<double>();
    ga.z = ctx->z.as<double>();
    ga.diag = ctx->diag.as<double>();
    ga.chi = ctx->chi.as<double>();
    ga.hard = ctx->hard.as<double>();
    LAUNCH(launch_gather(ga, batch, st));

    const double* d_vext = nullptr;
    if (has_ext) {
        CU(ctx->vext.ensure(L.NP * 8));
        VextArgs va{};
        va.x = ctx->x.as<double>();
        va.y = ctx->y.as<double>();
        va.z = ctx->z.as<double>();
        va.type = ctx->type.as<uint8_t>();
        va.NP = L.NP;
        va.m = m;
        va.pc_xyz = d_pc_xyz;
        va.pc_charge = d_pc_q;
        va.pc_type = ctx->pc_type.as<int>();
        va.field[0] = io.ext->field[0];
        va.field[1] = io.ext->field[1];
        va.field[2] = io.ext->field[2];
        va.v_out = ctx->vext.as<double>();
        LAUNCH(launch_vext(va, P.view, ctx->blob.as<char>(), P.blob_bytes, st));
        d_vext = ctx->vext.as<double>();
    }

    AssembleArgs aa{};
    aa.x = ctx->x.as<double>();
    aa.y = ctx->y.as<double>();
    aa.z = ctx->z.as<double>();
    aa.type = ctx->type.as<uint8_t>();
    aa.diag = ctx->diag.as<double>();
    aa.A = ctx->A.as<double>();
    aa.ld = L.ld;
    aa.a_stride = P.strideA;
    aa.pos_stride = L.NP;
    aa.NP = L.NP;
    aa.rcut2_max = P.rcut2_max;
    if (use_blocked(ctx, batch, L.NP)) {
        // distributed assembly: every rank produces only the tile columns it owns
        rc = ensure_panel_plan(ctx, L);
        if (rc) return rc;
        if (ctx->world > 1 && !all_columns) aa.col_mask = ctx->col_mask.as<uint8_t>();
    }
    LAUNCH(launch_assemble(aa, P.view, ctx->blob.as<char>(), P.blob_bytes, batch, st));
    LAUNCH(launch_rhs_rows(ctx->A.as<double>(), L.ld, P.strideA, L.NP, ctx->chi.as<double>(), d_vext, 0,
                           ctx->type.as<uint8_t>(), batch, st));
    CU(cudaEventRecord(ctx->ev[2], st));
    ctx->stats.bytes_assemble += 8.0 * (double)n * (double)(n + 1) / 2.0 * batch;
    return CHEQ_OK;
}

// Factorisation + SCF for the frames currently assembled in ctx->A.  Fills h_state[0..batch).
int factor_and_scf(cheq_ctx* ctx, const SolveIO& io, Prepared& P) {
    const Layout& L = P.L;
    const cheq_options& o = *io.opt;
    cudaStream_t st = ctx->stream;
    const int batch = P.batch;
    Dense D{ctx, ctx->A.as<double>(), L.ld, P.strideA, ctx->Linv.as<double>(), (long)L.NP * kTile, L.MT,
            ctx->state.as<ScfState>(), batch, ctx->sgn.as<double>(), (long)L.NP, ctx->sgn_flags.as<int>(),
            ctx->sgn_flags.as<int>() + (size_t)(L.NP / kTile) * batch, (long)L.NP / kTile};
    if (P.max_hardness > 0.0) D.pivot_tol = kPivotRel * P.max_hardness;
    double damping0 = 1.0;   // implementation.rs:297-301
    if (L.scf && o.damping_kind != CHEQ_DAMPING_NONE) damping0 = o.damping_value;
    LAUNCH(launch_scf_init(ctx->state.as<ScfState>(), damping0, batch, st));
    double* v = ctx->v.as<double>();
    double* xs = ctx->xsol.as<double>();
    double* q = ctx->q.as<double>();
    int rc;
    const bool blocked = use_blocked(ctx, batch, L.NP);
    const bool dist = blocked && ctx->world > 1;
    const PanelPlan& pp = ctx->pplan;
    if (dist) {
        // what the peers must see besides the trapezoid lives in one arena with the same layout on every rank
        const ArenaLayout AL(L.NP);
        const void* before = ctx->arena.p;
        CU(ctx->arena.ensure(AL.bytes));
        if (ctx->arena.p != before) CU(cudaMemsetAsync(ctx->arena.p, 0, 4096, st));   // fresh signal words
        char* ar = ctx->arena.as<char>();
        D.Linv = (double*)(ar + AL.linv);
        D.sgn = (double*)(ar + AL.sgn);
        D.neg_cnt = (int*)(ar + AL.neg_cnt);
        D.neg_idx = (int*)(ar + AL.neg_idx);
        rc = p2p_setup(ctx);   // start-of-solve barrier + peer mappings
        if (rc) return rc;
        // Data plane for this solve (CHEQ_DIST_P2P: 0 broadcast, 1 automatic, 2 push from the kernels' epilogues,
        // 3 push with the copy engines).  Measured on B200 (profiles/r1_bench_dist*_v[23]_*.json, r1_s100k_*):
        //   S20k, 2 GPUs: DMA push 0.510 s < epilogue push 0.537 < NCCL broadcast 0.560
        //   S20k, 8 GPUs: NCCL broadcast 0.432 s < DMA push 0.491 < epilogue push 0.593  (7 unicast copies of every
        //                 panel cost more than one broadcast when the panel is on the critical path)
        //   S100k, 8 GPUs: DMA push 6.92 s (24.4 TF/s per GPU) < epilogue push 7.52 < NCCL broadcast 8.38
        //                 (the panel is hidden behind long updates; pushing removes pack / broadcast / unpack)
        //   S20k, 4 GPUs: DMA push 0.455 s < NCCL broadcast 0.466
        const bool want_push = ctx->p2p_mode >= 2 || (ctx->p2p_mode == 1 && (ctx->world <= 4 || (long)L.NP >= 49152));
        ctx->push_now = ctx->links.enabled && want_push;
        ctx->push_dma = ctx->push_now && ctx->p2p_mode != 2;
        if (!ctx->push_now) CU(ctx->stage.ensure((size_t)L.MT * ctx->tpp_now * kTile * 8));
    }
    const uint8_t* mask_h = dist ? ctx->col_mask.as<uint8_t>() + L.nxp / kTile : nullptr;

    if (!L.scf) {
        // implementation.rs:284-292: one solve, iterations = 1, charges = raw solution
        if (blocked) {
            rc = factor_blocked(D, 0, (int)pp.begin.size());
        } else {
            D.plan = get_plan(ctx, L.MT, 0, L.NP, batch);
            rc = factor_panel(D, 0, L.NP);
        }
        if (rc) return rc;
        CU(cudaEventRecord(ctx->ev[3], st));
        LAUNCH(launch_mu(D.A, D.ld, D.strideA, L.NP, io.total_charge, v, L.NP, D.sgn, D.strideSgn, D.state, batch, st));
        rc = backsolve(D, L.NP, v, xs, L.NP);
        if (rc) return rc;
        LAUNCH(launch_scf_mix(xs, q, L.NP, ctx->type.as<uint8_t>(), L.NP, D.state, batch, st));
        CU(cudaMemsetAsync(ctx->active.p, 0, sizeof(int), st));
        LAUNCH(launch_scf_decide(D.state, INFINITY, CHEQ_DAMPING_NONE, batch, ctx->active.as<int>(), st));
        CU(cudaMemcpyAsync(ctx->h_state, ctx->state.p, sizeof(ScfState) * batch, cudaMemcpyDeviceToHost, st));
        CU(cudaEventRecord(ctx->ev[4], st));
        CU(cudaStreamSynchronize(st));
        ctx->stats.flops_factor += D.flops * batch;
        return CHEQ_OK;
    }

    // geometry-invariant part: factor the non-hydrogen block, form the hydrogen Schur complement S0 and the
    // reduced right-hand sides; snapshot them.
    if (blocked) {
        // right-looking over the X panels: the trailing updates reach into the owned hydrogen columns, which
        // therefore hold the Schur complement when the last X panel is done
        rc = factor_blocked(D, 0, pp.x_panels);
        if (rc) return rc;
    } else {
        D.plan = get_plan(ctx, L.MT, 0, L.nxp, batch);
        rc = factor_panel(D, 0, L.nxp);
        if (rc) return rc;
        D.plan = get_plan(ctx, L.MT, L.nxp, L.NP, batch);
        rc = schur_update(D, L.nxp, L.nhp);
        if (rc) return rc;
    }
    const long ld0 = L.MT - L.nxp;
    const long stride0 = ld0 * L.nhp;
    double* Ahh = D.A + (long)L.nxp + (long)L.nxp * L.ld;
    const int mt_h = (L.MT - L.nxp) / kTile, nt_h = L.nhp / kTile;
    LAUNCH(launch_copy_lower_tiles(Ahh, L.ld, D.strideA, ctx->S0.as<double>(), ld0, stride0, mt_h, nt_h, nullptr,
                                   batch, st, mask_h));
    CU(cudaEventRecord(ctx->ev[3], st));
    const double flops_once = D.flops;
    D.flops = 0.0;

    const uint8_t* type_h = ctx->type.as<uint8_t>() + L.nxp;
    const double* hard_h = ctx->hard.as<double>() + L.nxp;

    // Stale-factor iterations (kernels_krylov.cu): one frame, one GPU, a hydrogen block that is worth it, and room
    // for R (nhp^2 doubles).  Otherwise every iteration refactors the hydrogen block.
    Krylov K;
    bool kry = batch == 1 && ctx->krylov_mode && L.nhp >= ctx->krylov_min_nhp;
    if (kry) {
        // room for R (all of it on one GPU, the largest row chunk -- about half -- in a group); every rank of a group
        // must take the same decision, and they do: same sizes, same device type, same allocations so far
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        // fp64 R (the largest row chunk of a group: about half) + fp32 tile-major upper tiles of it + tile-major copy of
        // this rank's part of the lower triangle of S0
        const size_t nh2 = (size_t)L.nhp * L.nhp;
        const size_t need = nh2 * 8 / (dist ? 2 : 1) + (ctx->krylov_f32 ? nh2 * 2 / (dist ? 2 : 1) + nh2 * 4 / (dist ? ctx->world : 1) : 0) +
                            ((size_t)512 << 20);
        const size_t have_r = ctx->kr_R.cap + ctx->kr_Rf.cap + ctx->kr_S0t.cap;
        if (need > free_b + have_r) kry = false;
        if (dist) {
            int* flag = ctx->ipc_stage.as<int>() + 32;
            const int mine = kry ? 1 : 0;
            CU(cudaMemcpyAsync(flag, &mine, sizeof(int), cudaMemcpyHostToDevice, st));
            NC(g_nccl.AllReduce(flag, flag, 1, kNcclInt, 3 /* ncclMin */, ctx->comm, st));
            int all = 0;
            CU(cudaMemcpyAsync(&all, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            kry = all != 0;
        }
    }
    if (kry) {
        rc = krylov_alloc(ctx, K, L.nhp, dist, dist ? ctx->pplan.mask.data() + L.nxp / kTile : nullptr);
        if (rc) return rc;
        K.S0 = ctx->S0.as<double>();
        K.ld0 = ld0;
        K.sgn_h = D.sgn + L.nxp;
        K.col_mask = mask_h;
        if (dist) {
            const ArenaLayout AL(L.NP);
            K.kx_data = AL.kx_data;
            K.kx_slot = AL.kx_slot;
            const bool want_p2p = ctx->krylov_p2p > 0 || (ctx->krylov_p2p < 0 && ctx->world >= 4);
            K.p2p = want_p2p && ctx->links.enabled && ctx->arena.cap >= AL.bytes;
            CU(ctx->kx_local.ensure(64));
            CU(cudaMemsetAsync(ctx->kx_local.p, 0, 64, st));
        }
        LAUNCH(launch_krylov_setup(D.A, D.ld, L.NP, L.nxp, D.sgn, K.S0, ld0, L.nhp, io.total_charge, K.sc, K.tc, K.t1,
                                   mask_h, (!dist || ctx->rank == 0) ? 1 : 0, st));
        if (dist) {
            // the reduced right-hand sides live in the owners' columns: complete them on every rank
            rc = krylov_allreduce(ctx, K, K.tc);
            if (rc) return rc;
            rc = krylov_allreduce(ctx, K, K.t1);
            if (rc) return rc;
        }
    }
    // When to stop refactorising.  A refactorisation costs T_f; keeping a factor whose shift is d eV away costs about
    // 1.4 d extra GMRES steps in each of the ~14 iterations that follow (measured: 36 / 20 / 13 / 9 / 7 steps at
    // d = 8 / 6 / 3.4 / 0.8 / 0.2 eV on S20k), and a refactorisation roughly halves d.  So refactor while
    // d > T_f / (14 * 1.4 * T_step).  T_f and T_step come from a size model (identical on every rank of a group, which
    // must take the same decision), calibrated on S20k (1 and 8 GPUs) and S100k (8 GPUs): S20k -> 3.5..4 eV (freeze at
    // iteration 4, as tuned by hand), S100k on 8 GPUs -> 12 eV (freeze at iteration 2: a refactorisation there costs
    // 525 ms, a GMRES step 1.6 ms).
    double freeze_ev = ctx->krylov_freeze_ev;
    if (kry && !(freeze_ev > 0.0)) {
        const double w = dist ? ctx->world : 1;
        const double nh = (double)L.nhp, tiles = nh / kTile;
        const double t_f = nh * nh * nh / 3.0 / ((dist ? 22e12 : 25e12) * w) * 1e3 + tiles * (dist ? 0.17 : 0.075);
        const double t_step = 8.0 * nh * nh / (4e12 * w) * 1e3 + (dist ? 0.4 : 0.15);
        freeze_ev = std::min(12.0, std::max(3.5, t_f / (14.0 * 1.4 * t_step)));
    }
    cudaEvent_t evk0 = ctx->ev[5], evk1 = ctx->ev[6];
    float krylov_ms = 0.0f;
    bool have_factor = false;        // the trapezoid's hydrogen block holds the factor for the shift in K.d_prev
    uint32_t it = 0;
    double flops_iter_total = 0.0;
    for (it = 1; it <= o.max_iterations; ++it) {
        bool use_krylov = false;
        int gm_steps = 0;
        if (kry) {
            // this iteration's diagonal shift and how far it is from the factor we hold
            CU(cudaMemsetAsync(&K.sc->dmax_bits, 0, sizeof(unsigned long long), st));
            LAUNCH(launch_h_shift(hard_h, q + L.nxp, type_h, L.nhp, K.d_cur, K.frozen ? K.d_frozen : K.d_prev, K.sc, st));
            if (K.frozen) {
                use_krylov = true;
            } else if (have_factor && it >= 2) {
                CU(cudaMemcpyAsync(&ctx->h_ksc->dmax_bits, &K.sc->dmax_bits, sizeof(unsigned long long),
                                   cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                double dmax;
                std::memcpy(&dmax, &ctx->h_ksc->dmax_bits, 8);
                if (dmax <= freeze_ev || (int)it >= ctx->krylov_force_it) {
                    // the factor of the previous iteration becomes the preconditioner of all that follow
                    CU(cudaEventRecord(evk0, st));
                    D.flops = 0.0;
                    rc = krylov_freeze(ctx, D, K, L.nxp, K.d_prev);
                    if (rc) return rc;
                    flops_iter_total += D.flops;
                    CU(cudaEventRecord(evk1, st));
                    use_krylov = true;
                    CU(cudaStreamSynchronize(st));
                    float ms = 0;
                    cudaEventElapsedTime(&ms, evk0, evk1);
                    krylov_ms += ms;
                }
            }
        }
        if (use_krylov) {
            CU(cudaEventRecord(evk0, st));
            bool ok = false;
            rc = krylov_solve(ctx, K, 1e-13, &gm_steps, &ok);
            if (rc) return rc;
            if (K.p2p) {
                int kx_err = 0;
                CU(cudaMemcpyAsync(&kx_err, ctx->kx_local.as<int>() + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                if (kx_err) return fail(ctx, CHEQ_ERR_CUDA, "peer-memory all-reduce timed out waiting for a rank");
            }
            if (ok) {
                // x_x = L_xx^-T (S (yc_x - mu y1_x) - W^T x_h): the hydrogen part W^T x_h is one HBM-speed pass over
                // W = L[hydrogen rows, non-hydrogen columns] (split by row chunks over the ranks of a group), then the
                // back-substitution chain runs over the non-hydrogen blocks only
                if (int rcp = prof_begin(ctx, 2)) return rcp;
                const int nth = L.nhp / kTile, ntx = L.nxp / kTile;
                double* wx = nullptr;
                if (ntx > 0) {
                    CU(ctx->kr_wx.ensure((size_t)(nth + 1) * ntx * kTile * 8 + (size_t)(L.nxp + 8) * 8));
                    double* wpart = ctx->kr_wx.as<double>();
                    wx = wpart + (size_t)nth * ntx * kTile;
                    int wb = 0, we = nth;
                    if (dist) {
                        wb = (int)((long)nth * ctx->rank / ctx->world);
                        we = (int)((long)nth * (ctx->rank + 1) / ctx->world);
                    }
                    LAUNCH(launch_rect_matvec_t(D.A + L.nxp, D.ld, nth, ntx, wb, we, K.sol, wpart, wx, st));
                    if (dist)
                        if (int rca = krylov_allreduce_n(ctx, K, wx, L.nxp)) return rca;
                    K.bytes += 8.0 * (double)L.nhp * L.nxp / (dist ? ctx->world : 1);
                }
                LAUNCH(launch_krylov_finish(D.A, D.ld, L.NP, L.nxp, K.sol, L.nhp, v, xs, D.state, wx, st));
                CU(ctx->chain.ensure((size_t)(L.NP / kTile + 1) * sizeof(unsigned)));
                LAUNCH(launch_backsolve_chain(D.A, D.ld, D.strideA, D.Linv, D.strideLinv, D.sgn, D.strideSgn, L.nxp, v,
                                              xs, L.NP, ctx->chain.as<unsigned>(), D.state, 1, st));
                if (int rcp = prof_end(ctx)) return rcp;
            } else {
                // GMRES did not reach the target with this factor: refactor at the current shift instead
                use_krylov = false;
                K.frozen = false;
            }
            CU(cudaEventRecord(evk1, st));
            CU(cudaStreamSynchronize(st));
            float ms = 0;
            cudaEventElapsedTime(&ms, evk0, evk1);
            krylov_ms += ms;
        }
        if (!use_krylov) {
            if (it > 1)
                LAUNCH(launch_copy_lower_tiles(ctx->S0.as<double>(), ld0, stride0, Ahh, L.ld, D.strideA, mt_h, nt_h,
                                               D.state, batch, st, mask_h));
            LAUNCH(launch_set_h_diagonal(Ahh, L.ld, D.strideA, ctx->S0.as<double>(), ld0, stride0, hard_h,
                                         q + L.nxp, L.NP, type_h, L.nhp, D.state, batch, st, mask_h));
            D.flops = 0.0;
            rc = blocked ? factor_blocked(D, pp.x_panels, (int)pp.begin.size()) : factor_panel(D, L.nxp, L.nhp);
            if (rc) return rc;
            LAUNCH(launch_mu(D.A, D.ld, D.strideA, L.NP, io.total_charge, v, L.NP, D.sgn, D.strideSgn, D.state, batch, st));
            rc = backsolve(D, L.NP, v, xs, L.NP);
            if (rc) return rc;
            flops_iter_total += D.flops * (double)batch;   // upper bound for batches (finished frames skip)
            if (kry) {
                // remember which shift this factor belongs to, and its solution as the next initial guess
                std::swap(K.d_cur, K.d_prev);
                LAUNCH(launch_krylov_capture(xs, L.nxp, L.nhp, D.state, K.sol, st));
                have_factor = true;
            }
        }
        LAUNCH(launch_scf_mix(xs, q, L.NP, ctx->type.as<uint8_t>(), L.NP, D.state, batch, st));
        CU(cudaMemsetAsync(ctx->active.p, 0, sizeof(int), st));
        LAUNCH(launch_scf_decide(D.state, o.tolerance, o.damping_kind, batch, ctx->active.as<int>(), st));
        if (dist) {
            // every rank solved the same replicated triangular systems; rank 0's iterate and decision are made
            // authoritative so that the control flow of all ranks stays in lock-step
            cudaStream_t ps = ctx->panel_stream;
            CU(cudaEventRecord(ctx->ev_ready, st));
            CU(cudaStreamWaitEvent(ps, ctx->ev_ready, 0));
            NC(g_nccl.GroupStart());
            NC(g_nccl.Broadcast(D.state, D.state, sizeof(ScfState), kNcclChar, 0, ctx->comm, ps));
            NC(g_nccl.Broadcast(ctx->active.p, ctx->active.p, sizeof(int), kNcclChar, 0, ctx->comm, ps));
            NC(g_nccl.Broadcast(q, q, (size_t)L.NP, kNcclDouble, 0, ctx->comm, ps));
            // barrier: every rank has finished reading this iteration's factor before any owner pushes the next one
            NC(g_nccl.AllReduce(ctx->ipc_stage.as<int>() + 16, ctx->ipc_stage.as<int>() + 16, 1, kNcclInt, kNcclMax,
                                ctx->comm, ps));
            NC(g_nccl.GroupEnd());
            CU(cudaMemcpyAsync(ctx->h_active, ctx->active.p, sizeof(int), cudaMemcpyDeviceToHost, ps));
            if (batch == 1) CU(cudaMemcpyAsync(ctx->h_state, ctx->state.p, sizeof(ScfState), cudaMemcpyDeviceToHost, ps));
            CU(cudaStreamSynchronize(ps));
        } else {
            CU(cudaMemcpyAsync(ctx->h_active, ctx->active.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            if (batch == 1) CU(cudaMemcpyAsync(ctx->h_state, ctx->state.p, sizeof(ScfState), cudaMemcpyDeviceToHost, st));
        }
        if (ctx->h_oz_err) CU(cudaMemcpyAsync(ctx->h_oz_err, ctx->oz_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (ctx->h_oz_err && *ctx->h_oz_err) {
            const int code = *ctx->h_oz_err;
            cudaMemsetAsync(ctx->oz_err.p, 0, sizeof(int), st);
            return fail(ctx, CHEQ_ERR_CUDA, "tcgen05 slice-product pipeline timed out (role " + std::to_string(code) + ")");
        }
        if (batch == 1) ctx->trace.push_back({ctx->h_state[0].delta, use_krylov ? 1 : 0, gm_steps});
        if (*ctx->h_active == 0) break;
    }
    ctx->stats.krylov_ms += krylov_ms;
    ctx->stats.krylov_bytes += K.bytes;
    CU(cudaMemcpyAsync(ctx->h_state, ctx->state.p, sizeof(ScfState) * batch, cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(ctx->ev[4], st));
    CU(cudaStreamSynchronize(st));
    ctx->stats.flops_factor += flops_once * batch + flops_iter_total;
    return CHEQ_OK;
}

int frame_status(const ScfState& s) {
    if (s.info != 0) return CHEQ_ERR_LINALG;
    if (!std::isfinite(s.mu)) return CHEQ_ERR_LINALG;
    if (!s.converged && s.active) return CHEQ_ERR_NOT_CONVERGED;   // still active when the loop ran out
    return CHEQ_OK;
}

// Pivoted-LU path for ONE frame whose matrix is not positive definite: the reference's algorithm as written
// (fresh partial-pivot LU of the bordered matrix every SCF iteration, implementation.rs:303-314, 574-576).
// Expects a fresh assembly of that frame in ctx->A.
int lu_scf(cheq_ctx* ctx, const SolveIO& io, Prepared& P) {
    const Layout& L = P.L;
    const cheq_options& o = *io.opt;
    cudaStream_t st = ctx->stream;
    const int n1p = lu_padded_size(L.NP);
    CU(ctx->lu_f.ensure(lu_matrix_bytes(L.NP)));
    CU(ctx->lu_ut.ensure(lu_workspace_bytes(L.NP)));
    CU(ctx->lu_piv.ensure((size_t)n1p * sizeof(int)));
    ScfState* state = ctx->state.as<ScfState>();
    double damping0 = 1.0;
    if (L.scf && o.damping_kind != CHEQ_DAMPING_NONE) damping0 = o.damping_value;
    LAUNCH(launch_scf_init(state, damping0, 1, st));
    CU(cudaMemsetAsync(ctx->q.p, 0, (size_t)L.NP * 8, st));
    double* v = ctx->v.as<double>();
    double* q = ctx->q.as<double>();
    unsigned long long launches = 0;
    const uint32_t max_it = L.scf ? o.max_iterations : 1;
    for (uint32_t it = 1; it <= max_it; ++it) {
        CU(launch_lu_solve(ctx->A.as<double>(), L.ld, L.NP, L.nxp, L.scf ? 1 : 0, ctx->lu_f.as<double>(),
                           ctx->lu_ut.as<double>(), q,
                           ctx->hard.as<double>(), ctx->type.as<uint8_t>(), io.total_charge,
                           ctx->lu_piv.as<int>(), v, state, &launches, st));
        LAUNCH(launch_scf_mix(v, q, L.NP, ctx->type.as<uint8_t>(), L.NP, state, 1, st));
        CU(cudaMemsetAsync(ctx->active.p, 0, sizeof(int), st));
        LAUNCH(launch_scf_decide(state, L.scf ? o.tolerance : INFINITY, L.scf ? o.damping_kind : CHEQ_DAMPING_NONE,
                                 1, ctx->active.as<int>(), st));
        CU(cudaMemcpyAsync(ctx->h_active, ctx->active.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (*ctx->h_active == 0) break;
    }
    ctx->launches += launches;
    CU(cudaMemcpyAsync(ctx->h_state, ctx->state.p, sizeof(ScfState), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return CHEQ_OK;
}

// Scatter + download the charges of `batch` frames starting at global frame gf0, and fill the scalar outputs.
// `redo` (nullable): frames of the chunk that the pivoted-LU pass re-solves afterwards; their provisional status
// does not count towards `worst`.
int emit_results(cheq_ctx* ctx, const SolveIO& io, const Prepared& P, long gf0, int batch, int* worst,
                 const std::vector<int>* redo = nullptr) {
    cudaStream_t st = ctx->stream;
    const long n = io.n;
    double* d_out;
    if (io.device_ptrs) {
        d_out = io.charges_out + gf0 * n;
    } else {
        CU(ctx->out_q.ensure((size_t)n * 8 * batch));
        d_out = ctx->out_q.as<double>();
    }
    LAUNCH(launch_scatter_charges(P.L.NP, ctx->perm.as<int>(), ctx->q.as<double>(), P.L.NP, d_out, n, batch, st));
    CU(cudaEventRecord(ctx->ev[5], st));
    if (!io.device_ptrs)
        CU(cudaMemcpyAsync(io.charges_out + gf0 * n, d_out, (size_t)n * 8 * batch, cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(ctx->ev[6], st));
    CU(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]);
    ctx->stats.ms_d2h += ms;
    for (int f = 0; f < batch; ++f) {
        const ScfState& s = ctx->h_state[f];
        const int fs = frame_status(s);
        const long gf = gf0 + f;
        if (io.potential_out) io.potential_out[gf] = s.mu;
        if (io.iterations_out) io.iterations_out[gf] = P.L.scf ? (uint32_t)s.iteration : 1u;
        if (io.delta_out) io.delta_out[gf] = s.delta;
        if (io.status_out) io.status_out[gf] = fs;
        const bool redone = redo && std::find(redo->begin(), redo->end(), f) != redo->end();
        if (fs != CHEQ_OK && *worst == CHEQ_OK && !redone) {
            *worst = fs;
            if (fs == CHEQ_ERR_LINALG) {
                ctx->err = "Linear system solver panicked. Matrix might be singular.";
            } else {
                char buf[160];
                snprintf(buf, sizeof buf, "SCF failed to converge after %u iterations. Final charge difference: %.2e",
                         io.opt->max_iterations, s.delta);
                ctx->err = buf;
            }
        }
        ctx->stats.iterations = std::max<uint32_t>(ctx->stats.iterations, P.L.scf ? (uint32_t)s.iteration : 1u);
    }
    return CHEQ_OK;
}


// ---- matrix-free path --------------------------------------------------------------------------------------------
// QEqSolver::solve for systems whose (N + 1)^2 matrix does not fit (SURVEY.md section 8(e) last row, App. E.4).  The
// reference holds three dense copies of it (implementation.rs:464, 282, 575); here it is never formed:
//   K(q) [x; mu] = [c; Q],  K = [[diag(hardness + D(q)) + J, e], [e^T, 0]]
// is solved in every SCF iteration by restarted GMRES (the matrix is symmetric but indefinite, DESIGN.md section 5.2,
// so CG does not apply), K v recomputes the J tiles on the fly from the same device function as the assembly, rows
// sharded over the ranks of the group with one ncclAllGather of the result per application (the Krylov vectors are
// replicated, so dot products need no collective and all ranks stay bit-identical).  Preconditioner: block-Jacobi over
// 128-atom blocks of the space-filling-curve order (compact clusters of whole molecules), each block factorised
// A_b = L S L^T by the diagonal-tile kernel of the dense path, applied as M^T S M with M = S L^-1, border eliminated
// exactly as in the stale-factor iterations.  Measured on the CPU model (tools/exp_matrix_free.py): 103 / 107 / 159
// GMRES steps to 1e-12 at 3k / 8k / 20k atoms.

int solve_matfree(cheq_ctx* ctx, const SolveIO& io) {
    const long n = io.n;
    const cheq_options& o = *io.opt;
    cudaStream_t st = ctx->stream;
    const int world = ctx->world, rank = ctx->rank;
    const bool has_ext = !external_is_empty(io.ext);
    const long m = has_ext ? (long)io.ext->num_point_charges : 0;
    auto t_host = Clock::now();
    if (io.device_ptrs) return fail(ctx, CHEQ_ERR_INVALID_ARGUMENT, "the matrix-free path takes host pointers");

    // layout (pure Morton order, padded so that every rank owns a whole number of 128-row tiles and the column tiles
    // of 256 divide NP), tables
    Layout L;
    std::vector<int> pc_type;
    const int pad = 256 * world;
    int rc = make_layout(ctx, n, io.radius, io.nq, o.hydrogen_scf != 0, has_ext ? io.ext->radius : nullptr,
                         has_ext ? io.ext->n : nullptr, m, io.xyz, L, pc_type, pad);
    if (rc) return rc;
    const int NP = L.NP;
    PackedTables pk;
    rc = pack_tables(ctx, L.types, o.lambda_scale, o.basis, pk);
    if (rc) return rc;
    PairTableView view;
    rc = upload_tables(ctx, pk, view);
    if (rc) return rc;
    const int blob_bytes = (int)pk.blob.size();
    const double rcut2_max = (o.basis == CHEQ_BASIS_STO) ? pk.rcut2_max : INFINITY;
    std::vector<uint8_t> hflag(NP, 0);
    for (int i = 0; i < NP; ++i)
        if (L.perm[i] >= 0 && L.scf && io.nq[L.perm[i]] == 1) hflag[i] = 1;
    ctx->stats.n_atoms = n;
    ctx->stats.n_hydrogen = 0;
    for (uint8_t f : hflag) ctx->stats.n_hydrogen += f;
    ctx->stats.n_padded = NP;
    ctx->stats.pair_types = pk.pair_types;

    // device workspace
    if (!ctx->mf) ctx->mf = new MatfreeBuffers();
    MatfreeBuffers& MB = *ctx->mf;
    const int restart = std::min(kGmresMaxBasis, std::max(32, ctx->matfree_restart));
    const long vs = NP + 8;
    const int nb = NP / kTile;
    const int rows = NP / world, row0 = rank * rows;
    // column slices per row tile: enough CTAs (~8 waves of 4 resident CTAs per SM) that the last, partial wave does
    // not matter, but at least 4 column tiles of 256 per slice
    const int row_ctas = rows / kTile, col_tiles = NP / 256;
    int slices = std::max(1, std::min((8 * 4 * 148 + row_ctas - 1) / row_ctas, col_tiles / 4));
    slices = (col_tiles + (col_tiles + slices - 1) / slices - 1) / ((col_tiles + slices - 1) / slices);   // no empty slice
    CU(ctx->perm.ensure((size_t)NP * sizeof(int)));
    CU(ctx->type.ensure(NP));
    const size_t vec = (size_t)NP * 8;
    CU(ctx->x.ensure(vec));
    CU(ctx->y.ensure(vec));
    CU(ctx->z.ensure(vec));
    CU(ctx->q.ensure(vec));
    CU(ctx->xsol.ensure(vec));
    CU(ctx->diag.ensure(vec));
    CU(ctx->chi.ensure(vec));
    CU(ctx->hard.ensure(vec));
    CU(ctx->state.ensure(sizeof(ScfState)));
    CU(ctx->active.ensure(sizeof(int)));
    CU(ctx->kr_sc.ensure(sizeof(KrylovScalars)));
    if (!ctx->h_ksc) CU(cudaMallocHost((void**)&ctx->h_ksc, sizeof(KrylovScalars)));
    rc = ensure_host_state(ctx, 1);
    if (rc) return rc;
    if (!ctx->h_active) CU(cudaMallocHost((void**)&ctx->h_active, sizeof(int)));
    CU(MB.vec.ensure((size_t)10 * vs * 8));
    CU(MB.basis.ensure((size_t)(restart + 1) * vs * 8));
    CU(MB.part.ensure((size_t)slices * 2 * rows * 8));
    CU(MB.blocks.ensure((size_t)nb * kTile * kTile * 8));
    CU(MB.minv.ensure((size_t)nb * kTile * kTile * 8));
    CU(MB.sgn.ensure(vec));
    CU(MB.hflag.ensure(NP));
    double* v0 = MB.vec.as<double>();
    double *rhs = v0, *evec = v0 + vs, *wP = v0 + 2 * vs, *shift = v0 + 3 * vs, *sol = v0 + 4 * vs, *gr = v0 + 5 * vs,
           *gz = v0 + 6 * vs, *gw = v0 + 7 * vs, *tmp = v0 + 8 * vs, *kv = v0 + 9 * vs;
    (void)kv;
    CU(cudaMemsetAsync(MB.vec.p, 0, (size_t)10 * vs * 8, st));
    CU(cudaMemsetAsync(ctx->kr_sc.p, 0, sizeof(KrylovScalars), st));
    CU(cudaMemsetAsync(ctx->q.p, 0, vec, st));
    KrylovScalars* sc = (KrylovScalars*)ctx->kr_sc.p;

    // inputs -> device, internal order
    CU(cudaEventRecord(ctx->ev[0], st));
    CU(ctx->in_xyz.ensure((size_t)n * 24));
    CU(ctx->in_chi.ensure(n * 8));
    CU(ctx->in_hard.ensure(n * 8));
    CU(cudaMemcpyAsync(ctx->in_xyz.p, io.xyz, (size_t)n * 24, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->in_chi.p, io.chi, n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->in_hard.p, io.hard, n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->perm.p, L.perm.data(), (size_t)NP * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->type.p, L.type.data(), NP, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(MB.hflag.p, hflag.data(), NP, cudaMemcpyHostToDevice, st));
    if (m > 0) {
        CU(ctx->pc_type.ensure(m * sizeof(int)));
        CU(ctx->pc_xyz.ensure(m * 24));
        CU(ctx->pc_q.ensure(m * 8));
        CU(cudaMemcpyAsync(ctx->pc_type.p, pc_type.data(), m * sizeof(int), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->pc_xyz.p, io.ext->xyz, m * 24, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->pc_q.p, io.ext->charge, m * 8, cudaMemcpyHostToDevice, st));
    }
    CU(cudaStreamSynchronize(st));   // the copies above read pageable host vectors of this scope
    CU(cudaEventRecord(ctx->ev[1], st));
    ctx->stats.ms_host_prepare += ms_since(t_host);
    GatherArgs ga{};
    ga.NP = NP;
    ga.n = n;
    ga.pos_stride = NP;
    ga.perm = ctx->perm.as<int>();
    ga.xyz_in = ctx->in_xyz.as<double>();
    ga.chi_in = ctx->in_chi.as<double>();
    ga.hard_in = ctx->in_hard.as<double>();
    ga.x = ctx->x.as<double>();
    ga.y = ctx->y.as<double>();
    ga.z = ctx->z.as<double>();
    ga.diag = ctx->diag.as<double>();
    ga.chi = ctx->chi.as<double>();
    ga.hard = ctx->hard.as<double>();
    LAUNCH(launch_gather(ga, 1, st));
    const double* d_vext = nullptr;
    if (has_ext) {
        CU(ctx->vext.ensure(vec));
        VextArgs va{};
        va.x = ga.x;
        va.y = ga.y;
        va.z = ga.z;
        va.type = ctx->type.as<uint8_t>();
        va.NP = NP;
        va.m = m;
        va.pc_xyz = ctx->pc_xyz.as<double>();
        va.pc_charge = ctx->pc_q.as<double>();
        va.pc_type = ctx->pc_type.as<int>();
        va.field[0] = io.ext->field[0];
        va.field[1] = io.ext->field[1];
        va.field[2] = io.ext->field[2];
        va.v_out = ctx->vext.as<double>();
        LAUNCH(launch_vext(va, view, ctx->blob.as<char>(), blob_bytes, st));
        d_vext = ctx->vext.as<double>();
    }
    LAUNCH(launch_matfree_rhs(ga.chi, d_vext, ctx->type.as<uint8_t>(), NP, io.total_charge, rhs, evec, st));

    // block-Jacobi preconditioner: factor the 128 x 128 diagonal blocks of A(q = 0)
    MatfreeArgs ma{};
    ma.x = ga.x;
    ma.y = ga.y;
    ma.z = ga.z;
    ma.type = ctx->type.as<uint8_t>();
    ma.diag = ga.diag;
    ma.shift = shift;
    ma.NP = NP;
    ma.row0 = row0;
    ma.rows = rows;
    ma.slices = slices;
    ma.vstride = vs;
    ma.part = MB.part.as<double>();
    ma.rcut2_max = rcut2_max;
    LAUNCH(launch_assemble_diag_blocks(ma, view, ctx->blob.as<char>(), blob_bytes, MB.blocks.as<double>(), st));
    LAUNCH(launch_potrf_tile(MB.blocks.as<double>(), kTile, (long)kTile * kTile, MB.minv.as<double>(), (long)kTile * kTile,
                             MB.sgn.as<double>(), kTile, nullptr, nullptr, 0, nullptr, kPivotTol, nb, st));
    auto apply_P = [&](const double* in, double* out) -> int {
        LAUNCH(launch_block_jacobi_apply(MB.minv.as<double>(), MB.sgn.as<double>(), in, tmp, NP, 1, vs, st));
        LAUNCH(launch_precond_border(tmp, in, evec, wP, sc, out, NP, st));
        return CHEQ_OK;
    };
    LAUNCH(launch_block_jacobi_apply(MB.minv.as<double>(), MB.sgn.as<double>(), evec, wP, NP, 1, vs, st));
    LAUNCH(launch_precond_den(evec, wP, sc, NP, st));   // sc->g = 0: den = e . P_A^-1 e
    uint64_t matvecs = 0;
    auto apply_K = [&](const double* v, double* out) -> int {
        MatfreeArgs a = ma;
        a.vin = v;
        a.out = out;
        LAUNCH(launch_matfree_matvec(a, 1, view, ctx->blob.as<char>(), blob_bytes, st));
        if (world > 1)
            NC(g_nccl.AllGather(out + row0, out, (size_t)rows, kNcclDouble, ctx->comm, st));
        LAUNCH(launch_matfree_border(a, 1, st));
        ++matvecs;
        return CHEQ_OK;
    };
    CU(cudaEventRecord(ctx->ev[2], st));
    CU(cudaEventRecord(ctx->ev[3], st));

    GmresWork G;
    G.len = NP + 1;
    G.vs = vs;
    G.m = restart;
    G.max_steps = ctx->matfree_max_steps;
    G.r = gr;
    G.z = gz;
    G.w = gw;
    G.V = MB.basis.as<double>();
    G.sc = sc;
    G.stall_accept = 1e2;
    const double tol = ctx->matfree_tol;

    ScfState* state = ctx->state.as<ScfState>();
    double damping0 = 1.0;
    if (L.scf && o.damping_kind != CHEQ_DAMPING_NONE) damping0 = o.damping_value;
    LAUNCH(launch_scf_init(state, damping0, 1, st));
    double* q = ctx->q.as<double>();
    double* xs = ctx->xsol.as<double>();
    const uint32_t max_it = L.scf ? o.max_iterations : 1;
    int status = CHEQ_OK;
    for (uint32_t it = 1; it <= max_it; ++it) {
        if (L.scf) LAUNCH(launch_matfree_shift(ga.hard, q, MB.hflag.as<uint8_t>(), NP, shift, st));
        int steps = 0;
        bool ok = false;
        rc = gmres_solve(ctx, G, apply_K, apply_P, rhs, sol, tol, &steps, &ok);
        if (rc) return rc;
        if (!ok) {
            status = CHEQ_ERR_LINALG;
            ctx->err = "matrix-free GMRES did not converge in " + std::to_string(steps) + " operator applications";
            break;
        }
        LAUNCH(launch_matfree_unpack(sol, NP, xs, state, st));
        LAUNCH(launch_scf_mix(xs, q, NP, ctx->type.as<uint8_t>(), NP, state, 1, st));
        CU(cudaMemsetAsync(ctx->active.p, 0, sizeof(int), st));
        LAUNCH(launch_scf_decide(state, L.scf ? o.tolerance : INFINITY, L.scf ? o.damping_kind : CHEQ_DAMPING_NONE, 1,
                                 ctx->active.as<int>(), st));
        CU(cudaMemcpyAsync(ctx->h_active, ctx->active.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ctx->h_state, ctx->state.p, sizeof(ScfState), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        ctx->trace.push_back({ctx->h_state[0].delta, 2, steps});
        if (*ctx->h_active == 0) break;
    }
    CU(cudaMemcpyAsync(ctx->h_state, ctx->state.p, sizeof(ScfState), cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(ctx->ev[4], st));
    CU(cudaStreamSynchronize(st));
    ctx->stats.matfree_matvecs = matvecs;
    ctx->stats.matfree_pairs = (double)matvecs * (double)n * (double)(n - 1);
    if (status != CHEQ_OK) return status;

    Prepared P;
    P.L = L;
    int worst = CHEQ_OK;
    rc = emit_results(ctx, io, P, 0, 1, &worst);
    if (rc) return rc;
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->stats.ms_h2d += ms;
    cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]);
    ctx->stats.ms_assemble += ms;
    cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]);
    ctx->stats.ms_scf += ms;
    return worst;
}

int solve_impl(cheq_ctx* ctx, const SolveIO& io) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    LocalCallGuard local_guard(ctx, io.collective && io.frames == 1);
    auto t_total = Clock::now();
    ctx->err.clear();
    ctx->stats = cheq_stats{};
    ctx->launches = 0;
    ctx->prof_used = 0;
    ctx->prof_flops = 0.0;
    ctx->trace.clear();
    if (io.n == 0) return fail(ctx, CHEQ_ERR_NO_ATOMS, "Input validation failed: at least one atom is required for a calculation");
    int rc = validate_options(ctx, io.opt);
    if (rc) return rc;
    if (!io.xyz || !io.chi || !io.hard || !io.radius || !io.nq || !io.charges_out)
        return fail(ctx, CHEQ_ERR_INVALID_ARGUMENT, "NULL array argument");
    CU(cudaSetDevice(ctx->device));
    {
        // the dense direct path needs the lower trapezoid (+ the hydrogen snapshot): say so before allocating
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        const double np = (double)io.n + 2.0 * kTile;
        const double need = 8.0 * np * (np + kRhsRows);
        const bool too_big = need > 0.97 * (double)total_b;
        if (io.frames == 1 && !io.device_ptrs && (io.path == 2 || (io.path == 0 && too_big))) {
            // matrix-free path: chosen explicitly (cheq_solve_iterative) or because the matrix does not fit
            const int rcm = solve_matfree(ctx, io);
            ctx->stats.kernel_launches = ctx->launches;
            ctx->stats.ms_total = ms_since(t_total);
            return rcm;
        }
        if (too_big)
            return fail(ctx, CHEQ_ERR_OUT_OF_MEMORY,
                        "dense direct solve of " + std::to_string(io.n) + " atoms needs at least " +
                            std::to_string((size_t)(need / 1073741824.0)) + " GiB for the matrix; the device has " +
                            std::to_string(total_b >> 30) + " GiB (use the matrix-free path: cheq_solve_iterative)");
    }
    const long n = io.n;

    // frames per chunk from the memory budget (single solve: 1)
    long chunk = 1;
    Prepared P;
    if (io.frames > 1) {
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        const long NPmax = round_up_tile((int)n) + kTile;
        const double per_frame = 8.0 * ((double)(NPmax + kRhsRows) * NPmax * 2.0 + (double)NPmax * kTile + 8.0 * NPmax);
        const double budget = 0.8 * (double)(free_b + ctx->A.cap + ctx->S0.cap + ctx->Linv.cap);
        chunk = std::max<long>(1, std::min<long>(io.frames, (long)(budget / per_frame)));
        chunk = std::min<long>(chunk, 4096);
    }
    bool topo_ready = false;
    int worst = CHEQ_OK;
    for (long f0 = 0; f0 < io.frames; f0 += chunk) {
        const int batch = (int)std::min<long>(chunk, io.frames - f0);
        rc = prepare_and_assemble(ctx, io, f0, batch, io.opt->hydrogen_scf != 0, P, topo_ready);
        if (rc) return rc;
        topo_ready = true;
        rc = factor_and_scf(ctx, io, P);
        if (rc) return rc;
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
        ctx->stats.ms_h2d += ms;
        cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]);
        ctx->stats.ms_assemble += ms;
        cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]);
        ctx->stats.ms_factor_once += ms;
        cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]);
        ctx->stats.ms_scf += ms;
        std::vector<int> not_spd;
        for (int f = 0; f < batch; ++f)
            if (ctx->h_state[f].info != 0) not_spd.push_back(f);
        if (ctx->force_lu && batch == 1 && not_spd.empty()) not_spd.push_back(0);
        if (batch == 1 && !not_spd.empty()) {
            // single solve: redo with the pivoted LU
            rc = prepare_and_assemble(ctx, io, f0, 1, io.opt->hydrogen_scf != 0, P, true, true);
            if (rc) return rc;
            rc = lu_scf(ctx, io, P);
            if (rc) return rc;
            ctx->stats.lu_fallback_frames += 1;
            not_spd.clear();
        }
        rc = emit_results(ctx, io, P, f0, batch, &worst, &not_spd);
        if (rc) return rc;
        // batched frames that were not positive definite: one LU solve each, results overwrite the frame's slots
        for (int f : not_spd) {
            rc = prepare_and_assemble(ctx, io, f0 + f, 1, io.opt->hydrogen_scf != 0, P, true, true);
            if (rc) return rc;
            rc = lu_scf(ctx, io, P);
            if (rc) return rc;
            ctx->stats.lu_fallback_frames += 1;
            rc = emit_results(ctx, io, P, f0 + f, 1, &worst);
            if (rc) return rc;
        }
        // (the statuses of the chunk's other frames were folded into `worst` by emit_results: it skips exactly
        //  the frames listed in `redo`)
    }
    ctx->stats.kernel_launches = ctx->launches;
    if (ctx->profile) {
        double ms_sum[4] = {0.0, 0.0, 0.0, 0.0};
        uint64_t n_gemm = 0;
        FILE* log = nullptr;
        if (const char* path = getenv("CHEQ_GEMM_LOG")) log = fopen(path, "w");
        if (log) fprintf(log, "K,m_tiles,n_tiles,flops,ms,tcgen05\n");
        for (size_t i = 0; i + 1 < ctx->prof_used; i += 2) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ctx->prof_events[i], ctx->prof_events[i + 1]);
            const int cat = ctx->prof_cat[i / 2];
            ms_sum[cat] += ms;
            n_gemm += (cat == 0 || cat == 3);
            if (log && (cat == 0 || cat == 3)) {
                const auto& sh = ctx->prof_shape[i / 2];
                fprintf(log, "%d,%d,%d,%.6e,%.6f,%d\n", sh.K, sh.mt, sh.nt, sh.flops, ms, cat == 3 ? 1 : 0);
            }
        }
        if (log) fclose(log);
        ctx->stats.gemm_ms = ms_sum[0] + ms_sum[3];
        ctx->stats.tc_gemm_ms = ms_sum[3];
        ctx->stats.gemm_launches = n_gemm;
        ctx->stats.gemm_flops = ctx->prof_flops;
        ctx->stats.potrf_ms = ms_sum[1];
        ctx->stats.backsolve_ms = ms_sum[2];
    }
    ctx->stats.ms_total = ms_since(t_total);
    ctx->stats.krylov_iterations = 0;
    ctx->stats.krylov_steps = 0;
    for (const auto& te : ctx->trace)
        if (te.mode == 1) {
            ctx->stats.krylov_iterations += 1;
            ctx->stats.krylov_steps += (uint64_t)te.steps;
        }
    if (io.frames > 1 && io.status_out) return CHEQ_OK;
    return worst;
}

}  // namespace

// ================================================================================================================
extern "C" {

int32_t cheq_ctx_create(int32_t device, cheq_ctx** out) {
    if (!out) return CHEQ_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return CHEQ_ERR_CUDA;   // no CPU fallback by design
    }
    if (device < 0 || device >= count) return CHEQ_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(device) != cudaSuccess) return CHEQ_ERR_CUDA;
    cheq_ctx* ctx = new cheq_ctx();
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return CHEQ_ERR_CUDA;
    }
    for (auto& e : ctx->ev)
        if (cudaEventCreate(&e) != cudaSuccess) {
            delete ctx;
            return CHEQ_ERR_CUDA;
        }
    ctx->cur = ctx->stream;
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&ctx->panel_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_bcast, cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->copy_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->bulk_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_leaf, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming) != cudaSuccess) {
        delete ctx;
        return CHEQ_ERR_CUDA;
    }
    for (int i = 0; i < 64; ++i)
        if (cudaEventCreateWithFlags(&ctx->ev_c[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->ev_b[i], cudaEventDisableTiming) != cudaSuccess) {
            delete ctx;
            return CHEQ_ERR_CUDA;
        }
    if (configure_dense_kernels() != cudaSuccess) {
        cudaGetLastError();
        cheq_ctx_destroy(ctx);
        return CHEQ_ERR_CUDA;
    }
    if (const char* env = getenv("CHEQ_FORCE_LU")) ctx->force_lu = atoi(env) != 0;
    if (const char* env = getenv("CHEQ_OZAKI")) ctx->ozaki_slices = atoi(env);
    if (const char* env = getenv("CHEQ_OZAKI_R")) ctx->ozaki_slices_r = atoi(env);
    if (const char* env = getenv("CHEQ_OZAKI_MIN_K")) ctx->ozaki_min_k = std::max(128, atoi(env));
    if (ctx->ozaki_slices != 0 && (ctx->ozaki_slices < 4 || ctx->ozaki_slices > 8)) ctx->ozaki_slices = 7;
    if (ctx->ozaki_slices_r != 0 && (ctx->ozaki_slices_r < 4 || ctx->ozaki_slices_r > 8)) ctx->ozaki_slices_r = 0;
    if (configure_ozaki_kernels() != cudaSuccess || ctx->oz_err.ensure(64) != cudaSuccess ||
        cudaMemset(ctx->oz_err.p, 0, 64) != cudaSuccess ||
        cudaMallocHost((void**)&ctx->h_oz_err, sizeof(int)) != cudaSuccess) {
        cudaGetLastError();
        cheq_ctx_destroy(ctx);
        return CHEQ_ERR_CUDA;
    }
    *ctx->h_oz_err = 0;
    ctx->ozaki_now = ctx->ozaki_slices;
    if (const char* env = getenv("CHEQ_MORTON")) ctx->morton = atoi(env) != 0;
    if (const char* env = getenv("CHEQ_MATFREE_RESTART")) ctx->matfree_restart = atoi(env);
    if (const char* env = getenv("CHEQ_MATFREE_MAX_STEPS")) ctx->matfree_max_steps = std::max(10, atoi(env));
    if (const char* env = getenv("CHEQ_MATFREE_TOL")) ctx->matfree_tol = atof(env);
    if (const char* env = getenv("CHEQ_KRYLOV")) ctx->krylov_mode = atoi(env) != 0;
    if (const char* env = getenv("CHEQ_KRYLOV_F32")) ctx->krylov_f32 = atoi(env) != 0;
    if (const char* env = getenv("CHEQ_KRYLOV_P2P")) ctx->krylov_p2p = atoi(env);
    if (const char* env = getenv("CHEQ_KRYLOV_MIN_NH")) ctx->krylov_min_nhp = std::max(128, atoi(env));
    if (const char* env = getenv("CHEQ_KRYLOV_FREEZE_EV")) ctx->krylov_freeze_ev = atof(env);
    if (const char* env = getenv("CHEQ_KRYLOV_FORCE_IT")) ctx->krylov_force_it = std::max(2, atoi(env));
    if (const char* env = getenv("CHEQ_PANEL_FLAT")) ctx->panel_flat = atoi(env) != 0;
    if (const char* env = getenv("CHEQ_FACTOR_MODE")) ctx->factor_mode = std::max(0, std::min(2, atoi(env)));
    if (const char* env = getenv("CHEQ_TILES_PER_PANEL")) {
        ctx->tiles_per_panel = std::max(1, std::min(64, atoi(env)));
        ctx->tiles_per_panel_set = true;
    }
    *out = ctx;
    return CHEQ_OK;
}

void cheq_ctx_destroy(cheq_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->panel_stream) cudaStreamSynchronize(ctx->panel_stream);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    ctx->stage.release();
    ctx->own_list.release();
    ctx->col_mask.release();
    ctx->chain.release();
    p2p_close(ctx);
    ctx->arena.release();
    ctx->ipc_stage.release();
    if (ctx->bulk_stream) {
        cudaStreamSynchronize(ctx->bulk_stream);
        cudaStreamDestroy(ctx->bulk_stream);
    }
    for (int i = 0; i < 64; ++i) {
        if (ctx->ev_c[i]) cudaEventDestroy(ctx->ev_c[i]);
        if (ctx->ev_b[i]) cudaEventDestroy(ctx->ev_b[i]);
    }
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->ev_leaf) cudaEventDestroy(ctx->ev_leaf);
    if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->ev_ready) cudaEventDestroy(ctx->ev_ready);
    if (ctx->ev_bcast) cudaEventDestroy(ctx->ev_bcast);
    if (ctx->panel_stream) cudaStreamDestroy(ctx->panel_stream);
    DeviceBuffer* bufs[] = {&ctx->in_xyz, &ctx->in_chi, &ctx->in_hard, &ctx->in_radius, &ctx->in_nq, &ctx->perm,
                            &ctx->type, &ctx->x, &ctx->y, &ctx->z, &ctx->diag, &ctx->chi, &ctx->hard, &ctx->vext,
                            &ctx->A, &ctx->S0, &ctx->Linv, &ctx->v, &ctx->q, &ctx->scratch, &ctx->counters,
                            &ctx->state, &ctx->active, &ctx->blob, &ctx->pc_xyz, &ctx->pc_q, &ctx->pc_type,
                            &ctx->out_q, &ctx->misc0, &ctx->misc1, &ctx->misc2, &ctx->lu_f, &ctx->lu_ut, &ctx->lu_piv, &ctx->sgn, &ctx->sgn_flags, &ctx->xsol};
    for (DeviceBuffer* b : bufs) b->release();
    if (ctx->mf) {
        ctx->mf->release();
        delete ctx->mf;
        ctx->mf = nullptr;
    }
    ctx->kr_R.release();
    ctx->kr_Rf.release();
    ctx->kr_S0t.release();
    ctx->kr_tb.release();
    ctx->kr_wx.release();
    ctx->kx_local.release();
    ctx->kr_part.release();
    ctx->kr_vec.release();
    ctx->kr_sc.release();
    if (ctx->h_ksc) cudaFreeHost(ctx->h_ksc);
    if (ctx->h_state) cudaFreeHost(ctx->h_state);
    if (ctx->h_active) cudaFreeHost(ctx->h_active);
    if (ctx->h_oz_err) cudaFreeHost(ctx->h_oz_err);
    ctx->oz_ws.release();
    ctx->oz_err.release();
    for (auto& e : ctx->ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : ctx->prof_events)
        if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* cheq_last_error(const cheq_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

void cheq_options_default(cheq_options* o) {
    if (!o) return;
    // options.rs:90-100
    o->tolerance = 1.0e-6;
    o->lambda_scale = 0.5;
    o->damping_value = 0.4;
    o->max_iterations = 2000;
    o->hydrogen_scf = 1;
    o->basis = CHEQ_BASIS_STO;
    o->damping_kind = CHEQ_DAMPING_AUTO;
    o->reserved = 0;
}

int32_t cheq_get_stats(const cheq_ctx* ctx, cheq_stats* out) {
    if (!ctx || !out) return CHEQ_ERR_INVALID_ARGUMENT;
    *out = ctx->stats;
    return CHEQ_OK;
}

void* cheq_ctx_stream(const cheq_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int32_t cheq_set_profiling(cheq_ctx* ctx, int32_t enabled) {
    if (!ctx) return CHEQ_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->profile = enabled != 0;
    return CHEQ_OK;
}

int32_t cheq_get_scf_trace(const cheq_ctx* ctx, uint32_t capacity, double* delta_out, int32_t* mode_out,
                           int32_t* steps_out, uint32_t* count_out) {
    if (!ctx || !count_out) return CHEQ_ERR_INVALID_ARGUMENT;
    const uint32_t n = (uint32_t)ctx->trace.size();
    *count_out = n;
    for (uint32_t i = 0; i < n && i < capacity; ++i) {
        if (delta_out) delta_out[i] = ctx->trace[i].delta;
        if (mode_out) mode_out[i] = ctx->trace[i].mode;
        if (steps_out) steps_out[i] = ctx->trace[i].steps;
    }
    return CHEQ_OK;
}

int32_t cheq_set_krylov(cheq_ctx* ctx, int32_t enabled, int32_t min_hydrogen, double freeze_ev, int32_t force_iteration) {
    if (!ctx) return CHEQ_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (enabled >= 0) ctx->krylov_mode = enabled != 0;
    if (min_hydrogen > 0) ctx->krylov_min_nhp = std::max(128, (int)min_hydrogen);
    if (freeze_ev > 0.0) ctx->krylov_freeze_ev = freeze_ev;
    else if (freeze_ev < 0.0) ctx->krylov_freeze_ev = 0.0;   // back to the cost model
    if (force_iteration > 0) ctx->krylov_force_it = std::max(2, (int)force_iteration);
    return CHEQ_OK;
}

int32_t cheq_set_ozaki(cheq_ctx* ctx, int32_t slices, int32_t slices_r, int32_t min_k) {
    if (!ctx) return CHEQ_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if ((slices > 0 && (slices < 4 || slices > 8)) || (slices_r > 0 && (slices_r < 4 || slices_r > 8)))
        return fail(ctx, CHEQ_ERR_INVALID_ARGUMENT, "slice counts must be 0 (off / same) or 4..8");
    if (slices >= 0) ctx->ozaki_slices = slices;
    if (slices_r >= 0) ctx->ozaki_slices_r = slices_r;
    if (min_k > 0) ctx->ozaki_min_k = std::max(128, (int)min_k);
    ctx->ozaki_now = ctx->ozaki_slices;
    return CHEQ_OK;
}

int32_t cheq_dist_unique_id(uint8_t* id_out) {
    if (!id_out) return CHEQ_ERR_INVALID_ARGUMENT;
    std::string err;
    if (!nccl_load(err)) return CHEQ_ERR_CUDA;
    NcclId id;
    if (g_nccl.GetUniqueId(&id) != 0) return CHEQ_ERR_CUDA;
    std::memcpy(id_out, id.internal, CHEQ_DIST_ID_BYTES);
    return CHEQ_OK;
}

int32_t cheq_ctx_dist_init(cheq_ctx* ctx, int32_t rank, int32_t world, const uint8_t* id) {
    if (!ctx) return CHEQ_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->err.clear();
    if (world < 1 || rank < 0 || rank >= world || (world > 1 && !id))
        return fail(ctx, CHEQ_ERR_INVALID_ARGUMENT, "invalid rank / world / id");
    if (ctx->comm) return fail(ctx, CHEQ_ERR_INVALID_ARGUMENT, "context is already part of a group");
    if (world == 1) return CHEQ_OK;
    if (!nccl_load(ctx->err)) return CHEQ_ERR_CUDA;
    CU(cudaSetDevice(ctx->device));
    NcclId nid;
    std::memcpy(nid.internal, id, CHEQ_DIST_ID_BYTES);
    void* comm = nullptr;
    NC(g_nccl.CommInitRank(&comm, world, nid, rank));
    ctx->comm = comm;
    ctx->rank = rank;
    ctx->world = world;
    ctx->pplan = PanelPlan();
    return CHEQ_OK;
}

int32_t cheq_ctx_dist_info(const cheq_ctx* ctx, int32_t* rank_out, int32_t* world_out) {
    if (!ctx) return CHEQ_ERR_INVALID_ARGUMENT;
    if (rank_out) *rank_out = ctx->rank;
    if (world_out) *world_out = ctx->world;
    return CHEQ_OK;
}

int32_t cheq_set_factor_mode(cheq_ctx* ctx, int32_t mode, int32_t tiles_per_panel) {
    if (!ctx || mode < 0 || mode > 2 || tiles_per_panel < 0 || tiles_per_panel > 64) return CHEQ_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->factor_mode = mode;
    if (tiles_per_panel > 0) {
        ctx->tiles_per_panel = tiles_per_panel;
        ctx->tiles_per_panel_set = true;
    }
    return CHEQ_OK;
}

int32_t cheq_dist_plan(int32_t x_tiles, int32_t h_tiles, int32_t tiles_per_panel, int32_t world,
                       int32_t* tile_owner_out, int32_t* tile_panel_out) {
    if (x_tiles < 0 || h_tiles < 0 || tiles_per_panel < 1 || world < 1) return CHEQ_ERR_INVALID_ARGUMENT;
    PanelPlan pp;
    pp.build(x_tiles, h_tiles, tiles_per_panel, world, 0);
    for (size_t p = 0; p < pp.begin.size(); ++p)
        for (int t = pp.begin[p]; t < pp.end[p]; ++t) {
            if (tile_owner_out) tile_owner_out[t] = pp.owner[p];
            if (tile_panel_out) tile_panel_out[t] = (int32_t)p;
        }
    return CHEQ_OK;
}

int32_t cheq_solve(cheq_ctx* ctx, uint64_t n_atoms, const double* xyz, const double* chi, const double* hardness,
                   const double* radius, const uint8_t* n_quantum, double total_charge, const cheq_options* options,
                   const cheq_external* external, double* charges_out, double* potential_out,
                   uint32_t* iterations_out, double* delta_out) {
    if (!ctx) return CHEQ_ERR_INVALID_ARGUMENT;
    SolveIO io;
    io.n = (long)n_atoms;
    io.xyz = xyz;
    io.chi = chi;
    io.hard = hardness;
    io.radius = radius;
    io.nq = n_quantum;
    io.total_charge = total_charge;
    io.opt = options;
    io.ext = external;
    io.charges_out = charges_out;
    io.potential_out = potential_out;
    io.iterations_out = iterations_out;
    io.delta_out = delta_out;
    io.collective = true;
    return solve_impl(ctx, io);
}

int32_t cheq_solve_iterative(cheq_ctx* ctx, uint64_t n_atoms, const double* xyz, const double* chi,
                             const double* hardness, const double* radius, const uint8_t* n_quantum, double total_charge,
                             const cheq_options* options, const cheq_external* external, double* charges_out,
                             double* potential_out, uint32_t* iterations_out, double* delta_out) {
    if (!ctx) return CHEQ_ERR_INVALID_ARGUMENT;
    SolveIO io;
    io.n = (long)n_atoms;
    io.xyz = xyz;
    io.chi = chi;
    io.hard = hardness;
    io.radius = radius;
    io.nq = n_quantum;
    io.total_charge = total_charge;
    io.opt = options;
    io.ext = external;
    io.charges_out = charges_out;
    io.potential_out = potential_out;
    io.iterations_out = iterations_out;
    io.delta_out = delta_out;
    io.collective = true;
    io.path = 2;
    return solve_impl(ctx, io);
}

int32_t cheq_solve_device(cheq_ctx* ctx, uint64_t n_atoms, const double* d_xyz, const double* d_chi,
                          const double* d_hardness, const double* d_radius, const uint8_t* d_n_quantum,
                          double total_charge, const cheq_options* options, const cheq_external* d_external,
                          double* d_charges_out, double* potential_out, uint32_t* iterations_out,
                          double* delta_out) {
    if (!ctx) return CHEQ_ERR_INVALID_ARGUMENT;
    SolveIO io;
    io.device_ptrs = true;
    io.n = (long)n_atoms;
    io.xyz = d_xyz;
    io.chi = d_chi;
    io.hard = d_hardness;
    io.radius = d_radius;
    io.nq = d_n_quantum;
    io.total_charge = total_charge;
    io.opt = options;
    io.ext = d_external;
    io.charges_out = d_charges_out;
    io.potential_out = potential_out;
    io.iterations_out = iterations_out;
    io.delta_out = delta_out;
    io.collective = true;
    return solve_impl(ctx, io);
}

int32_t cheq_solve_batch(cheq_ctx* ctx, uint64_t n_frames, uint64_t n_atoms, const double* xyz, const double* chi,
                         const double* hardness, const double* radius, const uint8_t* n_quantum,
                         double total_charge, const cheq_options* options, double* charges_out,
                         double* potential_out, uint32_t* iterations_out, double* delta_out, int32_t* status_out) {
    if (!ctx) return CHEQ_ERR_INVALID_ARGUMENT;
    if (n_frames == 0) return CHEQ_OK;
    SolveIO io;
    io.n = (long)n_atoms;
    io.frames = (long)n_frames;
    io.xyz = xyz;
    io.chi = chi;
    io.hard = hardness;
    io.radius = radius;
    io.nq = n_quantum;
    io.total_charge = total_charge;
    io.opt = options;
    io.charges_out = charges_out;
    io.potential_out = potential_out;
    io.iterations_out = iterations_out;
    io.delta_out = delta_out;
    io.status_out = status_out;
    return solve_impl(ctx, io);
}

// ---- component entry points -------------------------------------------------------------------------------------

int32_t cheq_pair_integrals(cheq_ctx* ctx, uint64_t count, const double* distance, const uint8_t* n1,
                            const double* radius1, const uint8_t* n2, const double* radius2, double lambda_scale,
                            uint8_t basis, double* out_ev) {
    if (!ctx) return CHEQ_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->err.clear();
    ctx->launches = 0;
    if (count == 0) return CHEQ_OK;
    if (basis > 1) return fail(ctx, CHEQ_ERR_INVALID_ARGUMENT, "unknown basis");
    CU(cudaSetDevice(ctx->device));
    std::vector<ElementType> types;
    std::map<std::pair<uint8_t, uint64_t>, int> index;
    std::vector<int> ti(count), tj(count);
    for (uint64_t k = 0; k < count; ++k) {
        ti[k] = find_or_add_type(types, index, n1[k], radius1[k]);
        tj[k] = find_or_add_type(types, index, n2[k], radius2[k]);
    }
    PackedTables pk;
    int rc = pack_tables(ctx, types, lambda_scale, basis, pk);
    if (rc) return rc;
    PairTableView view;
    rc = upload_tables(ctx, pk, view);
    if (rc) return rc;
    CU(ctx->misc0.ensure(count * 8));
    CU(ctx->misc1.ensure(count * 8));
    CU(ctx->misc2.ensure(count * 8));
    int* d_ti = ctx->misc1.as<int>();
    int* d_tj = d_ti + count;
    cudaStream_t st = ctx->stream;
    CU(cudaMemcpyAsync(ctx->misc0.p, distance, count * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_ti, ti.data(), count * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_tj, tj.data(), count * 4, cudaMemcpyHostToDevice, st));
    LAUNCH(launch_pair_list((long)count, ctx->misc0.as<double>(), d_ti, d_tj, view, ctx->misc2.as<double>(), st));
    CU(cudaMemcpyAsync(out_ev, ctx->misc2.p, count * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return CHEQ_OK;
}

int32_t cheq_assemble_matrix(cheq_ctx* ctx, uint64_t n_atoms, const double* xyz, const double* hardness,
                             const double* radius, const uint8_t* n_quantum, double lambda_scale, uint8_t basis,
                             double* matrix_out) {
    if (!ctx) return CHEQ_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->err.clear();
    ctx->stats = cheq_stats{};
    ctx->launches = 0;
    if (n_atoms == 0) return fail(ctx, CHEQ_ERR_NO_ATOMS, "no atoms");
    LocalCallGuard local_guard(ctx, false);
    CU(cudaSetDevice(ctx->device));
    cheq_options o;
    cheq_options_default(&o);
    o.lambda_scale = lambda_scale;
    o.basis = basis;
    std::vector<double> chi(n_atoms, 0.0);
    SolveIO io;
    io.n = (long)n_atoms;
    io.xyz = xyz;
    io.chi = chi.data();
    io.hard = hardness;
    io.radius = radius;
    io.nq = n_quantum;
    io.opt = &o;
    Prepared P;
    int rc = prepare_and_assemble(ctx, io, 0, 1, true, P, false, true);
    if (rc) return rc;
    const long n = (long)n_atoms;
    CU(ctx->misc0.ensure((size_t)n * n * 8));
    LAUNCH(launch_export_matrix(P.L.NP, P.L.ld, ctx->A.as<double>(), ctx->perm.as<int>(), n, ctx->misc0.as<double>(),
                                ctx->stream));
    CU(cudaMemcpyAsync(matrix_out, ctx->misc0.p, (size_t)n * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]);
    ctx->stats.ms_assemble = ms;
    ctx->stats.kernel_launches = ctx->launches;
    return CHEQ_OK;
}

int32_t cheq_external_potential(cheq_ctx* ctx, uint64_t n_atoms, const double* xyz, const double* radius,
                                const uint8_t* n_quantum, const cheq_external* external, double lambda_scale,
                                uint8_t basis, double* v_out) {
    if (!ctx) return CHEQ_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->err.clear();
    ctx->launches = 0;
    if (n_atoms == 0) return fail(ctx, CHEQ_ERR_NO_ATOMS, "no atoms");
    if (!external) return fail(ctx, CHEQ_ERR_INVALID_ARGUMENT, "external is NULL");
    LocalCallGuard local_guard(ctx, false);
    CU(cudaSetDevice(ctx->device));
    cheq_options o;
    cheq_options_default(&o);
    o.lambda_scale = lambda_scale;
    o.basis = basis;
    const long n = (long)n_atoms;
    std::vector<double> zeros(n, 0.0), ones(n, 1.0);
    cheq_external ext = *external;
    // a zero field with no charges would be "empty"; force evaluation with an explicit flag instead
    SolveIO io;
    io.n = n;
    io.xyz = xyz;
    io.chi = zeros.data();
    io.hard = ones.data();
    io.radius = radius;
    io.nq = n_quantum;
    io.opt = &o;
    io.ext = &ext;
    Prepared P;
    if (external_is_empty(&ext)) {
        for (long i = 0; i < n; ++i) v_out[i] = 0.0;
        return CHEQ_OK;
    }
    int rc = prepare_and_assemble(ctx, io, 0, 1, false, P, false, true);
    if (rc) return rc;
    CU(ctx->out_q.ensure(n * 8));
    LAUNCH(launch_scatter_charges(P.L.NP, ctx->perm.as<int>(), ctx->vext.as<double>(), P.L.NP,
                                  ctx->out_q.as<double>(), n, 1, ctx->stream));
    CU(cudaMemcpyAsync(v_out, ctx->out_q.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return CHEQ_OK;
}

int32_t cheq_dbg_gemm_nt(cheq_ctx* ctx, uint64_t m, uint64_t n, uint64_t k, const double* a, const double* b,
                         double* c, int32_t subtract) {
    if (!ctx) return CHEQ_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->err.clear();
    if (m % kTile || n % kTile || k % kTile || !m || !n || !k)
        return fail(ctx, CHEQ_ERR_INVALID_ARGUMENT, "m, n, k must be positive multiples of 128");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CU(ctx->misc0.ensure(m * k * 8));
    CU(ctx->misc1.ensure(n * k * 8));
    CU(ctx->misc2.ensure(m * n * 8));
    CU(cudaMemcpyAsync(ctx->misc0.p, a, m * k * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->misc1.p, b, n * k * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->misc2.p, c, m * n * 8, cudaMemcpyHostToDevice, st));
    GemmArgs g{};
    g.A = ctx->misc0.as<double>();
    g.lda = (long)m;
    g.B = ctx->misc1.as<double>();
    g.ldb = (long)n;
    g.C = ctx->misc2.as<double>();
    g.ldc = (long)m;
    g.K = (int)k;
    g.lower = 0;
    g.subtract = subtract ? 1 : 0;
    LAUNCH(launch_gemm_nt(g, (int)(m / kTile), (int)(n / kTile), 1, st));
    CU(cudaMemcpyAsync(c, ctx->misc2.p, m * n * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return CHEQ_OK;
}

int32_t cheq_dbg_gemm_ozaki(cheq_ctx* ctx, uint64_t m, uint64_t n, uint64_t k, const double* a, const double* b,
                            double* c, int32_t subtract, int32_t lower, int32_t slices, const double* signs) {
    if (!ctx) return CHEQ_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->err.clear();
    if (m % kTile || n % kTile || k % kTile || !m || !n || !k)
        return fail(ctx, CHEQ_ERR_INVALID_ARGUMENT, "m, n, k must be positive multiples of 128");
    if (!ozaki_supported(slices, (int)std::min<uint64_t>(k, kOzakiMaxK)) || k > (uint64_t)kOzakiMaxK)
        return fail(ctx, CHEQ_ERR_INVALID_ARGUMENT, "4..8 slices, k at most 8192 per launch");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CU(ctx->misc0.ensure(m * k * 8));
    CU(ctx->misc1.ensure(n * k * 8));
    CU(ctx->misc2.ensure(m * n * 8));
    CU(cudaMemcpyAsync(ctx->misc0.p, a, m * k * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->misc1.p, b, n * k * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->misc2.p, c, m * n * 8, cudaMemcpyHostToDevice, st));
    // pivot signs as the factorisation stores them: per 128-wide tile of K a count and the in-tile indices
    std::vector<int> flags((size_t)(k / kTile) * (kTile + 1), 0);
    int* cnt = flags.data();
    int* idx = flags.data() + k / kTile;
    if (signs)
        for (uint64_t i = 0; i < k; ++i)
            if (signs[i] < 0.0) idx[(i / kTile) * kTile + cnt[i / kTile]++] = (int)(i % kTile);
    CU(ctx->sgn_flags.ensure(flags.size() * sizeof(int)));
    CU(cudaMemcpyAsync(ctx->sgn_flags.p, flags.data(), flags.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    GemmArgs g{};
    g.A = ctx->misc0.as<double>();
    g.lda = (long)m;
    g.B = ctx->misc1.as<double>();
    g.ldb = (long)n;
    g.C = ctx->misc2.as<double>();
    g.ldc = (long)m;
    g.K = (int)k;
    g.lower = lower ? 1 : 0;
    g.subtract = subtract ? 1 : 0;
    if (signs) {
        g.neg_cnt = ctx->sgn_flags.as<int>();
        g.neg_idx = ctx->sgn_flags.as<int>() + k / kTile;
    }
    const int mt = (int)(m / kTile), nt = (int)(n / kTile);
    CU(ctx->oz_ws.ensure(ozaki_workspace_bytes(slices, mt, nt, (int)k)));
    LAUNCH(launch_gemm_ozaki(g, mt, nt, slices, ctx->oz_ws.p, ctx->oz_ws.cap, ctx->oz_err.as<int>(), st));
    CU(cudaMemcpyAsync(c, ctx->misc2.p, m * n * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ctx->h_oz_err, ctx->oz_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (*ctx->h_oz_err) {
        cudaMemsetAsync(ctx->oz_err.p, 0, sizeof(int), st);
        return fail(ctx, CHEQ_ERR_CUDA, "tcgen05 slice-product pipeline timed out");
    }
    return CHEQ_OK;
}

int32_t cheq_dbg_cholesky_solve(cheq_ctx* ctx, uint64_t n_, const double* a, uint64_t nrhs, const double* b,
                                double* x_out, double* l_out) {
    if (!ctx) return CHEQ_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->err.clear();
    ctx->launches = 0;
    if (n_ == 0 || nrhs == 0 || nrhs > 2) return fail(ctx, CHEQ_ERR_INVALID_ARGUMENT, "need n > 0 and 1 <= nrhs <= 2");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const long n = (long)n_;
    const int NP = round_up_tile((int)n), MT = NP + kRhsRows;
    const long ld = MT;
    // host staging of the padded trapezoid (test path only)
    std::vector<double> h((size_t)ld * NP, 0.0);
    for (long j = 0; j < NP; ++j) {
        if (j < n) {
            for (long i = j; i < n; ++i) h[i + j * ld] = a[i + j * n];
            for (uint64_t r = 0; r < nrhs; ++r) h[NP + r + j * ld] = b[j + r * n];
        } else {
            h[j + j * ld] = 1.0;
        }
    }
    CU(ctx->A.ensure(h.size() * 8));
    CU(ctx->Linv.ensure((size_t)NP * kTile * 8));
    CU(ctx->v.ensure(NP * 8));
    CU(ctx->xsol.ensure(NP * 8));
    CU(ctx->scratch.ensure((size_t)backsolve_max_ctas() * kTile * 8));
    CU(ctx->counters.ensure(sizeof(unsigned)));
    CU(ctx->state.ensure(sizeof(ScfState)));
    int rc = ensure_host_state(ctx, 1);
    if (rc) return rc;
    CU(cudaMemsetAsync(ctx->counters.p, 0, sizeof(unsigned), st));
    CU(cudaMemcpyAsync(ctx->A.p, h.data(), h.size() * 8, cudaMemcpyHostToDevice, st));
    LAUNCH(launch_scf_init(ctx->state.as<ScfState>(), 1.0, 1, st));
    CU(ctx->sgn.ensure(NP * 8));
    CU(ctx->sgn_flags.ensure((size_t)(NP / kTile) * (kTile + 1) * sizeof(int)));
    Dense D{ctx, ctx->A.as<double>(), ld, (long)ld * NP, ctx->Linv.as<double>(), (long)NP * kTile, MT,
            ctx->state.as<ScfState>(), 1, ctx->sgn.as<double>(), (long)NP, ctx->sgn_flags.as<int>(),
            ctx->sgn_flags.as<int>() + NP / kTile, (long)NP / kTile};
    rc = factor_panel(D, 0, NP);
    if (rc) return rc;
    std::vector<double> xv(NP);
    for (uint64_t r = 0; r < nrhs; ++r) {
        // v = forward-substituted rhs r (row NP + r)
        CU(cudaMemcpy2DAsync(ctx->v.p, 8, ctx->A.as<double>() + NP + r, ld * 8, 8, NP, cudaMemcpyDeviceToDevice, st));
        rc = backsolve(D, NP, ctx->v.as<double>(), ctx->xsol.as<double>(), NP);
        if (rc) return rc;
        CU(cudaMemcpyAsync(xv.data(), ctx->xsol.p, NP * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (long i = 0; i < n; ++i) x_out[i + r * n] = xv[i];
    }
    CU(cudaMemcpyAsync(ctx->h_state, ctx->state.p, sizeof(ScfState), cudaMemcpyDeviceToHost, st));
    if (l_out) {
        CU(cudaMemcpyAsync(h.data(), ctx->A.p, h.size() * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (long j = 0; j < n; ++j)
            for (long i = 0; i < n; ++i) l_out[i + j * n] = (i >= j) ? h[i + j * ld] : 0.0;
    }
    CU(cudaStreamSynchronize(st));
    ctx->stats.kernel_launches = ctx->launches;
    if (ctx->h_state[0].info != 0) return fail(ctx, CHEQ_ERR_LINALG, "matrix is not positive definite");
    return CHEQ_OK;
}

int32_t cheq_dbg_gemm_bench(cheq_ctx* ctx, uint64_t m, uint64_t n, uint64_t k, int32_t lower, int32_t subtract,
                            int32_t reps, double* ms_out) {
    if (!ctx || !ms_out) return CHEQ_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->err.clear();
    if (m % kTile || n % kTile || k % kTile || !m || !n || !k || reps < 1)
        return fail(ctx, CHEQ_ERR_INVALID_ARGUMENT, "m, n, k must be positive multiples of 128");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CU(ctx->misc0.ensure(m * k * 8));
    CU(ctx->misc1.ensure(n * k * 8));
    CU(ctx->misc2.ensure(m * n * 8));
    CU(cudaMemsetAsync(ctx->misc0.p, 0x3f, m * k * 8, st));   // every double = 0.000473...
    CU(cudaMemsetAsync(ctx->misc1.p, 0x3f, n * k * 8, st));
    CU(cudaMemsetAsync(ctx->misc2.p, 0x3f, m * n * 8, st));
    GemmArgs g{};
    g.A = ctx->misc0.as<double>();
    g.lda = (long)m;
    g.B = ctx->misc1.as<double>();
    g.ldb = (long)n;
    g.C = ctx->misc2.as<double>();
    g.ldc = (long)m;
    g.K = (int)k;
    g.lower = lower & 1;
    g.subtract = subtract ? 1 : 0;
    if (lower & 0x30) set_gemm_warps((lower & 0x10) ? 8 : 16);   // debug: bit 4 -> 8 warps, bit 5 -> 16 warps
    if (lower & 0xC0) set_gemm_warps((lower & 0x40) ? 64 : 128);  // debug: bit 6 -> 64-row tiles, bit 7 -> 128-row
    LAUNCH(launch_gemm_nt(g, (int)(m / kTile), (int)(n / kTile), 1, st));   // warm-up
    CU(cudaEventRecord(ctx->ev[0], st));
    for (int r = 0; r < reps; ++r) LAUNCH(launch_gemm_nt(g, (int)(m / kTile), (int)(n / kTile), 1, st));
    CU(cudaEventRecord(ctx->ev[1], st));
    CU(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    *ms_out = ms / reps;
    return CHEQ_OK;
}

int32_t cheq_dbg_potrf_cycles(cheq_ctx* ctx, uint64_t* out8, int32_t reset) {
    if (!ctx || !out8) return CHEQ_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(potrf_debug_cycles((unsigned long long*)out8, reset));
    return CHEQ_OK;
}

int32_t cheq_dbg_dfma_rate(cheq_ctx* ctx, double* tflops_out) {
    if (!ctx || !tflops_out) return CHEQ_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->err.clear();
    CU(cudaSetDevice(ctx->device));
    const int blocks = 148 * 8, iters = 1 << 15;
    CU(ctx->misc0.ensure((size_t)blocks * 256 * 8));
    LAUNCH(launch_dfma_rate(ctx->misc0.as<double>(), blocks, 1024, ctx->stream));
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    LAUNCH(launch_dfma_rate(ctx->misc0.as<double>(), blocks, iters, ctx->stream));
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    *tflops_out = 2.0 * 8.0 * iters * (double)blocks * 256.0 / (ms * 1e-3) / 1e12;
    // DMMA issue rate at 2, 4 and 8 warps per SM sub-partition, reported through the error string (debug aid)
    std::string report;
    for (int threads : {128, 256, 512}) {
        const int b2 = 148, it2 = 1 << 14;
        LAUNCH(launch_dmma_rate(ctx->misc0.as<double>(), b2, threads, 256, ctx->stream));
        CU(cudaEventRecord(ctx->ev[0], ctx->stream));
        LAUNCH(launch_dmma_rate(ctx->misc0.as<double>(), b2, threads, it2, ctx->stream));
        CU(cudaEventRecord(ctx->ev[1], ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
        const double tf = 512.0 * 16.0 * it2 * (double)b2 * (threads / 32) / (ms * 1e-3) / 1e12;
        report += "dmma " + std::to_string(threads / 128) + " warps/SMSP: " + std::to_string(tf) + " TF/s; ";
    }
    {
        const int iters = 4096;
        long long* d_cyc = (long long*)(ctx->misc0.as<double>() + 64);
        LAUNCH(launch_latency_probe(ctx->misc0.as<double>(), d_cyc, iters, ctx->stream));
        long long h[5];
        CU(cudaMemcpyAsync(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        char buf[256];
        snprintf(buf, sizeof buf, "latency cycles: dfma %.1f shfl64 %.1f dmma %.1f lds+dadd %.1f rsqrtf+fadd %.1f",
                 (double)h[0] / iters, (double)h[1] / iters, (double)h[2] / iters, (double)h[3] / iters,
                 (double)h[4] / iters);
        report += buf;
    }
    ctx->err = report;
    return CHEQ_OK;
}

double cheq_host_sto_table_eval(int32_t n1, double zeta1, int32_t n2, double zeta2, double distance_bohr,
                                int32_t* kind_out, int32_t* terms_out, int32_t* coefs_out, double* max_abs_err_out) {
    static std::mutex mu;
    static std::map<TableKey, StoPairTable> cache;
    std::lock_guard<std::mutex> lock(mu);
    TableKey key{n1, n2, bits(zeta1), bits(zeta2)};
    auto it = cache.find(key);
    if (it == cache.end()) it = cache.emplace(key, build_sto_pair_table(n1, zeta1, n2, zeta2)).first;
    const StoPairTable& t = it->second;
    if (kind_out) *kind_out = t.kind;
    if (terms_out) *terms_out = (int32_t)t.terms.size();
    if (coefs_out) {
        int c = 0;
        for (const StoTerm& term : t.terms) c += (int)term.coef.size();
        *coefs_out = c;
    }
    if (max_abs_err_out) *max_abs_err_out = t.max_abs_err;
    return eval_sto_pair_table(t, distance_bohr);
}

}  // extern "C"
