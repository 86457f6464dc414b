// cheq_b200.hpp -- C++ host-side mirror of cheq's public API over the C ABI (cheq_b200.h).
//
// The reference is a Rust crate (compiled code) and this image has no Rust toolchain, so the host side above the
// C ABI is provided in C++ with the reference's names, argument meaning and error behaviour
// (/root/reference/src/lib.rs:98-104):
//   cheq::Atom, cheq::PointCharge, cheq::ExternalPotential      src/types.rs:35-40, 84-97, 136-292
//   cheq::SolverOptions, BasisType, DampingStrategy             src/solver/options.rs:9-101
//   cheq::ElementData, cheq::Parameters (TOML loader)           src/params.rs:21-59, 89-133, 182-348
//   cheq::CheqError                                              src/error.rs:18-68
//   cheq::QEqSolver::{solve, solve_in_field, with_options}       src/solver/implementation.rs:60-266
//   cheq::get_default_parameters                                 src/lib.rs:127-133
// Header-only; link with -lcheq_b200.  All numerics happen on the GPU behind cheq_solve.
#pragma once
#include <array>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "cheq_b200.h"

namespace cheq {

// ---- error.rs -------------------------------------------------------------------------------------------------
class CheqError : public std::runtime_error {
  public:
    enum class Kind { ParameterNotFound, NotConverged, LinalgError, IoError, DeserializationError, NoAtoms, Cuda };
    Kind kind;
    uint8_t atomic_number = 0;     // ParameterNotFound
    uint32_t max_iterations = 0;   // NotConverged
    double delta = 0.0;            // NotConverged
    CheqError(Kind k, const std::string& msg) : std::runtime_error(msg), kind(k) {}
    static CheqError parameter_not_found(uint8_t z) {
        CheqError e(Kind::ParameterNotFound,
                    "Atomic parameters not found for element with atomic number: " + std::to_string((int)z));
        e.atomic_number = z;
        return e;
    }
    static CheqError not_converged(uint32_t max_iterations, double delta) {
        char buf[160];
        std::snprintf(buf, sizeof buf, "SCF failed to converge after %u iterations. Final charge difference: %.2e",
                      max_iterations, delta);
        CheqError e(Kind::NotConverged, buf);
        e.max_iterations = max_iterations;
        e.delta = delta;
        return e;
    }
    static CheqError linalg(const std::string& m) {
        return CheqError(Kind::LinalgError, "Failed to solve the linear matrix system: " + m);
    }
    static CheqError no_atoms() {
        return CheqError(Kind::NoAtoms, "Input validation failed: at least one atom is required for a calculation");
    }
};

// ---- types.rs -------------------------------------------------------------------------------------------------
struct Atom {
    uint8_t atomic_number;
    std::array<double, 3> position;
};
// AtomView: any type with `uint8_t atomic_number() const` / `std::array<double,3> position() const` members or the
// two public fields of Atom works with QEqSolver::solve (resolved by the overloads below).
inline uint8_t atomic_number_of(const Atom& a) { return a.atomic_number; }
inline std::array<double, 3> position_of(const Atom& a) { return a.position; }
template <class A>
auto atomic_number_of(const A& a) -> decltype(a.atomic_number()) { return a.atomic_number(); }
template <class A>
auto position_of(const A& a) -> decltype(a.position()) { return a.position(); }

struct CalculationResult {
    std::vector<double> charges;
    double equilibrated_potential;
    uint32_t iterations;
};

struct PointCharge {
    uint8_t atomic_number;
    std::array<double, 3> position;
    double charge;
};

class ExternalPotential {
    std::vector<PointCharge> point_charges_;
    std::array<double, 3> uniform_field_{0.0, 0.0, 0.0};

  public:
    ExternalPotential() = default;
    static ExternalPotential from_point_charges(std::vector<PointCharge> c) {
        ExternalPotential e;
        e.point_charges_ = std::move(c);
        return e;
    }
    static ExternalPotential from_uniform_field(std::array<double, 3> f) {
        ExternalPotential e;
        e.uniform_field_ = f;
        return e;
    }
    ExternalPotential with_point_charges(std::vector<PointCharge> c) const {
        ExternalPotential e = *this;
        e.point_charges_ = std::move(c);
        return e;
    }
    ExternalPotential with_uniform_field(std::array<double, 3> f) const {
        ExternalPotential e = *this;
        e.uniform_field_ = f;
        return e;
    }
    const std::vector<PointCharge>& point_charges() const { return point_charges_; }
    std::array<double, 3> uniform_field() const { return uniform_field_; }
    bool is_empty() const {   // types.rs:281-286
        return point_charges_.empty() && uniform_field_[0] == 0.0 && uniform_field_[1] == 0.0 &&
               uniform_field_[2] == 0.0;
    }
    size_t num_point_charges() const { return point_charges_.size(); }
};

// ---- options.rs -----------------------------------------------------------------------------------------------
enum class BasisType { Gto = 0, Sto = 1 };
struct DampingStrategy {
    enum Kind { None = 0, Fixed = 1, Auto = 2 } kind = Auto;
    double value = 0.4;
    static DampingStrategy none() { return {None, 1.0}; }
    static DampingStrategy fixed(double d) { return {Fixed, d}; }
    static DampingStrategy automatic(double initial = 0.4) { return {Auto, initial}; }
};
struct SolverOptions {   // options.rs:90-100
    double tolerance = 1.0e-6;
    uint32_t max_iterations = 2000;
    double lambda_scale = 0.5;
    bool hydrogen_scf = true;
    BasisType basis_type = BasisType::Sto;
    DampingStrategy damping = DampingStrategy::automatic(0.4);
};

// ---- params.rs ------------------------------------------------------------------------------------------------
struct ElementData {
    double electronegativity;   // chi
    double hardness;            // j
    double radius;
    uint8_t principal_quantum_number;   // n
    bool operator==(const ElementData& o) const {
        return electronegativity == o.electronegativity && hardness == o.hardness && radius == o.radius &&
               principal_quantum_number == o.principal_quantum_number;
    }
};

inline int element_symbol_to_atomic_number(const std::string& s) {   // params.rs:226-348 (case-sensitive)
    static const char* sym[] = {"H","He","Li","Be","B","C","N","O","F","Ne","Na","Mg","Al","Si","P","S","Cl","Ar","K","Ca",
        "Sc","Ti","V","Cr","Mn","Fe","Co","Ni","Cu","Zn","Ga","Ge","As","Se","Br","Kr","Rb","Sr","Y","Zr","Nb","Mo","Tc",
        "Ru","Rh","Pd","Ag","Cd","In","Sn","Sb","Te","I","Xe","Cs","Ba","La","Ce","Pr","Nd","Pm","Sm","Eu","Gd","Tb","Dy",
        "Ho","Er","Tm","Yb","Lu","Hf","Ta","W","Re","Os","Ir","Pt","Au","Hg","Tl","Pb","Bi","Po","At","Rn","Fr","Ra","Ac",
        "Th","Pa","U","Np","Pu","Am","Cm","Bk","Cf","Es","Fm","Md","No","Lr","Rf","Db","Sg","Bh","Hs","Mt","Ds","Rg","Cn",
        "Nh","Fl","Mc","Lv","Ts","Og"};
    for (int i = 0; i < 118; ++i)
        if (s == sym[i]) return i + 1;
    return -1;
}

struct Parameters {
    std::map<uint8_t, ElementData> elements;

    // Minimal TOML reader for cheq's parameter files: one `[elements]` table whose entries are inline tables
    // `"O" = { chi = 8.741, j = 13.364, radius = 0.669, n = 2 }` keyed by symbol or atomic number.
    static Parameters load_from_str(const std::string& toml) {
        Parameters p;
        std::istringstream in(toml);
        std::string line;
        bool in_elements = false, saw_table = false;
        auto fail = [](const std::string& m) {
            return CheqError(CheqError::Kind::DeserializationError, "Failed to deserialize TOML parameters: " + m);
        };
        auto trim = [](std::string s) {
            size_t a = 0, b = s.size();
            while (a < b && std::isspace((unsigned char)s[a])) ++a;
            while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
            return s.substr(a, b - a);
        };
        while (std::getline(in, line)) {
            bool quoted = false;   // strip comments outside quotes
            for (size_t i = 0; i < line.size(); ++i) {
                if (line[i] == '"') quoted = !quoted;
                if (line[i] == '#' && !quoted) {
                    line.resize(i);
                    break;
                }
            }
            line = trim(line);
            if (line.empty()) continue;
            if (line.front() == '[') {
                if (line.back() != ']') throw fail("invalid table header: " + line);
                in_elements = trim(line.substr(1, line.size() - 2)) == "elements";
                saw_table |= in_elements;
                continue;
            }
            const size_t eq = line.find('=');
            if (eq == std::string::npos) throw fail("expected `key = value`: " + line);
            if (!in_elements) continue;
            std::string key = trim(line.substr(0, eq));
            if (key.size() >= 2 && key.front() == '"' && key.back() == '"') key = key.substr(1, key.size() - 2);
            std::string val = trim(line.substr(eq + 1));
            if (val.size() < 2 || val.front() != '{' || val.back() != '}') throw fail("expected an inline table for '" + key + "'");
            int z = -1;
            bool digits = !key.empty();
            for (char c : key) digits &= std::isdigit((unsigned char)c) != 0;
            if (digits && key.size() <= 3 && std::stoi(key) <= 255) z = std::stoi(key);
            if (z < 0) z = element_symbol_to_atomic_number(key);
            if (z < 0) throw fail("invalid element key: '" + key + "'");   // params.rs:203
            std::map<std::string, std::string> f;
            std::istringstream fields(val.substr(1, val.size() - 2));
            std::string item;
            while (std::getline(fields, item, ',')) {
                const size_t e2 = item.find('=');
                if (e2 == std::string::npos) {
                    if (trim(item).empty()) continue;
                    throw fail("malformed inline table for '" + key + "'");
                }
                f[trim(item.substr(0, e2))] = trim(item.substr(e2 + 1));
            }
            ElementData d{};
            for (const char* name : {"chi", "j", "radius", "n"})
                if (!f.count(name)) throw fail(std::string("missing field `") + name + "`");
            try {
                size_t pos = 0;
                d.electronegativity = std::stod(f["chi"], &pos);
                d.hardness = std::stod(f["j"], &pos);
                d.radius = std::stod(f["radius"], &pos);
                const long n = std::stol(f["n"], &pos);
                if (pos != f["n"].size() || n < 0 || n > 255) throw std::invalid_argument("n");
                d.principal_quantum_number = (uint8_t)n;
            } catch (const std::exception&) {
                throw fail("invalid value in element '" + key + "'");
            }
            p.elements[(uint8_t)z] = d;
        }
        if (!saw_table) throw fail("missing field `elements`");
        return p;
    }
    static Parameters load_from_file(const std::string& path) {
        std::ifstream in(path);
        if (!in) throw CheqError(CheqError::Kind::IoError, "I/O error at path '" + path + "': cannot open file");
        std::stringstream ss;
        ss << in.rdbuf();
        return load_from_str(ss.str());
    }
};

// lib.rs:127-133.  The table lives in cheq_b200/data/qeq_params.toml; `path` overrides the default location.
inline const Parameters& get_default_parameters(const std::string& path = "") {
    static std::once_flag once;
    static Parameters cached;
    std::call_once(once, [&] {
        const char* env = std::getenv("CHEQ_B200_PARAMS");
        cached = Parameters::load_from_file(!path.empty() ? path : env ? env : "cheq_b200/data/qeq_params.toml");
    });
    return cached;
}

// ---- solver/implementation.rs -----------------------------------------------------------------------------------
class QEqSolver {
    const Parameters* parameters_;
    SolverOptions options_;
    std::shared_ptr<cheq_ctx> ctx_;   // one CUDA stream + workspace; the library serialises calls on it

    void ensure_ctx(int device) {
        if (ctx_) return;
        cheq_ctx* raw = nullptr;
        if (cheq_ctx_create(device, &raw) != CHEQ_OK)
            throw CheqError(CheqError::Kind::Cuda, "cheq_ctx_create failed: a CUDA device is required (no CPU fallback)");
        ctx_ = std::shared_ptr<cheq_ctx>(raw, cheq_ctx_destroy);
    }

  public:
    explicit QEqSolver(const Parameters& parameters) : parameters_(&parameters) {}   // QEqSolver::new, :87-92
    QEqSolver with_options(SolverOptions o) const {                                   // :122-125
        QEqSolver s = *this;
        s.options_ = o;
        return s;
    }
    const SolverOptions& options() const { return options_; }

    // ---- one dense system over the GPUs of a node (no reference equivalent; include/cheq_b200.h) ----
    // Rank 0 creates the id, the caller ships the 128 bytes to every rank (MPI_Bcast, a file, ...), every rank
    // joins with its own solver; afterwards solve / solve_in_field are collective.
    static std::array<uint8_t, CHEQ_DIST_ID_BYTES> dist_unique_id() {
        std::array<uint8_t, CHEQ_DIST_ID_BYTES> id{};
        if (cheq_dist_unique_id(id.data()) != CHEQ_OK)
            throw CheqError(CheqError::Kind::Cuda, "cheq_dist_unique_id failed (NCCL not loadable)");
        return id;
    }
    void join_group(int rank, int world, const std::array<uint8_t, CHEQ_DIST_ID_BYTES>& id, int device) {
        ensure_ctx(device);
        if (cheq_ctx_dist_init(ctx_.get(), rank, world, id.data()) != CHEQ_OK)
            throw CheqError(CheqError::Kind::Cuda, std::string("cheq_ctx_dist_init: ") + cheq_last_error(ctx_.get()));
    }

    template <class A>
    CalculationResult solve(const std::vector<A>& atoms, double total_charge, int device = 0) {   // :174-188
        return solve_impl(atoms, total_charge, nullptr, device);
    }
    template <class A>
    CalculationResult solve_in_field(const std::vector<A>& atoms, double total_charge, const ExternalPotential& ext,
                                     int device = 0) {                                             // :239-266
        return solve_impl(atoms, total_charge, ext.is_empty() ? nullptr : &ext, device);
    }

  private:
    template <class A>
    CalculationResult solve_impl(const std::vector<A>& atoms, double total_charge, const ExternalPotential* ext,
                                 int device) {
        const size_t n = atoms.size();
        if (n == 0) throw CheqError::no_atoms();
        std::vector<double> xyz(3 * n), chi(n), hard(n), rad(n);
        std::vector<uint8_t> nq(n);
        for (size_t i = 0; i < n; ++i) {   // fetch_element_data, :348-362
            const uint8_t z = atomic_number_of(atoms[i]);
            auto it = parameters_->elements.find(z);
            if (it == parameters_->elements.end()) throw CheqError::parameter_not_found(z);
            const auto p = position_of(atoms[i]);
            xyz[3 * i] = p[0];
            xyz[3 * i + 1] = p[1];
            xyz[3 * i + 2] = p[2];
            chi[i] = it->second.electronegativity;
            hard[i] = it->second.hardness;
            rad[i] = it->second.radius;
            nq[i] = it->second.principal_quantum_number;
        }
        std::vector<double> pxyz, pq, prad;
        std::vector<uint8_t> pn;
        cheq_external e{};
        if (ext) {   // :378-387
            for (const PointCharge& pc : ext->point_charges()) {
                auto it = parameters_->elements.find(pc.atomic_number);
                if (it == parameters_->elements.end()) throw CheqError::parameter_not_found(pc.atomic_number);
                pxyz.insert(pxyz.end(), pc.position.begin(), pc.position.end());
                pq.push_back(pc.charge);
                prad.push_back(it->second.radius);
                pn.push_back(it->second.principal_quantum_number);
            }
            e.num_point_charges = pq.size();
            e.xyz = pxyz.data();
            e.charge = pq.data();
            e.radius = prad.data();
            e.n = pn.data();
            const auto f = ext->uniform_field();
            e.field[0] = f[0];
            e.field[1] = f[1];
            e.field[2] = f[2];
        }
        cheq_options o;
        cheq_options_default(&o);
        o.tolerance = options_.tolerance;
        o.lambda_scale = options_.lambda_scale;
        o.damping_value = options_.damping.value;
        o.max_iterations = options_.max_iterations;
        o.hydrogen_scf = options_.hydrogen_scf ? 1 : 0;
        o.basis = (uint8_t)options_.basis_type;
        o.damping_kind = (uint8_t)options_.damping.kind;
        ensure_ctx(device);
        CalculationResult r;
        r.charges.resize(n);
        double delta = 0.0;
        const int rc = cheq_solve(ctx_.get(), n, xyz.data(), chi.data(), hard.data(), rad.data(), nq.data(), total_charge,
                                  &o, ext ? &e : nullptr, r.charges.data(), &r.equilibrated_potential, &r.iterations,
                                  &delta);
        switch (rc) {
            case CHEQ_OK: return r;
            case CHEQ_ERR_NO_ATOMS: throw CheqError::no_atoms();
            case CHEQ_ERR_NOT_CONVERGED: throw CheqError::not_converged(options_.max_iterations, delta);
            case CHEQ_ERR_LINALG: throw CheqError::linalg(cheq_last_error(ctx_.get()));
            default: throw CheqError(CheqError::Kind::Cuda, cheq_last_error(ctx_.get()));
        }
    }
};

}  // namespace cheq
