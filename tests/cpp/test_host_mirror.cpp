// C++ host-mirror test, driven by tests/test_cpp_mirror.py.  Mode "cpu": everything that is decided before the
// FFI call (TOML loader, element lookup errors, builders).  Mode "gpu": the reference's own unit tests
// (src/solver/implementation.rs:620-1013) through cheq::QEqSolver on cuda:0.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "cheq_b200.hpp"

#define CHECK(cond)                                                         \
    do {                                                                    \
        if (!(cond)) {                                                      \
            std::fprintf(stderr, "CHECK failed at line %d: %s\n", __LINE__, #cond); \
            return 1;                                                       \
        }                                                                   \
    } while (0)

using namespace cheq;

static int cpu_tests(const Parameters& P) {
    // params.rs:350-501
    const char* toml =
        "[elements]\n\"1\" = { chi = 2.20, j = 13.60, radius = 0.37, n = 1 } # Hydrogen by atomic number\n"
        "\"Fe\" = { chi = 1.83, j = 7.90, radius = 1.26, n = 4 }\n\"O\" = { chi = 3.44, j = 13.62, radius = 0.60, n = 2 }\n";
    Parameters p = Parameters::load_from_str(toml);
    CHECK(p.elements.size() == 3 && p.elements.count(26) && p.elements[26].hardness == 7.90 &&
          p.elements[26].principal_quantum_number == 4 && p.elements[1].electronegativity == 2.20);
    try {
        Parameters::load_from_str("[elements]\n\"InvalidKey\" = { chi = 1.0, j = 1.0, radius = 1.0, n = 1 }\n");
        CHECK(false);
    } catch (const CheqError& e) {
        CHECK(e.kind == CheqError::Kind::DeserializationError &&
              std::strstr(e.what(), "invalid element key: 'InvalidKey'") != nullptr);
    }
    try {
        Parameters::load_from_str("[elements]\n\"1\" = { chi = 2.20, j = 13.60, radius = 0.37 }\n");
        CHECK(false);
    } catch (const CheqError& e) {
        CHECK(e.kind == CheqError::Kind::DeserializationError);
    }
    try {
        Parameters::load_from_str("this is not valid toml");
        CHECK(false);
    } catch (const CheqError& e) {
        CHECK(e.kind == CheqError::Kind::DeserializationError);
    }
    try {
        Parameters::load_from_file("non_existent_file.toml");
        CHECK(false);
    } catch (const CheqError& e) {
        CHECK(e.kind == CheqError::Kind::IoError);
    }
    CHECK(P.elements.size() == 103 && P.elements.at(1).hardness == 13.8904 && P.elements.at(8).radius == 0.669);
    CHECK(&get_default_parameters() == &P);   // cached (lib.rs:135-157)
    // types.rs:294-416
    CHECK(ExternalPotential().is_empty());
    auto e1 = ExternalPotential::from_point_charges({{8, {1.0, 0.0, 0.0}, -0.5}, {1, {2.0, 0.0, 0.0}, 0.25}});
    CHECK(!e1.is_empty() && e1.num_point_charges() == 2 && e1.point_charges()[1].charge == 0.25);
    auto e2 = ExternalPotential::from_uniform_field({0.0, 0.0, 0.001});
    CHECK(!e2.is_empty() && e2.num_point_charges() == 0);
    auto e3 = ExternalPotential().with_point_charges({{6, {0, 0, 0}, 0.1}}).with_uniform_field({0, 0, 1.0});
    CHECK(e3.num_point_charges() == 1 && e3.uniform_field()[2] == 1.0);
    CHECK(!ExternalPotential::from_point_charges({{1, {0, 0, 0}, 0.0}}).is_empty());
    // errors decided on the host (implementation.rs:180-182, 348-362, 378-387)
    QEqSolver solver(P);
    try {
        solver.solve(std::vector<Atom>{}, 0.0);
        CHECK(false);
    } catch (const CheqError& e) {
        CHECK(e.kind == CheqError::Kind::NoAtoms);
    }
    try {
        solver.solve(std::vector<Atom>{{118, {0, 0, 0}}}, 0.0);
        CHECK(false);
    } catch (const CheqError& e) {
        CHECK(e.kind == CheqError::Kind::ParameterNotFound && e.atomic_number == 118);
    }
    try {
        solver.solve_in_field(std::vector<Atom>{{6, {0, 0, 0}}}, 0.0,
                              ExternalPotential::from_point_charges({{200, {3.0, 0, 0}, 0.5}}));
        CHECK(false);
    } catch (const CheqError& e) {
        CHECK(e.kind == CheqError::Kind::ParameterNotFound && e.atomic_number == 200);
    }
    CHECK(std::string(CheqError::not_converged(100, 1.234e-3).what()) ==
          "SCF failed to converge after 100 iterations. Final charge difference: 1.23e-03");
    SolverOptions o;
    CHECK(o.tolerance == 1e-6 && o.max_iterations == 2000 && o.lambda_scale == 0.5 && o.hydrogen_scf &&
          o.basis_type == BasisType::Sto && o.damping.kind == DampingStrategy::Auto && o.damping.value == 0.4);
    return 0;
}

static int gpu_tests(const Parameters& P) {
    QEqSolver solver(P);
    auto r = solver.solve(std::vector<Atom>{{1, {0, 0, 0}}, {1, {0.74, 0, 0}}}, 0.0);
    CHECK(std::fabs(r.charges[0]) < 1e-6 && std::fabs(r.charges[0] - r.charges[1]) < 1e-10);
    std::vector<Atom> hf{{1, {0, 0, 0}}, {9, {0.917, 0, 0}}};
    r = solver.solve(hf, 0.0);
    CHECK(r.charges[0] > 0 && r.charges[1] < 0 && std::fabs(r.charges[0] + r.charges[1]) < 1e-6);
    // water, SURVEY App. B pinned values: 9 iterations, potential -6.721479 eV, q_H 0.3251
    const double h = 104.5 * M_PI / 180.0 / 2.0;
    std::vector<Atom> w{{8, {0, 0, 0}}, {1, {0.958 * std::cos(h), 0.958 * std::sin(h), 0}}, {1, {0.958 * std::cos(h), -0.958 * std::sin(h), 0}}};
    r = solver.solve(w, 0.0);
    CHECK(r.iterations == 9 && std::fabs(r.equilibrated_potential + 6.721479) < 1e-6 && std::fabs(r.charges[1] - 0.3251) < 5e-5);
    try {
        SolverOptions o;
        o.max_iterations = 1;
        solver.with_options(o).solve(std::vector<Atom>{{3, {0, 0, 0}}, {1, {1.595, 0, 0}}}, 0.0);
        CHECK(false);
    } catch (const CheqError& e) {
        CHECK(e.kind == CheqError::Kind::NotConverged && e.max_iterations == 1);
    }
    auto vac = solver.solve(std::vector<Atom>{{6, {0, 0, 0}}, {8, {1.2, 0, 0}}}, 0.0);
    auto fld = solver.solve_in_field(std::vector<Atom>{{6, {0, 0, 0}}, {8, {1.2, 0, 0}}}, 0.0,
                                     ExternalPotential::from_point_charges({{7, {3.0, 0, 0}, 0.5}}));
    CHECK(fld.charges[1] < vac.charges[1] && std::fabs(fld.charges[0] + fld.charges[1]) < 1e-6);
    auto same = solver.solve_in_field(hf, 0.0, ExternalPotential());
    auto base = solver.solve(hf, 0.0);
    CHECK(std::fabs(same.charges[0] - base.charges[0]) < 1e-10);
    SolverOptions g;
    g.basis_type = BasisType::Gto;
    r = solver.with_options(g).solve(hf, 0.0);
    CHECK(r.charges[0] > 0 && r.charges[1] < 0);
    std::printf("gpu ok\n");
    return 0;
}

int main(int argc, char** argv) {
    const std::string mode = argc > 1 ? argv[1] : "cpu";
    const std::string params = argc > 2 ? argv[2] : "cheq_b200/data/qeq_params.toml";
    try {
        const Parameters& P = get_default_parameters(params);
        const int rc = (mode == "gpu") ? gpu_tests(P) : cpu_tests(P);
        if (rc == 0) std::printf("%s tests passed\n", mode.c_str());
        return rc;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "unexpected exception: %s\n", e.what());
        return 2;
    }
}
