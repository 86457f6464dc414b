// Stand-alone check + timing of the int8-sliced FP64 GEMM (cheq_b200/csrc/kernels_ozaki.cu) on a B200.
//
//   tools/ozaki_probe                 correctness (small and ragged shapes, every slice count) + timings
//   tools/ozaki_probe quick           correctness only
//
// Correctness: C -= A S B^T against a long-double CPU product on sampled entries, error measured against
// rowmax(A) * rowmax(B) * K (the quantity the slicing bounds) and compared with the slice count's guarantee.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/ozaki_probe tools/ozaki_probe.cu
#include "../cheq_b200/csrc/kernels_ozaki.cu"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <random>
#include <vector>

using namespace cheq;

#define CK(x)                                                                                     \
    do {                                                                                          \
        cudaError_t e__ = (x);                                                                    \
        if (e__ != cudaSuccess) {                                                                 \
            printf("{\"error\": \"%s: %s (line %d)\"}\n", #x, cudaGetErrorString(e__), __LINE__); \
            exit(1);                                                                              \
        }                                                                                         \
    } while (0)

struct Case {
    int m_tiles, n_tiles, K, S, lower, subtract, samples;
    bool time_it;
};

static double run_case(const Case& c, int* d_err) {
    const long m = (long)c.m_tiles * 128, n = (long)c.n_tiles * 128;
    const int K = c.K;
    const long lda = m + 128, ldb = n + 256, ldc = m + 384;
    std::mt19937_64 rng(12345 + c.K + c.m_tiles * 7 + c.S);
    // cheap generators (the big cases have 10^8 entries): uniform in (-1, 1), and one with a wide dynamic range
    unsigned long long sm = 0x9E3779B97F4A7C15ull * (unsigned long long)(c.K + 3 * c.S + c.m_tiles);
    auto next = [&]() {
        unsigned long long z = (sm += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    };
    auto uni = [&]() { return (double)(next() >> 11) * (2.0 / 9007199254740992.0) - 1.0; };
    auto wide = [&]() { const unsigned long long z = next(); return std::ldexp((double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0, -(int)(z & 7)); };
    struct { std::function<double()> f; double operator()(std::mt19937_64&) { return f(); } } nd{uni};
    std::vector<double> A((size_t)lda * K), B((size_t)ldb * K), C((size_t)ldc * n), C0;
    std::vector<double> rs(m), cs(n);
    for (auto& v : rs) v = std::exp(2.0 * nd(rng));
    for (auto& v : cs) v = std::exp(2.0 * nd(rng));
    for (long k = 0; k < K; ++k) {
        for (long i = 0; i < lda; ++i) A[i + k * lda] = i < m ? rs[i] * wide() : 1e300;
        for (long i = 0; i < ldb; ++i) B[i + k * ldb] = i < n ? cs[i] * uni() : 1e300;
    }
    if (m >= 256)   // an all-zero row and a row with one huge entry
        for (long k = 0; k < K; ++k) {
            A[5 + k * lda] = 0.0;
            if (k == 3) A[6 + k * lda] = 1e12;
        }
    for (auto& v : C) v = nd(rng);
    C0 = C;
    std::vector<int> neg_cnt(K / 128, 0), neg_idx(K, 0);
    std::vector<double> sgn(K, 1.0);
    for (int k = 0; k < K; ++k)
        if (rng() % 3 == 0) {
            sgn[k] = -1.0;
            neg_idx[(k / 128) * 128 + neg_cnt[k / 128]++] = k % 128;
        }
    double *dA, *dB, *dC;
    int *dcnt, *didx;
    void* ws;
    const size_t wsb = ozaki_workspace_bytes(c.S, c.m_tiles, c.n_tiles, K);
    CK(cudaMalloc(&dA, A.size() * 8));
    CK(cudaMalloc(&dB, B.size() * 8));
    CK(cudaMalloc(&dC, C.size() * 8));
    CK(cudaMalloc(&dcnt, neg_cnt.size() * 4));
    CK(cudaMalloc(&didx, neg_idx.size() * 4));
    CK(cudaMalloc(&ws, wsb));
    CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dC, C.data(), C.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dcnt, neg_cnt.data(), neg_cnt.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(didx, neg_idx.data(), neg_idx.size() * 4, cudaMemcpyHostToDevice));
    GemmArgs g{};
    g.A = dA;
    g.B = dB;
    g.C = dC;
    g.lda = lda;
    g.ldb = ldb;
    g.ldc = ldc;
    g.K = K;
    g.lower = c.lower;
    g.subtract = c.subtract;
    g.neg_cnt = dcnt;
    g.neg_idx = didx;
    CK(launch_gemm_ozaki(g, c.m_tiles, c.n_tiles, c.S, ws, wsb, d_err, 0));
    CK(cudaDeviceSynchronize());
    int err = 0;
    CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost));
    // sampled check
    std::vector<double> amax(m, 0.0), bmax(n, 0.0);
    for (long k = 0; k < K; ++k) {
        for (long i = 0; i < m; ++i) amax[i] = std::max(amax[i], std::fabs(A[i + k * lda]));
        for (long j = 0; j < n; ++j) bmax[j] = std::max(bmax[j], std::fabs(B[j + k * ldb]));
    }
    double worst = 0.0, worst_untouched = 0.0;
    long bad_i = -1, bad_j = -1;
    for (int s = 0; s < c.samples; ++s) {
        long i = (long)(rng() % (unsigned long)m), j = (long)(rng() % (unsigned long)n);
        if (s < 8) { i = (s & 1) ? m - 1 - s : s; j = (s & 2) ? n - 1 - s : s; }
        if (s == 8 && m >= 256) { i = 5; }
        if (s == 9 && m >= 256) { i = 6; }
        long double ref = 0.0L;
        for (long k = 0; k < K; ++k) ref += (long double)A[i + k * lda] * (long double)(sgn[k] * B[j + k * ldb]);
        const double c0 = C0[i + j * ldc], got = C[i + j * ldc];
        const bool skipped = c.lower && (i / 128) * 128 + 127 < (j / 64) * 64;
        if (skipped) {
            worst_untouched = std::max(worst_untouched, std::fabs(got - c0));
            continue;
        }
        const long double want = c.subtract ? (long double)c0 - ref : ref;
        const double scale = amax[i] * bmax[j] * K + std::fabs(c0) * 1e-3 + 1e-300;
        const double e = (double)fabsl((long double)got - want) / scale;
        if (e > worst) { worst = e; bad_i = i; bad_j = j; }
    }
    double ms = 0.0, ms_gemm = 0.0;
    if (c.time_it) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        const int reps = 3;
        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r) CK(launch_gemm_ozaki(g, c.m_tiles, c.n_tiles, c.S, ws, wsb, d_err, 0));
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float f = 0;
        CK(cudaEventElapsedTime(&f, e0, e1));
        ms = f / reps;
        // the MMA kernel alone: slices are already in the workspace from the last call
        OzGemmArgs a{};
        {
            char* w = (char*)ws;
            unsigned long long* mxa = (unsigned long long*)w;
            double* sa = (double*)(mxa + m + n);
            size_t off = (size_t)(2 * (m + n)) * 8;
            off = (off + 1023) / 1024 * 1024;
            a.PA = (int8_t*)(w + off);
            a.PB = a.PA + (size_t)c.S * m * K;
            a.sa = sa;
            a.sb = sa + m;
            a.C = dC;
            a.ldc = ldc;
            a.nkb = K / 32;
            a.m_tiles = (int)(m / 128);
            a.n_tiles = (int)(n / 64);
            a.lower = c.lower;
            a.subtract = c.subtract;
            a.error = d_err;
        }
        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r) {
            const unsigned grid = (unsigned)(a.m_tiles * a.n_tiles);
            switch (c.S) {
                case 4: oz_gemm_kernel<4><<<grid, 128, oz_stages(4) * 4 * 6144>>>(a); break;
                case 5: oz_gemm_kernel<5><<<grid, 128, oz_stages(5) * 5 * 6144>>>(a); break;
                case 6: oz_gemm_kernel<6><<<grid, 128, oz_stages(6) * 6 * 6144>>>(a); break;
                case 7: oz_gemm_kernel<7><<<grid, 128, oz_stages(7) * 7 * 6144>>>(a); break;
                case 8: oz_gemm_kernel<8><<<grid, 128, oz_stages(8) * 8 * 6144>>>(a); break;
            }
        }
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&f, e0, e1));
        ms_gemm = f / reps;
        CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
    }
    const double flops = c.lower ? (double)n * n * K + 2.0 * (double)(m - n) * n * K : 2.0 * m * n * (double)K;
    const double bound = std::ldexp(1.0, -(8 * c.S - 6));   // generous: the guarantee is ~2^-(8 S - 3) + fp64 rounding
    const bool ok = err == 0 && worst <= std::max(bound, 4e-16) && worst_untouched == 0.0;
    printf("{\"case\": \"m %ld n %ld K %d S %d lower %d subtract %d\", \"pipeline_error\": %d, \"max_err_over_rowmax_colmax_K\": %.3e, "
           "\"bound\": %.3e, \"worst_at\": [%ld, %ld], \"untouched_diff\": %.1e, \"ok\": %s",
           m, n, K, c.S, c.lower, c.subtract, err, worst, bound, bad_i, bad_j, worst_untouched, ok ? "true" : "false");
    if (c.time_it)
        printf(", \"ms_total\": %.3f, \"ms_mma_kernel\": %.3f, \"tflops_total\": %.1f, \"tflops_mma_kernel\": %.1f", ms, ms_gemm,
               flops / ms * 1e-9, flops / ms_gemm * 1e-9);
    printf("}\n");
    fflush(stdout);
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dC);
    cudaFree(dcnt);
    cudaFree(didx);
    cudaFree(ws);
    if (err) CK(cudaMemset(d_err, 0, 4));
    return ok ? 0.0 : 1.0;
}

int main(int argc, char** argv) {
    const bool quick = argc > 1 && !strcmp(argv[1], "quick");
    CK(configure_ozaki_kernels());
    int* d_err;
    CK(cudaMalloc(&d_err, 4));
    CK(cudaMemset(d_err, 0, 4));
    double fails = 0;
    // one tile, one stage round; more K than stages; several tiles
    fails += run_case({1, 1, 128, 7, 0, 1, 400, false}, d_err);
    fails += run_case({1, 1, 1024, 7, 0, 0, 400, false}, d_err);
    for (int S = 4; S <= 8; ++S) fails += run_case({3, 2, 640, S, 0, 1, 600, false}, d_err);
    fails += run_case({5, 3, 1152, 7, 1, 1, 1500, false}, d_err);
    fails += run_case({20, 20, 2048, 7, 1, 1, 1500, !quick}, d_err);
    if (!quick) {
        fails += run_case({64, 64, 4096, 7, 0, 1, 300, true}, d_err);
        fails += run_case({64, 64, 4096, 8, 0, 1, 300, true}, d_err);
        fails += run_case({64, 64, 4096, 4, 0, 1, 300, true}, d_err);
        fails += run_case({105, 105, 6784, 7, 1, 1, 300, true}, d_err);   // the S20k Schur update
        fails += run_case({105, 24, 1024, 7, 0, 1, 300, true}, d_err);    // a mid-recursion update
    }
    printf("{\"failed_cases\": %d}\n", (int)fails);
    return fails > 0 ? 2 : 0;
}
