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
    std::vector<double> A((size_t)lda * K), B((size_t)ldb * K), C((size_t)ldc * n), C0;
    std::vector<double> rs(m), cs(n);
    for (auto& v : rs) v = std::exp(2.0 * uni());
    for (auto& v : cs) v = std::exp(2.0 * uni());
    for (long k = 0; k < K; ++k) {
        for (long i = 0; i < lda; ++i) A[i + k * lda] = i < m ? rs[i] * wide() : 1e300;
        for (long i = 0; i < ldb; ++i) B[i + k * ldb] = i < n ? cs[i] * uni() : 1e300;
    }
    if (m >= 256) {  // an all-zero row, a row with one huge entry, and a row with one NaN (must poison its row of C only)
        for (long k = 0; k < K; ++k) {
            A[5 + k * lda] = 0.0;
            if (k == 3) A[6 + k * lda] = 1e12;
        }
        A[7 + 9 * lda] = std::nan("");
    }
    for (auto& v : C) v = uni();
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
        if (s == 10 && m >= 256) { i = 7; }
        long double ref = 0.0L;
        for (long k = 0; k < K; ++k) ref += (long double)A[i + k * lda] * (long double)(sgn[k] * B[j + k * ldb]);
        const double c0 = C0[i + j * ldc], got = C[i + j * ldc];
        const bool skipped = c.lower && (i / 128) * 128 + 127 < (j / 64) * 64;
        if (skipped) {
            worst_untouched = std::max(worst_untouched, std::fabs(got - c0));
            continue;
        }
        if (m >= 256 && i == 7) {   // the NaN row
            if (!std::isnan(got)) { worst = 1.0; bad_i = i; bad_j = j; }
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
            const OzLayout L(c.S, (size_t)m, (size_t)n, (size_t)K);
            a.PA = (int8_t*)(w + L.pa);
            a.PB = (int8_t*)(w + L.pb);
            a.sa = (double*)(w + L.sa);
            a.sb = (double*)(w + L.sb);
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

// Column-list mode as the multi-GPU trailing update uses it: C and B are addressed from absolute tile column 0, the grid
// covers the tile columns in `list`; the update is issued as head + rest with the planes of A reused by the second call.
static double run_list_case(int S, int* d_err) {
    const int NTa = 12, e = 3, K = 512;
    const int m_tiles = NTa - e + 1;                 // rows from tile e down, plus one tile of right-hand-side rows
    const long m = (long)m_tiles * 128, nabs = (long)NTa * 128;
    const long lda = m, ldb = nabs + 128, ldc = m + 128;
    std::vector<int> list = {4, 6, 7, 10};
    unsigned long long sm = 77 + S;
    auto uni = [&]() {
        unsigned long long z = (sm += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        return (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
    };
    std::vector<double> A((size_t)lda * K), B((size_t)ldb * K), C((size_t)ldc * nabs), C0;
    for (auto& v : A) v = uni();
    for (auto& v : B) v = 3.0 * uni();
    for (auto& v : C) v = uni();
    C0 = C;
    std::vector<int> neg_cnt(K / 128, 0), neg_idx(K, 0);
    std::vector<double> sgn(K, 1.0);
    for (int k = 0; k < K; ++k)
        if ((k * 7) % 5 == 0) {
            sgn[k] = -1.0;
            neg_idx[(k / 128) * 128 + neg_cnt[k / 128]++] = k % 128;
        }
    double *dA, *dB, *dC;
    int *dcnt, *didx, *dlist;
    void* ws;
    const size_t wsb = ozaki_workspace_bytes(S, m_tiles, (int)list.size(), K);
    CK(cudaMalloc(&dA, A.size() * 8));
    CK(cudaMalloc(&dB, B.size() * 8));
    CK(cudaMalloc(&dC, C.size() * 8));
    CK(cudaMalloc(&dcnt, neg_cnt.size() * 4));
    CK(cudaMalloc(&didx, neg_idx.size() * 4));
    CK(cudaMalloc(&dlist, list.size() * 4));
    CK(cudaMalloc(&ws, wsb));
    CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dC, C.data(), C.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dcnt, neg_cnt.data(), neg_cnt.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(didx, neg_idx.data(), neg_idx.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dlist, list.data(), list.size() * 4, cudaMemcpyHostToDevice));
    GemmArgs g{};
    g.A = dA;
    g.B = dB;
    g.C = dC;
    g.lda = lda;
    g.ldb = ldb;
    g.ldc = ldc;
    g.K = K;
    g.lower = 1;
    g.subtract = 1;
    g.neg_cnt = dcnt;
    g.neg_idx = didx;
    g.row0 = e * 128;
    g.n_list = dlist;
    CK(launch_gemm_ozaki(g, m_tiles, 1, S, ws, wsb, d_err, 0, false));          // head: the first listed column
    g.n_list = dlist + 1;
    CK(launch_gemm_ozaki(g, m_tiles, (int)list.size() - 1, S, ws, wsb, d_err, 0, true));   // rest, planes of A reused
    CK(cudaDeviceSynchronize());
    int err = 0;
    CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost));
    double worst = 0.0, untouched = 0.0;
    for (long j = 0; j < nabs; ++j) {
        const bool listed = std::find(list.begin(), list.end(), (int)(j / 128)) != list.end();
        for (long i = 0; i < m; i += 3) {
            const double c0 = C0[i + j * ldc], got = C[i + j * ldc];
            const bool skipped = !listed || (e * 128 + (i / 128) * 128 + 127 < (j / 64) * 64);
            if (skipped) {
                untouched = std::max(untouched, std::fabs(got - c0));
                continue;
            }
            long double ref = 0.0L;
            for (long k = 0; k < K; ++k) ref += (long double)A[i + k * lda] * (long double)(sgn[k] * B[j + k * ldb]);
            worst = std::max(worst, (double)fabsl((long double)got - ((long double)c0 - ref)) / (3.0 * K));
        }
    }
    const double bound = std::max(std::ldexp(1.0, -(8 * S - 6)), 4e-16);
    const bool ok = err == 0 && worst <= bound && untouched == 0.0;
    printf("{\"case\": \"column list {4, 6, 7, 10} of 12, rows from tile 3, K 512, S %d, head + rest with reused A planes\", "
           "\"pipeline_error\": %d, \"max_err_over_rowmax_colmax_K\": %.3e, \"untouched_diff\": %.1e, \"ok\": %s}\n",
           S, err, worst, untouched, ok ? "true" : "false");
    fflush(stdout);
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dcnt); cudaFree(didx); cudaFree(dlist); cudaFree(ws);
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
    fails += run_list_case(7, d_err);
    fails += run_list_case(5, d_err);
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
