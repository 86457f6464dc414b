// tcgen05 feasibility probe for an Ozaki-style FP64 GEMM on B200.
//
// FP64 has no tcgen05.mma kind.  The only way to put the dense QEq factorisation on the 5th-generation tensor cores is
// to split every FP64 operand into int8 slices (error-free transformation), run the slice products on
// tcgen05.mma.kind::i8 with int32 accumulators in TMEM, and recombine in fp64: about s (s + 1) / 2 = 36 slice products
// for s = 8 slices per operand.  Whether that can beat DMMA (28.6 TF/s in situ, 35.4 cuBLAS) is decided by the int8
// rate the tensor cores actually sustain.  This program measures it: one CTA per SM, operands resident in shared
// memory (canonical no-swizzle K-major layout, written by the threads themselves -- no TMA, no HBM traffic), a single
// thread issues M128 x N256 x K32 int8 MMAs back to back into two TMEM accumulators, completion through
// tcgen05.commit on an mbarrier; the first K = 128 product is read back with tcgen05.ld and checked against the CPU.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tc05_probe tools/tc05_probe.cu
// Run:    tools/tc05_probe            (prints one JSON line)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int kM = 128, kN = 256, kKStep = 32;      // one tcgen05.mma.kind::i8 instruction
constexpr int kBK = 128;                            // bytes of K resident in shared memory (4 instructions)
constexpr int kSboA = (kBK / 16) * 128;             // byte stride between 8-row groups (core matrices along M / N)
constexpr int kLbo = 128;                           // byte stride between the two 16-byte K chunks of a core-matrix pair

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, SWIZZLE_NONE, K-major (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);                 // start address, bits [0, 14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;      // leading byte offset, bits [16, 30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;      // stride byte offset, bits [32, 46)
    d |= (uint64_t)1 << 46;                                // version = 1 (Blackwell), bits [46, 48)
    return d;                                              // base offset 0, lbo mode 0, layout type 0 (no swizzle)
}

// instruction descriptor for kind::i8: signed int8 x signed int8 -> int32, K-major A and B (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (2u << 4)            // c_format: S32
         | (1u << 7)            // a_format: signed 8 bit
         | (1u << 10)           // b_format: signed 8 bit
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u));
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity));
    return ok != 0;
}

__global__ void __launch_bounds__(128, 1)
tc05_probe_kernel(int outer, int inner, int32_t* check_out, unsigned long long* cycles_out, int* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    int8_t* sA = (int8_t*)smem;                       // 128 rows x 128 K-bytes = 16 KB
    int8_t* sB = (int8_t*)smem + kM * kBK;            // 256 rows x 128 K-bytes = 32 KB
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base_s;
    const int t = threadIdx.x, warp = t >> 5;

    // operands in the canonical layout: addr(r, k) = (r / 8) * SBO + (k / 16) * LBO + (r % 8) * 16 + (k % 16)
    for (int idx = t; idx < kM * kBK; idx += 128) {
        const int r = idx / kBK, k = idx % kBK;
        sA[(r / 8) * kSboA + (k / 16) * kLbo + (r % 8) * 16 + (k % 16)] = (int8_t)(((r * 7 + k * 3) % 11) - 5);
    }
    for (int idx = t; idx < kN * kBK; idx += 128) {
        const int r = idx / kBK, k = idx % kBK;
        sB[(r / 8) * kSboA + (k / 16) * kLbo + (r % 8) * 16 + (k % 16)] = (int8_t)(((r * 5 + k) % 7) - 3);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;\n");
    }
    // generic-proxy writes of the operands must be visible to the tensor-core (async) proxy
    asm volatile("fence.proxy.async.shared::cta;\n");
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = make_idesc(kM, kN);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB), barp = smem_u32(&bar);
    uint32_t parity = 0;
    int ok = 1;

    auto commit_and_wait = [&]() {
        if (t == 0 && ok) {
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(barp));
            const long long t0 = clock64();
            while (!mbar_try_wait(barp, parity)) {
                if (clock64() - t0 > 4000000000ll) {   // ~2 s: never hang the GPU
                    ok = 0;
                    break;
                }
            }
        }
        parity ^= 1u;
    };

    // ---- phase 1: one K = 128 product (4 instructions) into accumulator 0, read back and checked on the host ----
    if (t == 0) {
        for (int ks = 0; ks < kBK / kKStep; ++ks) {
            // K advances by 32 bytes = two 16-byte chunks = 2 * LBO
            const uint64_t da = make_desc(a0 + ks * 2 * kLbo, kLbo, kSboA);
            const uint64_t db = make_desc(b0 + ks * 2 * kLbo, kLbo, kSboA);
            mma_i8(tmem, da, db, idesc, ks > 0 ? 1u : 0u);
        }
    }
    commit_and_wait();
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    {
        // warp w reads TMEM lanes 32 w .. 32 w + 31 (accumulator rows), columns 0..7
        uint32_t v[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n");
        if (blockIdx.x == 0)
            for (int c = 0; c < 8; ++c) check_out[t * 8 + c] = (int32_t)v[c];
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();

    // ---- phase 2: sustained issue rate: `outer` batches of `inner` instructions, alternating two accumulators ----
    const long long c0 = clock64();
    for (int o = 0; o < outer; ++o) {
        if (t == 0 && ok) {
            for (int i = 0; i < inner; ++i) {
                const int ks = i & 3;
                const uint64_t da = make_desc(a0 + ks * 2 * kLbo, kLbo, kSboA);
                const uint64_t db = make_desc(b0 + ks * 2 * kLbo, kLbo, kSboA);
                mma_i8(tmem + (uint32_t)((i >> 2) & 1) * kN, da, db, idesc, 1u);
            }
        }
        commit_and_wait();
    }
    const long long c1 = clock64();
    __syncthreads();
    if (t == 0) {
        cycles_out[blockIdx.x] = (unsigned long long)(c1 - c0);
        if (!ok) atomicExch(status, 1);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512));
}

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e__ = (x);                                                         \
        if (e__ != cudaSuccess) {                                                      \
            printf("{\"error\": \"%s: %s\"}\n", #x, cudaGetErrorString(e__));          \
            return 1;                                                                  \
        }                                                                              \
    } while (0)

int main(int argc, char** argv) {
    const int outer = argc > 1 ? atoi(argv[1]) : 400, inner = argc > 2 ? atoi(argv[2]) : 64;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int32_t* d_check;
    unsigned long long* d_cycles;
    int* d_status;
    CK(cudaMalloc(&d_check, kM * 8 * sizeof(int32_t)));
    CK(cudaMalloc(&d_cycles, sms * sizeof(unsigned long long)));
    CK(cudaMalloc(&d_status, sizeof(int)));
    CK(cudaMemset(d_status, 0, sizeof(int)));
    const int smem = kM * kBK + kN * kBK + 1024;
    CK(cudaFuncSetAttribute(tc05_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    tc05_probe_kernel<<<sms, 128, smem>>>(8, inner, d_check, d_cycles, d_status);   // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    tc05_probe_kernel<<<sms, 128, smem>>>(outer, inner, d_check, d_cycles, d_status);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<int32_t> chk(kM * 8);
    std::vector<unsigned long long> cyc(sms);
    int status = 0;
    CK(cudaMemcpy(chk.data(), d_check, chk.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cyc.data(), d_cycles, cyc.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&status, d_status, 4, cudaMemcpyDeviceToHost));
    // CPU check of D[m][n] = sum_k A[m][k] B[n][k], k < 128, for n < 8
    int bad = 0;
    for (int m = 0; m < kM; ++m)
        for (int n = 0; n < 8; ++n) {
            int ref = 0;
            for (int k = 0; k < kBK; ++k) ref += (((m * 7 + k * 3) % 11) - 5) * (((n * 5 + k) % 7) - 3);
            if (ref != chk[m * 8 + n]) ++bad;
        }
    const double ops = 2.0 * kM * kN * kKStep * (double)outer * inner * sms;
    unsigned long long cmax = 0;
    for (auto c : cyc) cmax = c > cmax ? c : cmax;
    const double pops = ops / (ms * 1e-3) / 1e15;
    printf("{\"probe\": \"tcgen05.mma.cta_group::1.kind::i8 M128 N256 K32, operands resident in shared memory\", "
           "\"sms\": %d, \"instructions_per_sm\": %lld, \"ms\": %.4f, \"int8_pops\": %.4f, "
           "\"cycles_per_instruction\": %.2f, \"check_mismatches\": %d, \"timeout\": %d, "
           "\"fp64_equivalent_tflops_if_36_slice_products\": %.2f, \"fp64_equivalent_tflops_if_21_slice_products\": %.2f}\n",
           sms, (long long)outer * inner, ms, pops, (double)cmax / ((double)outer * inner), bad, status,
           pops * 1e3 / 36.0, pops * 1e3 / 21.0);
    return 0;
}
