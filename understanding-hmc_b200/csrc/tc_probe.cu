// Stage-1 probe for the tensor-core gradient path (not part of the library): one CTA computes
//   C[128 x N] = A[128 x K] * B[N x K]^T   (bf16 inputs, fp32 accumulation in TMEM)
// with hand-written tcgen05.mma from shared-memory operands in the canonical no-swizzle K-major layout, reads the
// accumulator back with tcgen05.ld (thread t <-> row t) and compares with the host.
//   build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tc_probe.cu -o ../bin/tc_probe
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int M = 128, N = 112, K = 112;          // K, N multiples of 16
constexpr int KC = K / 8;                         // 16-byte chunks (8 bf16) along K

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, no swizzle, K-major: canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;                      // version = 1 (Blackwell)
    return d;                                    // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

__global__ void __launch_bounds__(128) tc_probe_kernel(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B, float* __restrict__ C, int ts) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint4* As = reinterpret_cast<uint4*>(smem);                 // [KC][M] 16-byte chunks
    uint4* Bs = As + KC * M;                                    // [KC][N]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(Bs + KC * N);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;

    // operands -> canonical layout: chunk (kc, row) at index kc * rows + row
    for (int t = tid; t < KC * M; t += blockDim.x) { const int kc = t / M, r = t % M; As[t] = *reinterpret_cast<const uint4*>(A + (size_t)r * K + kc * 8); }
    for (int t = tid; t < KC * N; t += blockDim.x) { const int kc = t / N, r = t % N; Bs[t] = *reinterpret_cast<const uint4*>(B + (size_t)r * K + kc * 8); }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");            // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t ACOL = 128;                              // A operand in TMEM (ts mode): columns ACOL .. ACOL + K/2
    if (ts) {
        // thread t <-> row t: K bf16 values packed two per 32-bit column (low half = even k), 8 columns per tcgen05.st
        for (int c0 = 0; c0 < K / 2; c0 += 8) {
            uint32_t w[8];
            for (int c = 0; c < 8; ++c) w[c] = *reinterpret_cast<const uint32_t*>(A + (size_t)tid * K + 2 * (c0 + c));
            const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + ACOL + (uint32_t)c0;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                         ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;");
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
    }

    if (tid == 0) {
        // instruction descriptor: D = F32, A = B = BF16, both K-major, N, M
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t a0 = smem_u32(As), b0 = smem_u32(Bs);
        for (int ks = 0; ks < K / 16; ++ks) {                  // one MMA = 16 bf16 along K = 2 chunks
            const uint64_t da = make_desc(a0 + ks * 2 * M * 16, M * 16, 8 * 16);
            const uint64_t db = make_desc(b0 + ks * 2 * N * 16, N * 16, 8 * 16);
            const uint32_t acc = ks > 0 ? 1u : 0u;
            if (ts) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                    ::"r"(tmem), "r"(tmem + ACOL + 8u * ks), "l"(db), "r"(idesc), "r"(acc), "r"(0u));
                continue;
            }
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
    }
    // everybody waits for the accumulator
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(mbar)), "r"(0u) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    // thread t reads row t: lanes 32*warp .. 32*warp+31, 16 columns at a time
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int c = 0; c < 16; ++c) C[(size_t)tid * N + c0 + c] = __uint_as_float(v[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

int main() {
    std::vector<uint16_t> hA(M * K), hB(N * K);
    std::vector<float> fA(M * K), fB(N * K), hC(M * N), ref(M * N);
    auto bf = [](float x) { __nv_bfloat16 b = __float2bfloat16(x); uint16_t u; memcpy(&u, &b, 2); return u; };
    auto fb = [](uint16_t u) { uint32_t w = (uint32_t)u << 16; float f; memcpy(&f, &w, 4); return f; };
    srand(1);
    for (int i = 0; i < M * K; ++i) { hA[i] = bf((rand() % 2001 - 1000) * 1e-3f); fA[i] = fb(hA[i]); }
    for (int i = 0; i < N * K; ++i) { hB[i] = bf((rand() % 2001 - 1000) * 1e-3f); fB[i] = fb(hB[i]); }
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)fA[m * K + k] * fB[n * K + k]; ref[m * N + n] = (float)s; }
    uint16_t *dA, *dB; float* dC;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dC, hC.size() * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0, hC.size() * 4));
    const size_t smem = (size_t)KC * (M + N) * 16 + 64;
    CK(cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int ts = 0; ts < 2; ++ts) {
    CK(cudaMemset(dC, 0, hC.size() * 4));
    tc_probe_kernel<<<1, 128, smem>>>(dA, dB, dC, ts);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int i = 0; i < M * N; ++i) { maxerr = fmax(maxerr, fabs((double)hC[i] - ref[i])); maxref = fmax(maxref, fabs((double)ref[i])); }
    printf(ts ? "A operand from TMEM (tcgen05.st): " : "A operand from shared memory: ");
    printf("tcgen05 probe: max |C - ref| = %.3e (max |ref| = %.3f)  C[0][0..3] = %f %f %f %f  ref = %f %f %f %f\n", maxerr, maxref,
           hC[0], hC[1], hC[2], hC[3], ref[0], ref[1], ref[2], ref[3]);
    printf(maxerr < 1e-3 * maxref ? "PASS\n" : "FAIL\n");
    }
    return 0;
}
