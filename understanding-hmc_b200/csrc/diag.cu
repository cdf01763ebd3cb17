// Diagnostics as HBM-bound reductions over the sample stream q_chain[Nchain][L_chain][D]:
//   hmc_diag_moments   -- split-chain mean / ddof=1 std  (utils.convergence_stats, /root/reference/utils.py:88-118)
//   hmc_diag_variogram -- variogram numerators for a chunk of lags (utils.variogram, utils.py:161-179)
// One thread owns one (split chain, dimension) time series at a time; consecutive threads own consecutive
// dimensions, so every load of a warp is one contiguous row segment.  Per-thread partials are float64,
// combined per block in shared memory and added to the output with one float64 atomic per (block, value).
#include "hmc_common.cuh"

namespace {

constexpr int kDiagThreads = 256;

// `q` points at sample 0 of chain 0; `stride_chain` elements between chains; a split chain s = 2*m + h is
// samples [h*n, h*n + n) of chain m (utils.py:102-104).
template <typename T>
__global__ void __launch_bounds__(kDiagThreads) diag_moments_kernel(const T* __restrict__ q, long Nchain, long n, int D,
                                                                    long stride_chain, int spb, double* __restrict__ out) {
    extern __shared__ double sm[];   // [3][D]
    for (int t = threadIdx.x; t < 3 * D; t += blockDim.x) sm[t] = 0.0;
    __syncthreads();
    const int d = threadIdx.x % D;
    const int sl = threadIdx.x / D;
    double s_std = 0.0, s_mean = 0.0, s_mean2 = 0.0;
    if (sl < spb) {
        const long nseries = 2 * Nchain;
        for (long s = (long)blockIdx.x * spb + sl; s < nseries; s += (long)gridDim.x * spb) {
            const T* x = q + (s >> 1) * stride_chain + (s & 1) * n * D + d;
            const double x0 = (double)x[0];
            double a = 0.0, b = 0.0;
            long i = 0;
            for (; i + 4 <= n; i += 4) {
                const double v0 = (double)x[(i + 0) * D] - x0, v1 = (double)x[(i + 1) * D] - x0;
                const double v2 = (double)x[(i + 2) * D] - x0, v3 = (double)x[(i + 3) * D] - x0;
                a += (v0 + v1) + (v2 + v3);
                b += (v0 * v0 + v1 * v1) + (v2 * v2 + v3 * v3);
            }
            for (; i < n; ++i) { const double v = (double)x[i * D] - x0; a += v; b += v * v; }
            const double mean_s = a / (double)n;
            double var = (b - (double)n * mean_s * mean_s) / (double)(n - 1);   // ddof = 1 (utils.py:111)
            if (var < 0.0) var = 0.0;
            const double mean = mean_s + x0;
            s_std += sqrt(var);
            s_mean += mean;
            s_mean2 += mean * mean;
        }
        atomicAdd(&sm[d], s_std);
        atomicAdd(&sm[D + d], s_mean);
        atomicAdd(&sm[2 * D + d], s_mean2);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 3 * D; t += blockDim.x) atomicAdd(out + t, sm[t]);
}


// float32 stream, D % 4 == 0: one thread owns FOUR adjacent dimensions of one split chain and walks the time axis
// with 128-bit loads (8 rows in flight per thread), so that enough bytes are in flight to saturate HBM.
__global__ void __launch_bounds__(kDiagThreads) diag_moments_f32x4_kernel(const float* __restrict__ q, long Nchain, long n, int D,
                                                                          long stride_chain, int spb, double* __restrict__ out) {
    extern __shared__ double sm[];   // [3][D]
    for (int t = threadIdx.x; t < 3 * D; t += blockDim.x) sm[t] = 0.0;
    __syncthreads();
    const int D4 = D >> 2;
    const int dq = threadIdx.x % D4;
    const int sl = threadIdx.x / D4;
    double s_std[4] = {0, 0, 0, 0}, s_mean[4] = {0, 0, 0, 0}, s_mean2[4] = {0, 0, 0, 0};
    if (sl < spb) {
        const long nseries = 2 * Nchain;
        for (long s = (long)blockIdx.x * spb + sl; s < nseries; s += (long)gridDim.x * spb) {
            const float4* x = reinterpret_cast<const float4*>(q + (s >> 1) * stride_chain + (s & 1) * n * D) + dq;
            const float4 x0 = x[0];
            // float partial sums over blocks of 8 rows (shifted by the first sample), float64 across blocks
            double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
            long i = 0;
            for (; i + 8 <= n; i += 8) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = x[(i + u) * D4];
                float pa[4] = {0.f, 0.f, 0.f, 0.f}, pb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float e0 = v[u].x - x0.x, e1 = v[u].y - x0.y, e2 = v[u].z - x0.z, e3 = v[u].w - x0.w;
                    pa[0] += e0; pa[1] += e1; pa[2] += e2; pa[3] += e3;
                    pb[0] = fmaf(e0, e0, pb[0]); pb[1] = fmaf(e1, e1, pb[1]); pb[2] = fmaf(e2, e2, pb[2]); pb[3] = fmaf(e3, e3, pb[3]);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) { a[c] += (double)pa[c]; b[c] += (double)pb[c]; }
            }
            for (; i < n; ++i) {
                const float4 v = x[i * D4];
                const double e[4] = {(double)v.x - x0.x, (double)v.y - x0.y, (double)v.z - x0.z, (double)v.w - x0.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) { a[c] += e[c]; b[c] += e[c] * e[c]; }
            }
            const double xs[4] = {(double)x0.x, (double)x0.y, (double)x0.z, (double)x0.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const double mean_s = a[c] / (double)n;
                double var = (b[c] - (double)n * mean_s * mean_s) / (double)(n - 1);   // ddof = 1 (utils.py:111)
                if (var < 0.0) var = 0.0;
                const double mean = mean_s + xs[c];
                s_std[c] += sqrt(var); s_mean[c] += mean; s_mean2[c] += mean * mean;
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            atomicAdd(&sm[4 * dq + c], s_std[c]);
            atomicAdd(&sm[D + 4 * dq + c], s_mean[c]);
            atomicAdd(&sm[2 * D + 4 * dq + c], s_mean2[c]);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 3 * D; t += blockDim.x) atomicAdd(out + t, sm[t]);
}

constexpr int kShortThreads = 128;   // small blocks: several resident per SM, so loads and arithmetic of different blocks overlap

// Short series (n <= NMAX, all lags in one launch, lag0 = 1): every thread reads the n values of its (series, dimension)
// ONCE, back to back (n independent coalesced loads in flight), and forms the moments (optional) and all lag sums from
// registers: one HBM pass over the stored samples for the whole of utils.convergence_stats.  The loop over i branches
// on the (warp-uniform) series length, so exactly n (n-1) / 2 difference terms are evaluated; per lag the terms are
// added in increasing i, float partial sums as in the windowed kernel below.
template <typename T, int NMAX, bool MOMENTS>
__global__ void __launch_bounds__(kShortThreads, 4) diag_short_kernel(const T* __restrict__ q, long Nchain, long n, int D,
                                                                  long stride_chain, int spb, int nlags,
                                                                  double* __restrict__ mom_out, double* __restrict__ out) {
    extern __shared__ double sm[];   // [NMAX + 3][D]
    for (int t = threadIdx.x; t < (NMAX + 3) * D; t += blockDim.x) sm[t] = 0.0;
    __syncthreads();
    const int d = threadIdx.x % D;
    const int sl = threadIdx.x / D;
    // lag t = k + 1.  The partial sums stay in the stream's own precision over the ~150 series a thread visits (a few
    // thousand terms of like magnitude: 4e-6 relative in float32, below the float32 input's own rounding in n_eff) and
    // are widened once; float64 accumulators here cost 62 registers and a third of the resident warps.
    T acc[NMAX - 1];
#pragma unroll
    for (int k = 0; k < NMAX - 1; ++k) acc[k] = T(0);
    double s_std = 0.0, s_mean = 0.0, s_mean2 = 0.0;
    if (sl < spb) {
        const long nseries = 2 * Nchain;
        for (long s = (long)blockIdx.x * spb + sl; s < nseries; s += (long)gridDim.x * spb) {
            const T* x = q + (s >> 1) * stride_chain + (s & 1) * n * D + d;
            T v[NMAX];
#pragma unroll
            for (int i = 0; i < NMAX; ++i) v[i] = (i < n) ? x[(long)i * D] : T(0);
#pragma unroll
            for (int i = 1; i < NMAX; ++i) {
                if (i < n) {
#pragma unroll
                    for (int k = 0; k < i; ++k) { const T df = v[i] - v[i - k - 1]; acc[k] = fma(df, df, acc[k]); }
                }
            }
            if (MOMENTS) {           // per split chain mean and ddof = 1 standard deviation (utils.py:107-118)
                const double x0 = (double)v[0];
                double a = 0.0, b = 0.0;
#pragma unroll
                for (int i = 1; i < NMAX; ++i) {
                    if (i < n) { const double e = (double)v[i] - x0; a += e; b += e * e; }
                }
                const double mean_s = a / (double)n;
                double var = (b - (double)n * mean_s * mean_s) / (double)(n - 1);
                if (var < 0.0) var = 0.0;
                const double mean = mean_s + x0;
                s_std += sqrt(var); s_mean += mean; s_mean2 += mean * mean;
            }
        }
#pragma unroll
        for (int k = 0; k < NMAX - 1; ++k) if (k < nlags) atomicAdd(&sm[k * D + d], (double)acc[k]);
        if (MOMENTS) {
            atomicAdd(&sm[(NMAX + 0) * D + d], s_std);
            atomicAdd(&sm[(NMAX + 1) * D + d], s_mean);
            atomicAdd(&sm[(NMAX + 2) * D + d], s_mean2);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nlags * D; t += blockDim.x) atomicAdd(out + t, sm[t]);
    if (MOMENTS) for (int t = threadIdx.x; t < 3 * D; t += blockDim.x) atomicAdd(mom_out + t, sm[NMAX * D + t]);
}

// Lags t = lag0 + k, k < NL.  For every i the pair (x[i], x[i - lag0 - k]) contributes (x[i]-x[i-lag0-k])^2.
// The last NL delayed values live in a register window addressed with compile-time indices (the time loop is
// unrolled by NL); float partial sums are flushed into float64 every NL steps.
template <typename T, int NL>
__global__ void __launch_bounds__(kDiagThreads) diag_variogram_kernel(const T* __restrict__ q, long Nchain, long n, int D,
                                                                      long stride_chain, int spb, int lag0, int nlags,
                                                                      double* __restrict__ out) {
    extern __shared__ double sm[];   // [NL][D]
    for (int t = threadIdx.x; t < NL * D; t += blockDim.x) sm[t] = 0.0;
    __syncthreads();
    const int d = threadIdx.x % D;
    const int sl = threadIdx.x / D;
    double dacc[NL];
#pragma unroll
    for (int k = 0; k < NL; ++k) dacc[k] = 0.0;
    if (sl < spb) {
        const long nseries = 2 * Nchain;
        for (long s = (long)blockIdx.x * spb + sl; s < nseries; s += (long)gridDim.x * spb) {
            const T* x = q + (s >> 1) * stride_chain + (s & 1) * n * D + d;
            T w[NL];         // w[(i - lag0) % NL] = x[i - lag0]
            T acc[NL];
#pragma unroll
            for (int k = 0; k < NL; ++k) { w[k] = T(0); acc[k] = T(0); }
            long i0 = lag0;
            // first block: entry k is valid only once i - lag0 - k >= 0, i.e. k <= r (compile-time)
            {
#pragma unroll
                for (int r = 0; r < NL; ++r) {
                    const long i = i0 + r;
                    if (i < n) {
                        const T xa = x[i * D];
                        w[r] = x[(i - lag0) * D];
#pragma unroll
                        for (int k = 0; k < NL; ++k) {
                            if (k <= r) { const T df = xa - w[(r - k + NL) % NL]; acc[k] = fma(df, df, acc[k]); }
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < NL; ++k) { dacc[k] += (double)acc[k]; acc[k] = T(0); }
                i0 += NL;
            }
            for (; i0 + NL <= n; i0 += NL) {        // steady state, no conditions
                // both streams of the block are requested up front: 2 NL independent loads in flight per thread (the kernel
                // is bound by load latency, not by HBM bandwidth: one resident block per SM at this register count)
                T xa[NL], wn[NL];
#pragma unroll
                for (int r = 0; r < NL; ++r) { xa[r] = x[(i0 + r) * D]; wn[r] = x[(i0 + r - lag0) * D]; }
#pragma unroll
                for (int r = 0; r < NL; ++r) {
                    w[r] = wn[r];
#pragma unroll
                    for (int k = 0; k < NL; ++k) { const T df = xa[r] - w[(r - k + NL) % NL]; acc[k] = fma(df, df, acc[k]); }
                }
#pragma unroll
                for (int k = 0; k < NL; ++k) { dacc[k] += (double)acc[k]; acc[k] = T(0); }
            }
            if (i0 < n) {                            // tail
#pragma unroll
                for (int r = 0; r < NL; ++r) {
                    const long i = i0 + r;
                    if (i < n) {
                        const T xa = x[i * D];
                        w[r] = x[(i - lag0) * D];
#pragma unroll
                        for (int k = 0; k < NL; ++k) { const T df = xa - w[(r - k + NL) % NL]; acc[k] = fma(df, df, acc[k]); }
                    }
                }
#pragma unroll
                for (int k = 0; k < NL; ++k) { dacc[k] += (double)acc[k]; acc[k] = T(0); }
            }
        }
#pragma unroll
        for (int k = 0; k < NL; ++k) if (k < nlags) atomicAdd(&sm[k * D + d], dacc[k]);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nlags * D; t += blockDim.x) atomicAdd(out + t, sm[t]);
}

int grid_for(long nseries, int spb, int per_sm = 8) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long want = (nseries + spb - 1) / spb;
    long cap = (long)sms * per_sm;
    return (int)(want < cap ? want : cap);
}

}  // namespace

#define HMC_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            hmc_set_error(__VA_ARGS__);   \
            return HMC_E_BADARG;          \
        }                                 \
    } while (0)

extern "C" int hmc_diag_moments(int32_t dtype, const void* q, int64_t Nchain, int64_t n, int32_t D, int64_t stride_chain,
                                double* out3xD, void* cuda_stream) {
    HMC_REQUIRE(q && out3xD, "NULL buffer");
    HMC_REQUIRE(Nchain >= 1 && n >= 2 && D >= 1 && D <= kDiagThreads, "need Nchain >= 1, n >= 2, 1 <= D <= %d", kDiagThreads);
    HMC_REQUIRE(stride_chain >= 2 * n * D, "stride_chain too small");
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    const int spb = kDiagThreads / D;
    const int grid = grid_for(2 * Nchain, spb);
    const size_t smem = sizeof(double) * 3 * D;
    HMC_CUDA_CHECK(cudaMemsetAsync(out3xD, 0, smem, stream));
    const bool vec4 = dtype == HMC_F32 && (D % 4) == 0 && (stride_chain % 4) == 0 && (n * D) % 4 == 0 &&
                      (reinterpret_cast<uintptr_t>(q) % 16) == 0;
    if (vec4) {
        const int spb4 = kDiagThreads / (D / 4);
        diag_moments_f32x4_kernel<<<grid_for(2 * Nchain, spb4), kDiagThreads, smem, stream>>>((const float*)q, Nchain, n, D, stride_chain, spb4, out3xD);
    } else if (dtype == HMC_F32)
        diag_moments_kernel<float><<<grid, kDiagThreads, smem, stream>>>((const float*)q, Nchain, n, D, stride_chain, spb, out3xD);
    else
        diag_moments_kernel<double><<<grid, kDiagThreads, smem, stream>>>((const double*)q, Nchain, n, D, stride_chain, spb, out3xD);
    HMC_CUDA_CHECK(cudaGetLastError());
    return HMC_OK;
}

extern "C" int hmc_diag_variogram(int32_t dtype, const void* q, int64_t Nchain, int64_t n, int32_t D, int64_t stride_chain,
                                  int32_t lag0, int32_t nlags, double* out, void* cuda_stream) {
    constexpr int NL = 32;
    HMC_REQUIRE(q && out, "NULL buffer");
    HMC_REQUIRE(Nchain >= 1 && n >= 2 && D >= 1 && D <= kDiagThreads, "need Nchain >= 1, n >= 2, 1 <= D <= %d", kDiagThreads);
    HMC_REQUIRE(lag0 >= 1 && nlags >= 1 && nlags <= NL, "need lag0 >= 1 and 1 <= nlags <= %d", NL);
    HMC_REQUIRE(stride_chain >= 2 * n * D, "stride_chain too small");
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    const int spb = kDiagThreads / D;
    const int grid = grid_for(2 * Nchain, spb);
    const size_t smem = sizeof(double) * NL * D;
    HMC_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(double) * nlags * D, stream));
    if (lag0 == 1 && n <= NL && D <= kShortThreads) {    // short series: one pass, values held in registers
        const size_t smem2 = sizeof(double) * (NL + 3) * D;
        if (dtype == HMC_F32) {
            HMC_CUDA_CHECK(cudaFuncSetAttribute(diag_short_kernel<float, NL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            diag_short_kernel<float, NL, false><<<grid_for(2 * Nchain, kShortThreads / D, 8), kShortThreads, smem2, stream>>>((const float*)q, Nchain, n, D, stride_chain, kShortThreads / D, nlags, nullptr, out);
        } else {
            HMC_CUDA_CHECK(cudaFuncSetAttribute(diag_short_kernel<double, NL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            diag_short_kernel<double, NL, false><<<grid_for(2 * Nchain, kShortThreads / D, 8), kShortThreads, smem2, stream>>>((const double*)q, Nchain, n, D, stride_chain, kShortThreads / D, nlags, nullptr, out);
        }
        HMC_CUDA_CHECK(cudaGetLastError());
        return HMC_OK;
    }
    if (dtype == HMC_F32) {
        HMC_CUDA_CHECK(cudaFuncSetAttribute(diag_variogram_kernel<float, NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        diag_variogram_kernel<float, NL><<<grid, kDiagThreads, smem, stream>>>((const float*)q, Nchain, n, D, stride_chain, spb, lag0, nlags, out);
    } else {
        HMC_CUDA_CHECK(cudaFuncSetAttribute(diag_variogram_kernel<double, NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        diag_variogram_kernel<double, NL><<<grid, kDiagThreads, smem, stream>>>((const double*)q, Nchain, n, D, stride_chain, spb, lag0, nlags, out);
    }
    HMC_CUDA_CHECK(cudaGetLastError());
    return HMC_OK;
}

extern "C" int hmc_diag_short_series(int32_t dtype, const void* q, int64_t Nchain, int64_t n, int32_t D, int64_t stride_chain,
                                     int32_t nlags, double* out3xD, double* out_lags, void* cuda_stream) {
    constexpr int NL = 32;
    HMC_REQUIRE(q && out3xD && out_lags, "NULL buffer");
    HMC_REQUIRE(Nchain >= 1 && n >= 2 && n <= NL && D >= 1 && D <= kShortThreads, "need Nchain >= 1, 2 <= n <= %d, 1 <= D <= %d", NL, kShortThreads);
    HMC_REQUIRE(nlags >= 1 && nlags < NL, "need 1 <= nlags <= %d", NL - 1);
    HMC_REQUIRE(stride_chain >= 2 * n * D, "stride_chain too small");
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    const int spb = kShortThreads / D;
    const int grid = grid_for(2 * Nchain, spb, 8);            // 4 resident blocks per SM: two full waves
    const size_t smem = sizeof(double) * (NL + 3) * D;
    HMC_CUDA_CHECK(cudaMemsetAsync(out3xD, 0, sizeof(double) * 3 * D, stream));
    HMC_CUDA_CHECK(cudaMemsetAsync(out_lags, 0, sizeof(double) * nlags * D, stream));
    if (dtype == HMC_F32) {
        HMC_CUDA_CHECK(cudaFuncSetAttribute(diag_short_kernel<float, NL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        diag_short_kernel<float, NL, true><<<grid, kShortThreads, smem, stream>>>((const float*)q, Nchain, n, D, stride_chain, spb, nlags, out3xD, out_lags);
    } else if (dtype == HMC_F64) {
        HMC_CUDA_CHECK(cudaFuncSetAttribute(diag_short_kernel<double, NL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        diag_short_kernel<double, NL, true><<<grid, kShortThreads, smem, stream>>>((const double*)q, Nchain, n, D, stride_chain, spb, nlags, out3xD, out_lags);
    } else {
        hmc_set_error("dtype must be HMC_F32 or HMC_F64");
        return HMC_E_BADARG;
    }
    HMC_CUDA_CHECK(cudaGetLastError());
    return HMC_OK;
}
