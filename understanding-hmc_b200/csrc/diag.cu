// Diagnostics as HBM-bound reductions over the sample stream q_chain[Nchain][L_chain][D]:
//   hmc_diag_moments   -- split-chain mean / ddof=1 std  (utils.convergence_stats, /root/reference/utils.py:88-118)
//   hmc_diag_variogram -- variogram numerators for a chunk of lags (utils.variogram, utils.py:161-179)
// One thread owns one (split chain, dimension) time series at a time; consecutive threads own consecutive
// dimensions, so every load of a warp is one contiguous row segment.  Per-thread partials are float64,
// combined per block in shared memory and added to the output with one float64 atomic per (block, value).
#include "hmc_common.cuh"
#include <cstdlib>

namespace {

constexpr int kDiagThreads = 256;

// `q` points at sample 0 of chain 0; `stride_chain` elements between chains; a split chain s = 2*m + h is
// samples [h*n, h*n + n) of chain m (utils.py:102-104).
template <typename T>
__global__ void __launch_bounds__(kDiagThreads) diag_moments_kernel(const T* __restrict__ q, long Nchain, long n, int Dfull, long pitch,
                                                                    long stride_chain, int d0, int D, int spb, double* __restrict__ out) {
    // (D = dimensions of this tile, first dimension d0; `pitch` elements between consecutive samples)
    extern __shared__ double sm[];   // [3][D]
    for (int t = threadIdx.x; t < 3 * D; t += blockDim.x) sm[t] = 0.0;
    __syncthreads();
    const int d = threadIdx.x % D;
    const int sl = threadIdx.x / D;
    double s_std = 0.0, s_mean = 0.0, s_mean2 = 0.0;
    // chain means are accumulated relative to the first sample of this device's first chain (row 3 of `out`): sum mean^2 -
    // m mean^2 then does not cancel when the chains sit far from zero; the host combines devices with their own shifts
    const double cshift = (double)q[d0 + d];
    if (blockIdx.x == 0 && sl == 0) out[3 * Dfull + d0 + d] = cshift;
    if (sl < spb) {
        const long nseries = 2 * Nchain;
        for (long s = (long)blockIdx.x * spb + sl; s < nseries; s += (long)gridDim.x * spb) {
            const T* x = q + (s >> 1) * stride_chain + (s & 1) * n * pitch + d0 + d;
            const double x0 = (double)x[0];
            double a = 0.0, b = 0.0;
            long i = 0;
            for (; i + 4 <= n; i += 4) {
                const double v0 = (double)x[(i + 0) * pitch] - x0, v1 = (double)x[(i + 1) * pitch] - x0;
                const double v2 = (double)x[(i + 2) * pitch] - x0, v3 = (double)x[(i + 3) * pitch] - x0;
                a += (v0 + v1) + (v2 + v3);
                b += (v0 * v0 + v1 * v1) + (v2 * v2 + v3 * v3);
            }
            for (; i < n; ++i) { const double v = (double)x[i * pitch] - x0; a += v; b += v * v; }
            const double mean_s = a / (double)n;
            double var = (b - (double)n * mean_s * mean_s) / (double)(n - 1);   // ddof = 1 (utils.py:111)
            if (var < 0.0) var = 0.0;
            const double mean = mean_s + (x0 - cshift);
            s_std += sqrt(var);
            s_mean += mean;
            s_mean2 += mean * mean;
        }
        atomicAdd(&sm[d], s_std);
        atomicAdd(&sm[D + d], s_mean);
        atomicAdd(&sm[2 * D + d], s_mean2);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 3 * D; t += blockDim.x) atomicAdd(out + (t / D) * Dfull + d0 + (t % D), sm[t]);
}


// float32 stream, D % 4 == 0: one thread owns FOUR adjacent dimensions of one split chain and walks the time axis
// with 128-bit loads (8 rows in flight per thread), so that enough bytes are in flight to saturate HBM.
__global__ void __launch_bounds__(kDiagThreads) diag_moments_f32x4_kernel(const float* __restrict__ q, long Nchain, long n, int D, long pitch,
                                                                          long stride_chain, int spb, double* __restrict__ out) {
    extern __shared__ double sm[];   // [3][D]
    for (int t = threadIdx.x; t < 3 * D; t += blockDim.x) sm[t] = 0.0;
    __syncthreads();
    const int D4 = D >> 2;
    const int dq = threadIdx.x % D4;
    const int sl = threadIdx.x / D4;
    double s_std[4] = {0, 0, 0, 0}, s_mean[4] = {0, 0, 0, 0}, s_mean2[4] = {0, 0, 0, 0};
    double cshift[4] = {0, 0, 0, 0};             // see diag_moments_kernel
    if (sl < spb) {
        const float4 c4 = reinterpret_cast<const float4*>(q)[dq];
        cshift[0] = c4.x; cshift[1] = c4.y; cshift[2] = c4.z; cshift[3] = c4.w;
        if (blockIdx.x == 0 && sl == 0) { for (int c = 0; c < 4; ++c) out[3 * D + 4 * dq + c] = cshift[c]; }
        const long nseries = 2 * Nchain;
        for (long s = (long)blockIdx.x * spb + sl; s < nseries; s += (long)gridDim.x * spb) {
            const float4* x = reinterpret_cast<const float4*>(q + (s >> 1) * stride_chain + (s & 1) * n * pitch) + dq;
            const long P4 = pitch >> 2;
            const float4 x0 = x[0];
            // float partial sums over blocks of 8 rows (shifted by the first sample), float64 across blocks
            double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
            long i = 0;
            for (; i + 8 <= n; i += 8) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = x[(i + u) * P4];
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
                const float4 v = x[i * P4];
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
                const double mean = mean_s + (xs[c] - cshift[c]);
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
__global__ void __launch_bounds__(kShortThreads, 4) diag_short_kernel(const T* __restrict__ q, long Nchain, long n, int Dfull, long pitch,
                                                                  long stride_chain, int d0, int D, int spb, int nlags,
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
    const double cshift = (double)q[d0 + d];     // see diag_moments_kernel
    if (MOMENTS && blockIdx.x == 0 && sl == 0) mom_out[3 * Dfull + d0 + d] = cshift;
    if (sl < spb) {
        const long nseries = 2 * Nchain;
        for (long s = (long)blockIdx.x * spb + sl; s < nseries; s += (long)gridDim.x * spb) {
            const T* x = q + (s >> 1) * stride_chain + (s & 1) * n * pitch + d0 + d;
            T v[NMAX];
#pragma unroll
            for (int i = 0; i < NMAX; ++i) v[i] = (i < n) ? x[(long)i * pitch] : T(0);
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
                const double mean = mean_s + (x0 - cshift);
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
    for (int t = threadIdx.x; t < nlags * D; t += blockDim.x) atomicAdd(out + (t / D) * Dfull + d0 + (t % D), sm[t]);
    if (MOMENTS) for (int t = threadIdx.x; t < 3 * D; t += blockDim.x) atomicAdd(mom_out + (t / D) * Dfull + d0 + (t % D), sm[NMAX * D + t]);
}

// Lags t = lag0 + k, k < NL.  For every i the pair (x[i], x[i - lag0 - k]) contributes (x[i]-x[i-lag0-k])^2.
// The last NL delayed values live in a register window addressed with compile-time indices (the time loop is
// unrolled by NL); float partial sums are flushed into float64 every NL steps.
template <typename T, int NL>
__global__ void __launch_bounds__(kDiagThreads) diag_variogram_kernel(const T* __restrict__ q, long Nchain, long n, int Dfull, long pitch,
                                                                      long stride_chain, int d0, int D, int spb, int lag0, int nlags,
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
            const T* x = q + (s >> 1) * stride_chain + (s & 1) * n * pitch + d0 + d;
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
                        const T xa = x[i * pitch];
                        w[r] = x[(i - lag0) * pitch];
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
                for (int r = 0; r < NL; ++r) { xa[r] = x[(i0 + r) * pitch]; wn[r] = x[(i0 + r - lag0) * pitch]; }
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
                        const T xa = x[i * pitch];
                        w[r] = x[(i - lag0) * pitch];
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
    for (int t = threadIdx.x; t < nlags * D; t += blockDim.x) atomicAdd(out + (t / D) * Dfull + d0 + (t % D), sm[t]);
}


// ---------------------------------------------------------------------------------------------------------------------
// Packed float32 kernels (D even): one thread owns TWO adjacent dimensions of one split chain, loads them as one 64-bit
// word per sample and works on both with the packed FADD2 / FFMA2 instructions: one instruction per (sample, lag, dim).
// ---------------------------------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 sqacc2(f32x2 d, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(r) : "l"(d), "l"(c)); return r; }
__device__ __forceinline__ float lo_of(f32x2 v) { return __uint_as_float((unsigned int)v); }
__device__ __forceinline__ float hi_of(f32x2 v) { return __uint_as_float((unsigned int)(v >> 32)); }

// Windowed variogram, lags t = lag0 + k, k < NL (utils.py:161-179).  Steps j = i - lag0 = 0, 1, ...: the pair
// (x[lag0 + j], x[j - k]) contributes for k <= j.  The last NL delayed values live in a register window addressed with
// compile-time indices (the step loop is unrolled by NL); both streams of half a block (NL / 2 steps) are requested up front.
// Partial sums stay float32 within one split chain (n terms of like magnitude) and are widened once per chain.
template <int NL>
__global__ void __launch_bounds__(256, 2) diag_variogram_f32x2_kernel(const float* __restrict__ q, long Nchain, long n, int D, long pitch,
                                                                    long stride_chain, int d0, int Dt, int spb, int lag0, int nlags,
                                                                    double* __restrict__ out) {
    extern __shared__ double sm[];   // [NL][Dt]
    for (int t = threadIdx.x; t < NL * Dt; t += blockDim.x) sm[t] = 0.0;
    __syncthreads();
    const int D2 = Dt >> 1;
    const int dp = threadIdx.x % D2, sl = threadIdx.x / D2;
    constexpr int H = NL / 2;
    if (sl < spb) {
        const long nseries = 2 * Nchain;
        const long p2 = pitch >> 1;
        for (long s = (long)blockIdx.x * spb + sl; s < nseries; s += (long)gridDim.x * spb) {
            const f32x2* x = reinterpret_cast<const f32x2*>(q + (s >> 1) * stride_chain + (s & 1) * n * pitch + d0) + dp;
            f32x2 w[NL], acc[NL];
#pragma unroll
            for (int k = 0; k < NL; ++k) { w[k] = 0ull; acc[k] = 0ull; }
            const long nsteps = n - lag0;                // j = 0 .. nsteps - 1
            long j0 = 0;
            if (nsteps > 0) {                            // first block: entry k valid only for k <= r (compile time)
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                    f32x2 xa[H], wn[H];
#pragma unroll
                    for (int r = 0; r < H; ++r) {
                        const long j = hb * H + r;
                        xa[r] = (j < nsteps) ? x[(lag0 + j) * p2] : 0ull;
                        wn[r] = (j < nsteps) ? x[j * p2] : 0ull;
                    }
#pragma unroll
                    for (int r = 0; r < H; ++r) {
                        const int rr = hb * H + r;
                        if (rr < nsteps) {
                            w[rr] = wn[r];
#pragma unroll
                            for (int k = 0; k < NL; ++k)
                                if (k <= rr) acc[k] = sqacc2(sub2(xa[r], w[(rr - k + NL) % NL]), acc[k]);
                        }
                    }
                }
                j0 = NL;
            }
            for (; j0 + NL <= nsteps; j0 += NL) {        // steady state, no conditions
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                    f32x2 xa[H], wn[H];
#pragma unroll
                    for (int r = 0; r < H; ++r) { xa[r] = x[(lag0 + j0 + hb * H + r) * p2]; wn[r] = x[(j0 + hb * H + r) * p2]; }
#pragma unroll
                    for (int r = 0; r < H; ++r) {
                        const int rr = hb * H + r;
                        w[rr] = wn[r];
#pragma unroll
                        for (int k = 0; k < NL; ++k) acc[k] = sqacc2(sub2(xa[r], w[(rr - k + NL) % NL]), acc[k]);
                    }
                }
            }
            if (j0 < nsteps) {                           // tail
#pragma unroll
                for (int rr = 0; rr < NL; ++rr) {
                    const long j = j0 + rr;
                    if (j < nsteps) {
                        const f32x2 xa = x[(lag0 + j) * p2];
                        w[rr] = x[j * p2];
#pragma unroll
                        for (int k = 0; k < NL; ++k) acc[k] = sqacc2(sub2(xa, w[(rr - k + NL) % NL]), acc[k]);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                if (k < nlags) {
                    atomicAdd(&sm[k * Dt + 2 * dp], (double)lo_of(acc[k]));
                    atomicAdd(&sm[k * Dt + 2 * dp + 1], (double)hi_of(acc[k]));
                }
            }
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nlags * Dt; t += blockDim.x) atomicAdd(out + (t / Dt) * D + d0 + (t % Dt), sm[t]);
}

__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { return (f32x2)__float_as_uint(lo) | ((f32x2)__float_as_uint(hi) << 32); }

// Windowed variogram in CROSS-PRODUCT form: with y = x - x[0] (the chain's first sample: the sums are shift invariant),
//   sum_{i >= t} (y_i - y_{i-t})^2 = sum_{i >= t} y_i^2 + sum_{i < n-t} y_i^2 - 2 sum_{i >= t} y_i y_{i-t},
// so a (sample, lag) pair costs ONE fused multiply-add instead of a subtraction and a multiply-add: the float32 pipe drops
// below the HBM time of a pass.  The two sums of squares of lag t = lag0 + k are those of lag0 (accumulated along the two
// streams the kernel reads anyway) minus the first / last k terms, which are folded into the lag's accumulator (half each)
// from 2 (NL - 1) re-read samples per chain.  The float32 cancellation costs ~1e-6 / (1 - rho_t) relative accuracy of V_t,
// i.e. ~1e-6 absolute in rho_t (measured: n_eff of the Case 3c bench run equal to the difference form's to 1e-9).  Same
// window / unrolling scheme as diag_variogram_f32x2_kernel.  MEASURED: with 16 lags per pass it is no faster than the
// difference form (end-to-end step 250 vs 245 ms: both ride the HBM time of a pass at 400-sample chains; at 125 samples
// the per-chain prologue / epilogue makes it slower, 2.53 vs 2.20 ms), and a 32-lags-in-one-pass instantiation (window +
// accumulators = 128 registers, one 320-thread block per SM) fell into local memory and took 9.7 ms per 32 lags against
// 4.4 ms for two 16-lag passes.  So the difference form stays the default; HMC_B200_DIAG_CROSS=1 selects this kernel.
template <int NL, int H, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) diag_variogram_cross_kernel(const float* __restrict__ q, long Nchain, long n, int D, long pitch,
                                                                    long stride_chain, int d0, int Dt, int spb, int lag0, int nlags,
                                                                    double* __restrict__ out) {
    extern __shared__ double sm[];   // [NL][Dt]
    for (int t = threadIdx.x; t < NL * Dt; t += blockDim.x) sm[t] = 0.0;
    __syncthreads();
    const int D2 = Dt >> 1;
    const int dp = threadIdx.x % D2, sl = threadIdx.x / D2;
    constexpr int NHB = NL / H;                 // the streams are requested H steps at a time
    if (sl < spb) {
        const long nseries = 2 * Nchain;
        const long p2 = pitch >> 1;
        for (long s = (long)blockIdx.x * spb + sl; s < nseries; s += (long)gridDim.x * spb) {
            const f32x2* x = reinterpret_cast<const f32x2*>(q + (s >> 1) * stride_chain + (s & 1) * n * pitch + d0) + dp;
            const f32x2 c = x[0];
            f32x2 w[NL], acc[NL];       // acc[k] = sum y_i y_{i - lag0 - k}  (+ half the excluded squares, folded in below)
            f32x2 sa = 0ull, sb = 0ull;  // sum of squares of the two streams: y_i (i >= lag0) and y_j (j < n - lag0)
#pragma unroll
            for (int k = 0; k < NL; ++k) { w[k] = 0ull; acc[k] = 0ull; }
            const long nsteps = n - lag0;
            long j0 = 0;
            if (nsteps > 0) {
#pragma unroll
                for (int hb = 0; hb < NHB; ++hb) {
                    f32x2 xa[H], wn[H];
#pragma unroll
                    for (int r = 0; r < H; ++r) {
                        const long j = hb * H + r;
                        xa[r] = (j < nsteps) ? sub2(x[(lag0 + j) * p2], c) : 0ull;
                        wn[r] = (j < nsteps) ? sub2(x[j * p2], c) : 0ull;
                    }
#pragma unroll
                    for (int r = 0; r < H; ++r) {
                        const int rr = hb * H + r;
                        if (rr < nsteps) {
                            w[rr] = wn[r];
                            sa = fma2(xa[r], xa[r], sa); sb = fma2(wn[r], wn[r], sb);
#pragma unroll
                            for (int k = 0; k < NL; ++k)
                                if (k <= rr) acc[k] = fma2(xa[r], w[(rr - k + NL) % NL], acc[k]);
                        }
                    }
                }
                j0 = NL;
            }
            for (; j0 + NL <= nsteps; j0 += NL) {
#pragma unroll
                for (int hb = 0; hb < NHB; ++hb) {
                    f32x2 xa[H], wn[H];
#pragma unroll
                    for (int r = 0; r < H; ++r) { xa[r] = sub2(x[(lag0 + j0 + hb * H + r) * p2], c); wn[r] = sub2(x[(j0 + hb * H + r) * p2], c); }
#pragma unroll
                    for (int r = 0; r < H; ++r) {
                        const int rr = hb * H + r;
                        w[rr] = wn[r];
                        sa = fma2(xa[r], xa[r], sa); sb = fma2(wn[r], wn[r], sb);
#pragma unroll
                        for (int k = 0; k < NL; ++k) acc[k] = fma2(xa[r], w[(rr - k + NL) % NL], acc[k]);
                    }
                }
            }
            if (j0 < nsteps) {
#pragma unroll
                for (int rr = 0; rr < NL; ++rr) {
                    const long j = j0 + rr;
                    if (j < nsteps) {
                        const f32x2 xa = sub2(x[(lag0 + j) * p2], c);
                        w[rr] = sub2(x[j * p2], c);
                        sa = fma2(xa, xa, sa); sb = fma2(w[rr], w[rr], sb);
#pragma unroll
                        for (int k = 0; k < NL; ++k) acc[k] = fma2(xa, w[(rr - k + NL) % NL], acc[k]);
                    }
                }
            }
            // lag lag0 + k leaves out the first k samples of the upper stream and the last k of the lower one
            if (nsteps > 0) {
                f32x2 pre = 0ull, suf = 0ull;
                const f32x2 half2 = pack2(0.5f, 0.5f);
#pragma unroll
                for (int t = 0; t < NL - 1; ++t) {
                    if (t < nsteps) {
                        const f32x2 a1 = sub2(x[(lag0 + t) * p2], c), b1 = sub2(x[(nsteps - 1 - t) * p2], c);
                        pre = fma2(a1, a1, pre); suf = fma2(b1, b1, suf);
                        acc[t + 1] = fma2(half2, pre, acc[t + 1]);
                        acc[t + 1] = fma2(half2, suf, acc[t + 1]);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                if (k < nlags && k < nsteps) {       // sum (y_i - y_{i-t})^2 = sa + sb - 2 (cross + half the excluded squares)
                    atomicAdd(&sm[k * Dt + 2 * dp], ((double)lo_of(sa) + (double)lo_of(sb)) - 2.0 * (double)lo_of(acc[k]));
                    atomicAdd(&sm[k * Dt + 2 * dp + 1], ((double)hi_of(sa) + (double)hi_of(sb)) - 2.0 * (double)hi_of(acc[k]));
                }
            }
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nlags * Dt; t += blockDim.x) atomicAdd(out + (t / Dt) * D + d0 + (t % Dt), sm[t]);
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

// Dimensions are processed in tiles of at most kDiagThreads (the thread <-> (series, dimension) mapping of the generic
// kernels); D is limited only by the 32-bit tile arithmetic.
static const int kMaxD = 1 << 20;

static bool packed_ok(int32_t dtype, const void* q, int32_t D, int64_t pitch, int64_t stride_chain, int64_t n) {
    return dtype == HMC_F32 && (D % 2) == 0 && (pitch % 2) == 0 && (stride_chain % 2) == 0 && ((n * pitch) % 2) == 0 &&
           (reinterpret_cast<uintptr_t>(q) % 8) == 0 && getenv("HMC_B200_DIAG_GENERIC") == nullptr;
}

extern "C" int hmc_diag_moments(int32_t dtype, const void* q, int64_t Nchain, int64_t n, int32_t D, int64_t stride_chain,
                                double* out4xD, void* cuda_stream) {
    HMC_REQUIRE(dtype == HMC_F32 || dtype == HMC_F64, "dtype must be HMC_F32 or HMC_F64");
    HMC_REQUIRE(q && out4xD, "NULL buffer");
    HMC_REQUIRE(Nchain >= 1 && n >= 2 && D >= 1 && D <= kMaxD, "need Nchain >= 1, n >= 2, 1 <= D <= %d", kMaxD);
    HMC_REQUIRE(stride_chain >= 2 * n * D, "stride_chain too small");
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    const long pitch = D;
    HMC_CUDA_CHECK(cudaMemsetAsync(out4xD, 0, sizeof(double) * 4 * D, stream));
    const bool vec4 = dtype == HMC_F32 && (D % 4) == 0 && D <= 4 * kDiagThreads && (stride_chain % 4) == 0 && (n * D) % 4 == 0 &&
                      (reinterpret_cast<uintptr_t>(q) % 16) == 0;
    if (vec4) {
        const int spb4 = kDiagThreads / (D / 4);
        diag_moments_f32x4_kernel<<<grid_for(2 * Nchain, spb4), kDiagThreads, sizeof(double) * 3 * D, stream>>>((const float*)q, Nchain, n, D, pitch, stride_chain, spb4, out4xD);
    } else {
        for (int d0 = 0; d0 < D; d0 += kDiagThreads) {
            const int Dt = (D - d0 < kDiagThreads) ? D - d0 : kDiagThreads;
            const int spb = kDiagThreads / Dt;
            const int grid = grid_for(2 * Nchain, spb);
            const size_t smem = sizeof(double) * 3 * Dt;
            if (dtype == HMC_F32)
                diag_moments_kernel<float><<<grid, kDiagThreads, smem, stream>>>((const float*)q, Nchain, n, D, pitch, stride_chain, d0, Dt, spb, out4xD);
            else
                diag_moments_kernel<double><<<grid, kDiagThreads, smem, stream>>>((const double*)q, Nchain, n, D, pitch, stride_chain, d0, Dt, spb, out4xD);
        }
    }
    HMC_CUDA_CHECK(cudaGetLastError());
    return HMC_OK;
}

template <typename T, bool MOMENTS>
static int launch_short_generic(const T* q, int64_t Nchain, int64_t n, int32_t D, long pitch, int64_t stride_chain, int nlags,
                                double* mom, double* out, cudaStream_t stream) {
    constexpr int NL = 32;
    for (int d0 = 0; d0 < D; d0 += kShortThreads) {
        const int Dt = (D - d0 < kShortThreads) ? D - d0 : kShortThreads;
        const int spb = kShortThreads / Dt;
        const size_t smem = sizeof(double) * (NL + 3) * Dt;
        HMC_CUDA_CHECK(cudaFuncSetAttribute(diag_short_kernel<T, NL, MOMENTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        diag_short_kernel<T, NL, MOMENTS><<<grid_for(2 * Nchain, spb, 8), kShortThreads, smem, stream>>>(q, Nchain, n, D, pitch, stride_chain, d0, Dt, spb, nlags, mom, out);
    }
    return HMC_OK;
}

extern "C" int hmc_diag_variogram(int32_t dtype, const void* q, int64_t Nchain, int64_t n, int32_t D, int64_t stride_chain,
                                  int32_t lag0, int32_t nlags, double* out, void* cuda_stream) {
    constexpr int NL = 32;
    HMC_REQUIRE(dtype == HMC_F32 || dtype == HMC_F64, "dtype must be HMC_F32 or HMC_F64");
    HMC_REQUIRE(q && out, "NULL buffer");
    HMC_REQUIRE(Nchain >= 1 && n >= 2 && D >= 1 && D <= kMaxD, "need Nchain >= 1, n >= 2, 1 <= D <= %d", kMaxD);
    HMC_REQUIRE(lag0 >= 1 && nlags >= 1 && nlags <= NL, "need lag0 >= 1 and 1 <= nlags <= %d", NL);
    HMC_REQUIRE(stride_chain >= 2 * n * D, "stride_chain too small");
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    const long pitch = D;
    HMC_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(double) * nlags * D, stream));
    const bool packed = packed_ok(dtype, q, D, pitch, stride_chain, n);
    if (lag0 == 1 && n <= NL) {                  // short series: one pass, values held in registers
        int rc;
        if (dtype == HMC_F32) rc = launch_short_generic<float, false>((const float*)q, Nchain, n, D, pitch, stride_chain, nlags, nullptr, out, stream);
        else rc = launch_short_generic<double, false>((const double*)q, Nchain, n, D, pitch, stride_chain, nlags, nullptr, out, stream);
        if (rc) return rc;
        HMC_CUDA_CHECK(cudaGetLastError());
        return HMC_OK;
    }
    if (packed) {
        const bool force_cross = getenv("HMC_B200_DIAG_CROSS") != nullptr;
        // 16 lags per pass, two resident blocks per SM (the 32-lag difference-form window needs 190+ registers and leaves one
        // block per SM waiting on its loads); a 32-lag request is two passes, the second one reads the samples again
        constexpr int NLP = 16;
        for (int l0 = 0; l0 < nlags; l0 += NLP) {
            const int nl = (nlags - l0 < NLP) ? nlags - l0 : NLP;
            for (int d0 = 0; d0 < D; d0 += 512) {
                const int Dt = (D - d0 < 512) ? D - d0 : 512;
                const int spb = 256 / (Dt / 2);
                const size_t smem = sizeof(double) * NLP * Dt;
                if (!force_cross) {
                    HMC_CUDA_CHECK(cudaFuncSetAttribute(diag_variogram_f32x2_kernel<NLP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    diag_variogram_f32x2_kernel<NLP><<<grid_for(2 * Nchain, spb, 2), 256, smem, stream>>>((const float*)q, Nchain, n, D, pitch, stride_chain, d0, Dt, spb, lag0 + l0, nl, out + (size_t)l0 * D);
                } else {
                    auto kern = diag_variogram_cross_kernel<NLP, 8, 256, 2>;
                    HMC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    kern<<<grid_for(2 * Nchain, spb, 2), 256, smem, stream>>>((const float*)q, Nchain, n, D, pitch, stride_chain, d0, Dt, spb, lag0 + l0, nl, out + (size_t)l0 * D);
                }
            }
        }
        HMC_CUDA_CHECK(cudaGetLastError());
        return HMC_OK;
    }
    for (int d0 = 0; d0 < D; d0 += kDiagThreads) {
        const int Dt = (D - d0 < kDiagThreads) ? D - d0 : kDiagThreads;
        const int spb = kDiagThreads / Dt;
        const int grid = grid_for(2 * Nchain, spb);
        const size_t smem = sizeof(double) * NL * Dt;
        if (dtype == HMC_F32) {
            HMC_CUDA_CHECK(cudaFuncSetAttribute(diag_variogram_kernel<float, NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            diag_variogram_kernel<float, NL><<<grid, kDiagThreads, smem, stream>>>((const float*)q, Nchain, n, D, pitch, stride_chain, d0, Dt, spb, lag0, nlags, out);
        } else {
            HMC_CUDA_CHECK(cudaFuncSetAttribute(diag_variogram_kernel<double, NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            diag_variogram_kernel<double, NL><<<grid, kDiagThreads, smem, stream>>>((const double*)q, Nchain, n, D, pitch, stride_chain, d0, Dt, spb, lag0, nlags, out);
        }
    }
    HMC_CUDA_CHECK(cudaGetLastError());
    return HMC_OK;
}

extern "C" int hmc_diag_short_series(int32_t dtype, const void* q, int64_t Nchain, int64_t n, int32_t D, int64_t stride_chain,
                                     int32_t nlags, double* out4xD, double* out_lags, void* cuda_stream) {
    constexpr int NL = 32;
    HMC_REQUIRE(dtype == HMC_F32 || dtype == HMC_F64, "dtype must be HMC_F32 or HMC_F64");
    HMC_REQUIRE(q && out4xD && out_lags, "NULL buffer");
    HMC_REQUIRE(Nchain >= 1 && n >= 2 && n <= NL && D >= 1 && D <= kMaxD, "need Nchain >= 1, 2 <= n <= %d, 1 <= D <= %d", NL, kMaxD);
    HMC_REQUIRE(nlags >= 1 && nlags < NL, "need 1 <= nlags <= %d", NL - 1);
    HMC_REQUIRE(stride_chain >= 2 * n * D, "stride_chain too small");
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    const long pitch = D;
    HMC_CUDA_CHECK(cudaMemsetAsync(out4xD, 0, sizeof(double) * 4 * D, stream));
    HMC_CUDA_CHECK(cudaMemsetAsync(out_lags, 0, sizeof(double) * nlags * D, stream));
    int rc;
    if (dtype == HMC_F32) rc = launch_short_generic<float, true>((const float*)q, Nchain, n, D, pitch, stride_chain, nlags, out4xD, out_lags, stream);
    else rc = launch_short_generic<double, true>((const double*)q, Nchain, n, D, pitch, stride_chain, nlags, out4xD, out_lags, stream);
    if (rc) return rc;
    HMC_CUDA_CHECK(cudaGetLastError());
    return HMC_OK;
}
