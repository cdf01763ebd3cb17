// All lags of the variogram in ONE pass over the sample stream (utils.variogram for every t, /root/reference/utils.py:141-152,
// 161-179): the O(n T) windowed differences of diag.cu become one 1024-point FFT per (chain, dimension).
//
//   sum_i (y[i+t] - y[i])^2 = sum_{i >= t} y_i^2 + sum_{i < n-t} y_i^2 - 2 R_t,      R_t = sum_i y_i y_{i+t},
//
// with y = x - x[0] per split chain (the variogram is shift invariant; the shift keeps the float32 cancellation small).
// The two halves of a chain (split chains 2m and 2m+1, utils.py:102-104) ride in ONE complex transform z = a + i b, zero
// padded to N = 1024 >= 2n - 1 (n <= 512): |Z_k|^2 + |Z_{N-k}|^2 = 2 (|A_k|^2 + |B_k|^2), so the cosine transform of
// P_k = sum over chains of |Z_k|^2 is R_t(a) + R_t(b) summed over the chains -- the cross term is odd in k and drops out.
//
// Geometry: a block owns FOUR adjacent dimensions and walks over chains (stride = blocks per dimension tile); the chain's 2n
// rows x 16 bytes arrive by cp.async while the previous chain is transformed.  One warp = one dimension: N = 32 x 32,
//   step 1  lane n2 transforms x[32 n1 + n2] over n1 (32-point radix-2 DIF in registers; inputs n1 >= 16 are zero),
//   step 2  twiddle W_1024^(n2 k1) from a shared table laid out [k1][n2],
//   step 3  32 x 32 transpose through a padded shared tile (real plane, then imaginary plane),
//   step 4  lane k1 transforms over n2, |X|^2 is added to 32 float accumulators per lane (P_k for k = k1 + 32 k2),
// and every kFlush chains the accumulators are widened into the float64 power spectrum P[D][1024] (fire-and-forget
// atomics).  The per-row sums of squares Q[2n][D] are kept by the thread that loads the row.  A small float64 kernel
// turns (P, Q) into the variogram numerators of all lags.  ~60 kflop per (chain, dimension) instead of 2 n T.
#include "hmc_common.cuh"
#include <cstdlib>

namespace {

constexpr int kN = 1024;          // transform length
constexpr int kHalf = 512;        // largest n (samples per split chain)
constexpr int kThreads = 128;     // 4 warps = 4 dimensions
constexpr int kZRow = 32 * 33 / 2;// float2 per z row (>= kHalf): the row doubles as a 32 x 33 float transpose plane
constexpr int kFlush = 32;        // chains between two widenings of the float accumulators
constexpr int kStepSync = 8;      // chains between two check-ins of the blocks that read the same rows (see `bars`)
constexpr int kStepLead = 1;      // check-ins a block may be ahead of its slowest sibling (absorbs the jitter between SMs)
constexpr int kMaxGroups = 64;

// cos(2 pi j / 32), j = 0..8
__host__ __device__ __forceinline__ constexpr float c32(int j) {
    return j == 0 ? 1.0f : j == 1 ? 0.98078528040323044913f : j == 2 ? 0.92387953251128675613f : j == 3 ? 0.83146961230254523708f
         : j == 4 ? 0.70710678118654752440f : j == 5 ? 0.55557023301960222474f : j == 6 ? 0.38268343236508977173f
         : j == 7 ? 0.19509032201612826785f : 0.0f;
}
// W_32^j = exp(-2 pi i j / 32), j = 0..15
__host__ __device__ __forceinline__ constexpr float w32re(int j) { return j <= 8 ? c32(j) : -c32(16 - j); }
__host__ __device__ __forceinline__ constexpr float w32im(int j) { return j <= 8 ? -c32(8 - j) : -c32(j - 8); }
__host__ __device__ __forceinline__ constexpr int brev5(int p) {
    return ((p & 1) << 4) | ((p & 2) << 2) | (p & 4) | ((p & 8) >> 2) | ((p & 16) >> 4);
}

// (dr + i di) * W_32^j with the trivial factors folded at compile time (j is a constant after unrolling)
__host__ __device__ __forceinline__ void mul_w32(float dr, float di, int j, float& outr, float& outi) {
    if (j == 0) { outr = dr; outi = di; }
    else if (j == 8) { outr = di; outi = -dr; }                 // W = -i
    else if (j == 4) { const float s = 0.70710678118654752440f; outr = s * (dr + di); outi = s * (di - dr); }
    else if (j == 12) { const float s = 0.70710678118654752440f; outr = s * (di - dr); outi = -s * (dr + di); }
    else {
        const float wr = w32re(j), wi = w32im(j);
        outr = fmaf(dr, wr, -di * wi);
        outi = fmaf(dr, wi, di * wr);
    }
}

// In-place 32-point radix-2 decimation-in-frequency transform; the result for frequency brev5(p) is left at position p.
// ZERO_TOP: inputs 16..31 are known to be zero (first stage degenerates to one complex multiply per pair).
template <bool ZERO_TOP>
__host__ __device__ __forceinline__ void fft32(float (&re)[32], float (&im)[32]) {
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int half = 16 >> s;
#pragma unroll
        for (int g = 0; g < (1 << s); ++g) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const int a = g * 2 * half + j, b = a + half;
                if (ZERO_TOP && s == 0) {
                    mul_w32(re[a], im[a], j, re[b], im[b]);
                } else {
                    const float dr = re[a] - re[b], di = im[a] - im[b];
                    re[a] += re[b];
                    im[a] += im[b];
                    mul_w32(dr, di, j << s, re[b], im[b]);
                }
            }
        }
    }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Shared memory: twiddles float2 [32][32] | z float2 [4][512] | transpose float [4][32 * 33] | raw float4 [2n] | Qs float4 [2n]
//
// `bars` (cooperative launch only, else NULL): the D / 4 blocks of a group read the SAME rows of the same chains, 16 bytes each,
// and every 64-byte DRAM atom serves four of them -- if they pass by within the L2's reach.  Left alone they drift apart (an
// SM holds two or three blocks) and the stream is fetched 3.4 times (ncu: 70 GB for 21 GB).  So the blocks of a group check in
// at a counter every kStepSync chains and wait while they are more than kStepLead check-ins ahead of the slowest sibling; a
// cooperative launch guarantees that all of them are resident, and a spin limit turns a missing sibling into lost sharing
// instead of a hang.  Measured DRAM reads for the 21 GB stream: 21.1 GB in strict lock step (lead 0), 42 GB with the one check-in
// of lead shipped here, 70 GB without check-ins -- at the same kernel time (the kernel is issue / shared-memory-latency bound).
__global__ void __launch_bounds__(kThreads, 3) diag_fft_power_kernel(const float* __restrict__ q, long Nchain, int n, int D, long stride_chain,
                                                                    int groups, double* __restrict__ P, double* __restrict__ Q,
                                                                    unsigned int* __restrict__ bars) {
    extern __shared__ __align__(16) unsigned char smraw[];
    float2* tw = reinterpret_cast<float2*>(smraw);                       // [p][lane] = W_1024^(lane * brev5(p))
    float2* z = tw + 32 * 32;                                            // [4][kZRow]: 512 complex inputs; after step 1 the row is the
                                                                         // imaginary plane of the warp's 32 x 33 transpose tile
    float* tr = reinterpret_cast<float*>(z + 4 * kZRow);                 // [4][32 * 33]
    float4* raw = reinterpret_cast<float4*>(tr + 4 * 32 * 33);           // [2n]
    float4* Qs = raw + 2 * n;                                            // [2n]
    __shared__ int sync_lost;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int dtile = blockIdx.x / groups, grp = blockIdx.x % groups;
    if (tid == 0) sync_lost = 0;
    const int d0 = dtile * 4;
    const int rows = 2 * n;

    for (int t = tid; t < 32 * 32; t += kThreads) {
        const int p = t >> 5, l = t & 31;
        float s, c;
        sincospif(-(float)((l * brev5(p)) & (kN - 1)) * (2.0f / kN), &s, &c);
        tw[t] = make_float2(c, s);
    }
    for (int t = tid; t < 4 * kZRow; t += kThreads) z[t] = make_float2(0.f, 0.f);
    for (int t = tid; t < rows; t += kThreads) Qs[t] = make_float4(0.f, 0.f, 0.f, 0.f);

    float acc[32];
#pragma unroll
    for (int p = 0; p < 32; ++p) acc[p] = 0.f;

    auto issue = [&](long chain) {
        const float* src = q + chain * stride_chain + d0;
        for (int r = tid; r < rows; r += kThreads) cp_async16(raw + r, src + (long)r * D);
    };
    auto flush = [&]() {
        double* Pd = P + (long)(d0 + warp) * kN + lane;
#pragma unroll
        for (int p = 0; p < 32; ++p) {
            atomicAdd(Pd + 32 * brev5(p), (double)acc[p]);
            acc[p] = 0.f;
        }
        for (int r = tid; r < rows; r += kThreads) {                     // rows are owned by the thread that loads them
            const float4 v = Qs[r];
            double* Qd = Q + (long)r * D + d0;
            atomicAdd(Qd + 0, (double)v.x);
            atomicAdd(Qd + 1, (double)v.y);
            atomicAdd(Qd + 2, (double)v.z);
            atomicAdd(Qd + 3, (double)v.w);
            Qs[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };

    long chain = grp;
    if (chain < Nchain) issue(chain);
    int since = 0, step = 0;
    const unsigned int siblings = gridDim.x / groups;
    bool in_step = bars != nullptr;
    float* mytr = tr + warp * (32 * 33);
    const float2* myz = z + warp * kZRow;
    float* myzp = reinterpret_cast<float*>(z + warp * kZRow);
    for (; chain < Nchain; chain += groups) {
        cp_async_wait_all();
        __syncthreads();                                   // rows of this chain have landed; every warp is done with z
        {
            const float4 s0 = raw[0], s1 = raw[n];         // shifts: first sample of each split chain
            float* zf = reinterpret_cast<float*>(z);
            for (int r = tid; r < rows; r += kThreads) {
                const bool second = r >= n;
                const float4 v = raw[r];
                const float4 sh = second ? s1 : s0;
                const float y0 = v.x - sh.x, y1 = v.y - sh.y, y2 = v.z - sh.z, y3 = v.w - sh.w;
                float4 acc4 = Qs[r];
                acc4.x = fmaf(y0, y0, acc4.x); acc4.y = fmaf(y1, y1, acc4.y); acc4.z = fmaf(y2, y2, acc4.z); acc4.w = fmaf(y3, y3, acc4.w);
                Qs[r] = acc4;
                const int i = second ? r - n : r, comp = second ? 1 : 0;
                zf[(0 * kZRow + i) * 2 + comp] = y0;
                zf[(1 * kZRow + i) * 2 + comp] = y1;
                zf[(2 * kZRow + i) * 2 + comp] = y2;
                zf[(3 * kZRow + i) * 2 + comp] = y3;
            }
        }
        __syncthreads();                                   // z complete, raw free
        if (chain + groups < Nchain) issue(chain + groups);

        float re[32], im[32];
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            const float2 v = (32 * n1 + lane < n) ? myz[32 * n1 + lane] : make_float2(0.f, 0.f);   // (beyond n the row holds the previous
            re[n1] = v.x; im[n1] = v.y;                                                             //  chain's transpose plane)
        }
#pragma unroll
        for (int n1 = 16; n1 < 32; ++n1) { re[n1] = 0.f; im[n1] = 0.f; }
        fft32<true>(re, im);
#pragma unroll
        for (int p = 0; p < 32; ++p) {                     // twiddle (all table loads ahead of the transpose's stores: the compiler
            const float2 w = tw[p * 32 + lane];            //  cannot tell that the two shared arrays do not alias)
            const float r2 = fmaf(re[p], w.x, -im[p] * w.y);
            im[p] = fmaf(re[p], w.y, im[p] * w.x);
            re[p] = r2;
        }
        __syncwarp();                                      // every lane has read its part of z: the row doubles as the second plane
#pragma unroll
        for (int p = 0; p < 32; ++p) { mytr[brev5(p) * 33 + lane] = re[p]; myzp[brev5(p) * 33 + lane] = im[p]; }
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 32; ++n2) { re[n2] = mytr[lane * 33 + n2]; im[n2] = myzp[lane * 33 + n2]; }
        __syncwarp();
        fft32<false>(re, im);
#pragma unroll
        for (int p = 0; p < 32; ++p) acc[p] = fmaf(re[p], re[p], fmaf(im[p], im[p], acc[p]));

        if (++since == kFlush) { flush(); since = 0; }
        if (in_step && (++step % kStepSync) == 0) {
            __syncthreads();
            if (tid == 0) {
                const int behind = step / kStepSync - kStepLead;          // everybody must have finished this check-in
                const unsigned int target = behind > 0 ? (unsigned int)behind * siblings : 0u;
                atomicAdd(bars + grp, 1u);
                const long long t0 = clock64();
                while (*reinterpret_cast<volatile unsigned int*>(bars + grp) < target) {
                    __nanosleep(100);
                    if (clock64() - t0 > 200000000ll) { *reinterpret_cast<volatile int*>(&sync_lost) = 1; break; }
                }
            }
            __syncthreads();
            if (*reinterpret_cast<volatile int*>(&sync_lost)) in_step = false;      // (same decision in every warp: read after the barrier)
        }
    }
    if (since) flush();
}

// Variogram numerators of lags 1..nlags from the power spectrum and the per-row sums of squares (float64):
//   out[t - 1][d] = sum_{i >= t} Qt[i] + sum_{i < n - t} Qt[i] - (2 / N) sum_k P[d][k] cos(2 pi k t / N),   Qt[i] = Q[i] + Q[n + i].
__global__ void __launch_bounds__(kHalf) diag_fft_finish_kernel(const double* __restrict__ P, const double* __restrict__ Q, int n, int D,
                                                                int nlags, double* __restrict__ out) {
    __shared__ double Pd[kN], ct[kN], pre[kHalf + 1];
    const int d = blockIdx.x, tid = threadIdx.x;
    for (int k = tid; k < kN; k += blockDim.x) {
        Pd[k] = P[(long)d * kN + k];
        ct[k] = cospi(2.0 * k / kN);
    }
    for (int i = tid; i < n; i += blockDim.x) pre[i + 1] = Q[(long)i * D + d] + Q[(long)(n + i) * D + d];
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        pre[0] = 0.0;
        for (int i = 1; i <= n; ++i) { s += pre[i]; pre[i] = s; }
    }
    __syncthreads();
    const int t = tid + 1;
    if (t <= nlags && t < n) {
        double r = 0.0;
        for (int k = 0; k < kN; ++k) r = fma(Pd[k], ct[(k * t) & (kN - 1)], r);
        out[(long)(t - 1) * D + d] = (pre[n] - pre[t]) + pre[n - t] - 2.0 * r / kN;
    }
}

}  // namespace

#ifdef DIAG_FFT_HOST_TEST
// Host emulation of one warp's 1024-point transform (same functions, lanes run one after another): checks the index
// algebra of steps 1-4 against a direct DFT.  Build: nvcc -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a
// -DDIAG_FFT_HOST_TEST -o /tmp/fft_test diag_fft.cu
void hmc_set_error(const char*, ...) {}
#include <cstdio>
#include <cmath>
#include <vector>
#include <complex>
int main() {
    const int n = 400;
    std::vector<std::complex<double>> x(kN, 0.0);
    unsigned s = 12345u;
    for (int i = 0; i < n; ++i) {
        s = s * 1664525u + 1013904223u; double a = (double)(s >> 8) / (1 << 24) - 0.5;
        s = s * 1664525u + 1013904223u; double b = (double)(s >> 8) / (1 << 24) - 0.5;
        x[i] = {a, b};
    }
    static float T_re[32][33], T_im[32][33];
    for (int lane = 0; lane < 32; ++lane) {
        float re[32], im[32];
        for (int n1 = 0; n1 < 16; ++n1) { re[n1] = (float)x[32 * n1 + lane].real(); im[n1] = (float)x[32 * n1 + lane].imag(); }
        for (int n1 = 16; n1 < 32; ++n1) { re[n1] = 0.f; im[n1] = 0.f; }
        fft32<true>(re, im);
        for (int p = 0; p < 32; ++p) {
            const double ang = -2.0 * M_PI * ((lane * brev5(p)) & (kN - 1)) / kN;
            const float wx = (float)cos(ang), wy = (float)sin(ang);
            const float r2 = fmaf(re[p], wx, -im[p] * wy);
            im[p] = fmaf(re[p], wy, im[p] * wx);
            re[p] = r2;
            T_re[brev5(p)][lane] = re[p];
            T_im[brev5(p)][lane] = im[p];
        }
    }
    double maxerr = 0, maxmag = 0;
    std::vector<double> Pk(kN);
    for (int lane = 0; lane < 32; ++lane) {
        float re[32], im[32];
        for (int n2 = 0; n2 < 32; ++n2) { re[n2] = T_re[lane][n2]; im[n2] = T_im[lane][n2]; }
        fft32<false>(re, im);
        for (int p = 0; p < 32; ++p) {
            const int k = lane + 32 * brev5(p);
            std::complex<double> ref = 0.0;
            for (int i = 0; i < n; ++i) ref += x[i] * std::polar(1.0, -2.0 * M_PI * (double)((long)k * i % kN) / kN);
            maxerr = fmax(maxerr, std::abs(ref - std::complex<double>(re[p], im[p])));
            maxmag = fmax(maxmag, std::abs(ref));
            Pk[k] = (double)re[p] * re[p] + (double)im[p] * im[p];
        }
    }
    printf("max |X - ref| = %.3e (max |X| = %.3e)\n", maxerr, maxmag);
    // autocorrelation check at a few lags
    for (int t : {1, 7, 100, 399}) {
        double r = 0; for (int k = 0; k < kN; ++k) r += Pk[k] * cos(2.0 * M_PI * (double)((long)k * t % kN) / kN);
        r /= kN;
        double ref = 0; for (int i = 0; i + t < n; ++i) ref += x[i].real() * x[i + t].real() + x[i].imag() * x[i + t].imag();
        printf("lag %d: R = %.9f ref %.9f\n", t, r, ref);
    }
    return maxerr < 1e-4 * maxmag ? 0 : 1;
}
#endif

#define HMC_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            hmc_set_error(__VA_ARGS__);   \
            return HMC_E_BADARG;          \
        }                                 \
    } while (0)

extern "C" int64_t hmc_diag_variogram_all_workspace_bytes(int64_t n, int32_t D) {
    return (int64_t)sizeof(double) * ((int64_t)D * kN + 2 * n * (int64_t)D + kMaxGroups);
}

extern "C" int hmc_diag_variogram_all(int32_t dtype, const void* q, int64_t Nchain, int64_t n, int32_t D, int64_t stride_chain,
                                      int32_t nlags, double* out_nlags_x_D, void* workspace, int64_t workspace_bytes, void* cuda_stream) {
    HMC_REQUIRE(q && out_nlags_x_D && workspace, "NULL buffer");
    HMC_REQUIRE(Nchain >= 1 && n >= 2 && D >= 1, "need Nchain >= 1, n >= 2, D >= 1");
    HMC_REQUIRE(nlags >= 1 && nlags <= n - 1, "need 1 <= nlags <= n - 1");
    HMC_REQUIRE(stride_chain >= 2 * n * D, "stride_chain too small");
    if (!(dtype == HMC_F32 && n <= kHalf && (D % 4) == 0 && (stride_chain % 4) == 0 && (reinterpret_cast<uintptr_t>(q) % 16) == 0)) {
        hmc_set_error("hmc_diag_variogram_all covers float32 streams with n <= %d, D %% 4 == 0 and 16-byte aligned rows; use hmc_diag_variogram", kHalf);
        return HMC_E_UNSUPPORTED;
    }
    HMC_REQUIRE(workspace_bytes >= hmc_diag_variogram_all_workspace_bytes(n, D), "workspace too small");
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    double* P = (double*)workspace;
    double* Q = P + (size_t)D * kN;
    HMC_CUDA_CHECK(cudaMemsetAsync(workspace, 0, (size_t)hmc_diag_variogram_all_workspace_bytes(n, D), stream));
    int dev = 0, sms = 148;
    HMC_CUDA_CHECK(cudaGetDevice(&dev));
    HMC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int dtiles = D / 4;
    int groups = (3 * sms) / dtiles;                     // three resident blocks per SM, one wave
    if (groups < 1) groups = 1;
    if (groups > Nchain) groups = (int)Nchain;
    const size_t smem = sizeof(float2) * (32 * 32 + 4 * kZRow) + sizeof(float) * 4 * 32 * 33 + sizeof(float4) * 4 * (size_t)n;
    HMC_CUDA_CHECK(cudaFuncSetAttribute(diag_fft_power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (groups > kMaxGroups) groups = kMaxGroups;
    // cooperative launch when every block fits at once (then the sibling blocks may wait for each other); plain launch otherwise
    unsigned int* bars = reinterpret_cast<unsigned int*>(Q + 2 * (size_t)n * D);
    int coop = 0, per_sm = 0;
    HMC_CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    HMC_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, diag_fft_power_kernel, kThreads, smem));
    const float* qf = (const float*)q;
    long nch = Nchain;
    int ni = (int)n, gr = groups;
    bool launched = false;
    if (coop && (long)per_sm * sms >= (long)dtiles * groups && dtiles > 1 && getenv("HMC_B200_DIAG_FFT_FREE") == nullptr) {
        void* kargs[] = {&qf, &nch, &ni, &D, &stride_chain, &gr, &P, &Q, &bars};
        launched = cudaLaunchCooperativeKernel((const void*)diag_fft_power_kernel, dim3(dtiles * groups), dim3(kThreads), kargs, smem, stream) == cudaSuccess;
        if (!launched) (void)cudaGetLastError();
    }
    if (!launched) {
        unsigned int* none = nullptr;
        diag_fft_power_kernel<<<dtiles * groups, kThreads, smem, stream>>>(qf, nch, ni, D, stride_chain, gr, P, Q, none);
    }
    diag_fft_finish_kernel<<<D, kHalf, 0, stream>>>(P, Q, (int)n, D, nlags, out_nlags_x_D);
    HMC_CUDA_CHECK(cudaGetLastError());
    return HMC_OK;
}
