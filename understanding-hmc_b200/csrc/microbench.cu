// Design probes for the fused trajectory kernel (not part of the library): measures, on the box, the FP32 FMA
// peak and the achieved FMA rate of candidate inner loops (operands in shared memory, accumulators in
// registers) so that the register-tile shape is chosen from measurements, not guesses.
//   build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo microbench.cu -o ../../gpurun_out/microbench
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
    unsigned long long D = *reinterpret_cast<unsigned long long*>(&d);
    const unsigned long long A = *reinterpret_cast<const unsigned long long*>(&a);
    const unsigned long long B = *reinterpret_cast<const unsigned long long*>(&b);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(D) : "l"(A), "l"(B));
    d = *reinterpret_cast<float2*>(&D);
}

// ---- (0) raw FMA peak ------------------------------------------------------------------------------------
template <bool PACKED>
__global__ void __launch_bounds__(512) peak_kernel(float* out, int iters) {
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    float2 a = make_float2(0.999f, 0.998f), b = make_float2(0.001f, 0.002f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (PACKED) { float2 t = acc[i]; acc[i] = b; ffma2(acc[i], a, t); }
            else { acc[i].x = fmaf(a.x, acc[i].x, b.x); acc[i].y = fmaf(a.y, acc[i].y, b.y); }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- (A) warp tile: 6 chain-groups x 5 dim-groups per warp, thread tile 2 chains x 20 dims -------------------
// P [K][100] plain in smem (shared by all warps); per-warp q tile [K][6][4] = (q0,q0,q1,q1) duplicated.
template <bool PACKED, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) warp_tile_kernel(float* out, int steps) {
    constexpr int K = 100, D = 100;
    extern __shared__ __align__(16) float sm[];
    float* P = sm;                                   // K * D
    float* Q = sm + K * D + (threadIdx.x >> 5) * (K * 24);
    for (int t = threadIdx.x; t < K * D; t += blockDim.x) P[t] = 1e-3f * ((t * 7) % 13 - 6);
    const int lane = threadIdx.x & 31;
    for (int t = lane; t < K * 24; t += 32) Q[t] = 1e-2f * ((t * 5) % 11 - 5);
    __syncthreads();
    const int cg = lane / 5, dg = lane % 5;
    const bool active = lane < 30;
    float2 acc[2][10];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < 10; ++j) acc[c][j] = make_float2(0.f, 0.f);
    const float* qp = Q + (active ? cg : 0) * 4;
    const float* pp = P + (active ? dg : 0) * 20;
    for (int s = 0; s < steps; ++s) {
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            const float4 qv = *reinterpret_cast<const float4*>(qp + k * 24);
            float4 pv[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) pv[i] = *reinterpret_cast<const float4*>(pp + k * D + 4 * i);
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                if (PACKED) {
                    ffma2(acc[0][2 * i], make_float2(qv.x, qv.y), make_float2(pv[i].x, pv[i].y));
                    ffma2(acc[0][2 * i + 1], make_float2(qv.x, qv.y), make_float2(pv[i].z, pv[i].w));
                    ffma2(acc[1][2 * i], make_float2(qv.z, qv.w), make_float2(pv[i].x, pv[i].y));
                    ffma2(acc[1][2 * i + 1], make_float2(qv.z, qv.w), make_float2(pv[i].z, pv[i].w));
                } else {
                    acc[0][2 * i].x = fmaf(qv.x, pv[i].x, acc[0][2 * i].x); acc[0][2 * i].y = fmaf(qv.x, pv[i].y, acc[0][2 * i].y);
                    acc[0][2 * i + 1].x = fmaf(qv.x, pv[i].z, acc[0][2 * i + 1].x); acc[0][2 * i + 1].y = fmaf(qv.x, pv[i].w, acc[0][2 * i + 1].y);
                    acc[1][2 * i].x = fmaf(qv.z, pv[i].x, acc[1][2 * i].x); acc[1][2 * i].y = fmaf(qv.z, pv[i].y, acc[1][2 * i].y);
                    acc[1][2 * i + 1].x = fmaf(qv.z, pv[i].z, acc[1][2 * i + 1].x); acc[1][2 * i + 1].y = fmaf(qv.z, pv[i].w, acc[1][2 * i + 1].y);
                }
            }
        }
        // feed something back so the loop is not hoisted
        if (acc[0][0].x == 123.456f) Q[lane] = acc[1][3].y;
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < 10; ++j) s += acc[c][j].x + acc[c][j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- (B) warp tile variant: 3 chain-groups x 10 dim-groups, thread tile 4 chains x 10 dims --------------------
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) warp_tile_b_kernel(float* out, int steps) {
    constexpr int K = 100, D = 100;
    extern __shared__ __align__(16) float sm[];
    float* P = sm;
    float* Q = sm + K * D + (threadIdx.x >> 5) * (K * 24);   // [K][3 cg][8] = 4 chains duplicated
    for (int t = threadIdx.x; t < K * D; t += blockDim.x) P[t] = 1e-3f * ((t * 7) % 13 - 6);
    const int lane = threadIdx.x & 31;
    for (int t = lane; t < K * 24; t += 32) Q[t] = 1e-2f * ((t * 5) % 11 - 5);
    __syncthreads();
    const bool active = lane < 30;
    const int cg = active ? lane / 10 : 0, dg = active ? lane % 10 : 0;
    float2 acc[4][5];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int j = 0; j < 5; ++j) acc[c][j] = make_float2(0.f, 0.f);
    const float* qp = Q + cg * 8;
    const float* pp = P + dg * 10;
    for (int s = 0; s < steps; ++s) {
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            const float4 qa = *reinterpret_cast<const float4*>(qp + k * 24);
            const float4 qb = *reinterpret_cast<const float4*>(qp + k * 24 + 4);
            float2 pv[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) pv[i] = *reinterpret_cast<const float2*>(pp + k * D + 2 * i);
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                ffma2(acc[0][i], make_float2(qa.x, qa.y), pv[i]);
                ffma2(acc[1][i], make_float2(qa.z, qa.w), pv[i]);
                ffma2(acc[2][i], make_float2(qb.x, qb.y), pv[i]);
                ffma2(acc[3][i], make_float2(qb.z, qb.w), pv[i]);
            }
        }
        if (acc[0][0].x == 123.456f) Q[lane] = acc[1][3].y;
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int j = 0; j < 5; ++j) s += acc[c][j].x + acc[c][j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- (C) CTA tile: 8 chains x 4 dims per thread, pairs along chains, P duplicated in smem ---------------------
// q tile [K][2][NCG][4], Pdup [K][2][32 dg slots][4]; warp = 4 cg x 8 dg.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) cta_tile_kernel(float* out, int steps) {
    constexpr int K = 100, NCG = 20, NDGS = 32;
    extern __shared__ __align__(16) float sm[];
    float* Pd = sm;                          // K * 2 * NDGS * 4
    float* Q = sm + K * 2 * NDGS * 4;        // K * 2 * NCG * 4
    for (int t = threadIdx.x; t < K * 2 * NDGS * 4; t += blockDim.x) Pd[t] = 1e-3f * ((t * 7) % 13 - 6);
    for (int t = threadIdx.x; t < K * 2 * NCG * 4; t += blockDim.x) Q[t] = 1e-2f * ((t * 5) % 11 - 5);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cg = (warp % 5) * 4 + (lane >> 3);
    const int dg = (warp / 5) * 8 + (lane & 7);
    float2 acc[4][4];    // [chain pair][dim]
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[c][j] = make_float2(0.f, 0.f);
    const float* qp = Q + cg * 4;
    const float* pp = Pd + dg * 4;
    for (int s = 0; s < steps; ++s) {
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            const float4 q0 = *reinterpret_cast<const float4*>(qp + (k * 2 + 0) * NCG * 4);
            const float4 q1 = *reinterpret_cast<const float4*>(qp + (k * 2 + 1) * NCG * 4);
            const float4 p0 = *reinterpret_cast<const float4*>(pp + (k * 2 + 0) * NDGS * 4);
            const float4 p1 = *reinterpret_cast<const float4*>(pp + (k * 2 + 1) * NDGS * 4);
            const float2 qq[4] = {make_float2(q0.x, q0.y), make_float2(q0.z, q0.w), make_float2(q1.x, q1.y), make_float2(q1.z, q1.w)};
            const float2 pd[4] = {make_float2(p0.x, p0.y), make_float2(p0.z, p0.w), make_float2(p1.x, p1.y), make_float2(p1.z, p1.w)};
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j) ffma2(acc[c][j], qq[c], pd[j]);
        }
        if (acc[0][0].x == 123.456f) Q[lane] = acc[1][3].y;
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += acc[c][j].x + acc[c][j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// ---- (D) lane-owns-chains: thread = 2 chains (one FFMA2 pair) x all dims in slices of TN; P from the constant
// bank through uniform registers (LDCU + FFMA2 with a scalar UR operand); q tile per warp [K][66] in smem.
__constant__ float Pc[100 * 100];
template <int TN, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) lane_chain_kernel(float* out, int steps) {
    constexpr int K = 100, D = 100, QS = 66;
    extern __shared__ __align__(16) float sm[];
    float* Q = sm + (threadIdx.x >> 5) * (K * QS);
    const int lane = threadIdx.x & 31;
    for (int t = lane; t < K * QS; t += 32) Q[t] = 1e-2f * ((t * 5) % 11 - 5);
    __syncwarp();
    float2 tot = make_float2(0.f, 0.f);
    const float* qp = Q + 2 * lane;
    for (int s = 0; s < steps; ++s) {
#pragma unroll 1
        for (int sl = 0; sl < D / TN; ++sl) {
            float2 acc[TN];
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll 4
            for (int k = 0; k < K; ++k) {
                const float2 qv = *reinterpret_cast<const float2*>(qp + k * QS);
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    const float pj = Pc[k * D + sl * TN + j];
                    ffma2(acc[j], qv, make_float2(pj, pj));
                }
            }
#pragma unroll
            for (int j = 0; j < TN; ++j) { tot.x += acc[j].x; tot.y += acc[j].y; }
        }
        if (tot.x == 123.456f) Q[lane] = tot.y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = tot.x + tot.y;
}


// ---- (E) generic per-warp register tile: thread = TM chains (FFMA2 pairs along chains) x TN dims, the P value
// enters FFMA2 as a scalar-broadcast operand (no duplication); warp = NCGW chain-groups x NDGW dim-groups.
template <int TM, int TN, int NCGW, int NDGW, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) warp_generic_kernel(float* out, int steps) {
    constexpr int K = 100, QS = NCGW * TM, PS = NDGW * TN;
    static_assert(TM % 2 == 0 && NCGW * NDGW <= 32, "tile");
    extern __shared__ __align__(16) float sm[];
    float* P = sm;                                   // K * PS
    float* Q = sm + K * PS + (threadIdx.x >> 5) * (K * QS);
    for (int t = threadIdx.x; t < K * PS; t += blockDim.x) P[t] = 1e-3f * ((t * 7) % 13 - 6);
    const int lane = threadIdx.x & 31;
    for (int t = lane; t < K * QS; t += 32) Q[t] = 1e-2f * ((t * 5) % 11 - 5);
    __syncthreads();
    const bool active = lane < NCGW * NDGW;
    const int cg = active ? lane / NDGW : 0, dg = active ? lane % NDGW : 0;
    float2 acc[TM / 2][TN];
#pragma unroll
    for (int c = 0; c < TM / 2; ++c)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[c][j] = make_float2(0.f, 0.f);
    const float* qp = Q + cg * TM;
    const float* pp = P + dg * TN;
    for (int s = 0; s < steps; ++s) {
#pragma unroll 2
        for (int k = 0; k < K; ++k) {
            float qv[TM], pv[TN];
            if constexpr (TM % 4 == 0) {
#pragma unroll
                for (int i = 0; i < TM / 4; ++i) *reinterpret_cast<float4*>(&qv[4 * i]) = *reinterpret_cast<const float4*>(qp + k * QS + 4 * i);
            } else {
#pragma unroll
                for (int i = 0; i < TM / 2; ++i) *reinterpret_cast<float2*>(&qv[2 * i]) = *reinterpret_cast<const float2*>(qp + k * QS + 2 * i);
            }
            if constexpr (TN % 4 == 0) {
#pragma unroll
                for (int i = 0; i < TN / 4; ++i) *reinterpret_cast<float4*>(&pv[4 * i]) = *reinterpret_cast<const float4*>(pp + k * PS + 4 * i);
            } else {
#pragma unroll
                for (int i = 0; i < TN / 2; ++i) *reinterpret_cast<float2*>(&pv[2 * i]) = *reinterpret_cast<const float2*>(pp + k * PS + 2 * i);
            }
#pragma unroll
            for (int c = 0; c < TM / 2; ++c)
#pragma unroll
                for (int j = 0; j < TN; ++j) ffma2(acc[c][j], make_float2(qv[2 * c], qv[2 * c + 1]), make_float2(pv[j], pv[j]));
        }
        if (acc[0][0].x == 123.456f) Q[lane] = acc[0][1].y;
    }
    float r = 0.f;
#pragma unroll
    for (int c = 0; c < TM / 2; ++c)
#pragma unroll
        for (int j = 0; j < TN; ++j) r += acc[c][j].x + acc[c][j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}


// ---- (F) like (E) with TM=8, TN=10 but explicit register double buffering of the operands (loads of step k+1
// issued before the FMAs of step k): how efficient is the loop with ONE warp per scheduler?
template <int WARPS, int UNR>
__global__ void __launch_bounds__(WARPS * 32) warp_db_kernel(float* out, int steps) {
    constexpr int TM = 8, TN = 10, NCGW = 3, NDGW = 10, K = 100, QS = 28, PS = NDGW * TN;
    extern __shared__ __align__(16) float sm[];
    float* P = sm;
    float* Q = sm + K * PS + (threadIdx.x >> 5) * (K * QS);
    for (int t = threadIdx.x; t < K * PS; t += blockDim.x) P[t] = 1e-3f * ((t * 7) % 13 - 6);
    const int lane = threadIdx.x & 31;
    for (int t = lane; t < K * QS; t += 32) Q[t] = 1e-2f * ((t * 5) % 11 - 5);
    __syncthreads();
    const bool active = lane < NCGW * NDGW;
    const int cg = active ? lane / NDGW : 0, dg = active ? lane % NDGW : 0;
    float2 acc[TM / 2][TN];
#pragma unroll
    for (int c = 0; c < TM / 2; ++c)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[c][j] = make_float2(0.f, 0.f);
    const float* qp = Q + cg * TM;
    const float* pp = P + dg * TN;
    auto load = [&](int k, float (&qv)[TM], float (&pv)[TN]) {
#pragma unroll
        for (int i = 0; i < TM / 4; ++i) *reinterpret_cast<float4*>(&qv[4 * i]) = *reinterpret_cast<const float4*>(qp + k * QS + 4 * i);
#pragma unroll
        for (int i = 0; i < TN / 2; ++i) *reinterpret_cast<float2*>(&pv[2 * i]) = *reinterpret_cast<const float2*>(pp + k * PS + 2 * i);
    };
    auto fmas = [&](const float (&qv)[TM], const float (&pv)[TN]) {
#pragma unroll
        for (int c = 0; c < TM / 2; ++c)
#pragma unroll
            for (int j = 0; j < TN; ++j) ffma2(acc[c][j], make_float2(qv[2 * c], qv[2 * c + 1]), make_float2(pv[j], pv[j]));
    };
    for (int s = 0; s < steps; ++s) {
        float qa[TM], pa[TN], qb[TM], pb[TN];
        load(0, qa, pa);
#pragma unroll UNR
        for (int k = 0; k < K; k += 2) {
            load(k + 1, qb, pb);
            fmas(qa, pa);
            load((k + 2 < K) ? k + 2 : 0, qa, pa);
            fmas(qb, pb);
        }
        if (acc[0][0].x == 123.456f) Q[lane] = acc[0][1].y;
    }
    float r = 0.f;
#pragma unroll
    for (int c = 0; c < TM / 2; ++c)
#pragma unroll
        for (int j = 0; j < TN; ++j) r += acc[c][j].x + acc[c][j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F>
double time_ms(F launch, int reps = 3) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    int clk = 0;
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("SMs %d, max clock %d kHz, nominal FP32 peak %.2f TFLOP/s\n", sms, clk, 2.0 * 128 * sms * clk * 1e3 / 1e12);
    float* out;
    CK(cudaMalloc(&out, sizeof(float) * sms * 8 * 1024));
    {
        const int iters = 20000, blocks = sms * 4, threads = 512;
        double ms = time_ms([&] { peak_kernel<false><<<blocks, threads>>>(out, iters); });
        printf("peak FFMA   : %.2f TFLOP/s\n", 2.0 * 32 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { peak_kernel<true><<<blocks, threads>>>(out, iters); });
        printf("peak FFMA2  : %.2f TFLOP/s\n", 2.0 * 32 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12);
    }
    const int steps = 2000;
    {
        constexpr int W = 16;
        const size_t smem = sizeof(float) * (100 * 100 + W * 100 * 24);
        auto k1 = warp_tile_kernel<true, W>;
        auto k2 = warp_tile_kernel<false, W>;
        CK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const double flop = 2.0 * 100 * 100 * 12.0 * W * sms * steps;   // 12 chains per warp
        double ms = time_ms([&] { k1<<<sms, W * 32, smem>>>(out, steps); });
        printf("(A) warp tile 6x5 (2x20) FFMA2, %d warps: %.2f TFLOP/s useful\n", W, flop / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { k2<<<sms, W * 32, smem>>>(out, steps); });
        printf("(A) warp tile 6x5 (2x20) FFMA , %d warps: %.2f TFLOP/s useful\n", W, flop / (ms * 1e-3) / 1e12);
    }
    {
        constexpr int W = 12;
        const size_t smem = sizeof(float) * (100 * 100 + W * 100 * 24);
        auto k1 = warp_tile_kernel<true, W>;
        CK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const double flop = 2.0 * 100 * 100 * 12.0 * W * sms * steps;
        double ms = time_ms([&] { k1<<<sms, W * 32, smem>>>(out, steps); });
        printf("(A) warp tile 6x5 (2x20) FFMA2, %d warps: %.2f TFLOP/s useful\n", W, flop / (ms * 1e-3) / 1e12);
    }
    {
        constexpr int W = 16;
        const size_t smem = sizeof(float) * (100 * 100 + W * 100 * 24);
        auto k1 = warp_tile_b_kernel<W>;
        CK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const double flop = 2.0 * 100 * 100 * 12.0 * W * sms * steps;
        double ms = time_ms([&] { k1<<<sms, W * 32, smem>>>(out, steps); });
        printf("(B) warp tile 3x10 (4x10) FFMA2, %d warps: %.2f TFLOP/s useful\n", W, flop / (ms * 1e-3) / 1e12);
    }
    {
        constexpr int W = 15;   // 5 cg-blocks x 3 dg-blocks (96 of 100 dims; the probe ignores the last dim group)
        const size_t smem = sizeof(float) * (100 * 2 * 32 * 4 + 100 * 2 * 20 * 4);
        auto k1 = cta_tile_kernel<W>;
        CK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const double flop = 2.0 * 100 * 96 * 160.0 * sms * steps;
        double ms = time_ms([&] { k1<<<sms, W * 32, smem>>>(out, steps); });
        printf("(C) CTA tile 8x4 Pdup FFMA2, %d warps: %.2f TFLOP/s useful\n", W, flop / (ms * 1e-3) / 1e12);
    }

    {
        static float hP[100 * 100];
        for (int i = 0; i < 100 * 100; ++i) hP[i] = 1e-3f * ((i * 7) % 13 - 6);
        CK(cudaMemcpyToSymbol(Pc, hP, sizeof(hP)));
        const double flop_per_warp = 2.0 * 100 * 100 * 64.0 * steps;
#define RUN_D(TN, W, CTAS)                                                                                   \
        {                                                                                                    \
            const size_t smem = sizeof(float) * W * 100 * 66;                                                \
            auto kd = lane_chain_kernel<TN, W>;                                                              \
            CK(cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
            double ms = time_ms([&] { kd<<<sms * CTAS, W * 32, smem>>>(out, steps); });                      \
            printf("(D) lane-owns-2-chains, slice %d dims, %d warps x %d CTA/SM: %.2f TFLOP/s useful\n", TN, W, CTAS, \
                   flop_per_warp * W * CTAS * sms / (ms * 1e-3) / 1e12);                                     \
        }
        RUN_D(25, 4, 1)
        RUN_D(20, 4, 1)
        RUN_D(50, 4, 1)
        RUN_D(25, 1, 4)
        RUN_D(25, 2, 2)
        RUN_D(25, 8, 1)
        RUN_D(10, 8, 1)
    }

    {
#define RUN_E(TM, TN, NCGW, NDGW, W)                                                                          \
        {                                                                                                    \
            const size_t smem = sizeof(float) * (100 * NDGW * TN + W * 100 * NCGW * TM);                     \
            auto ke = warp_generic_kernel<TM, TN, NCGW, NDGW, W>;                                            \
            CK(cudaFuncSetAttribute(ke, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
            double ms = time_ms([&] { ke<<<sms, W * 32, smem>>>(out, steps); });                             \
            printf("(E) TM=%d TN=%d warp %dx%d, %d warps (smem %zu KB): %.2f TFLOP/s useful\n", TM, TN, NCGW, NDGW, W, smem / 1024, \
                   2.0 * 100 * (NDGW * TN) * (double)(NCGW * TM) * W * sms * steps / (ms * 1e-3) / 1e12);   \
        }
        RUN_E(8, 10, 3, 10, 8)
        RUN_E(8, 10, 3, 10, 12)
        RUN_E(6, 10, 3, 10, 12)
        RUN_E(4, 20, 6, 5, 12)
        RUN_E(4, 10, 3, 10, 16)
        RUN_E(8, 4, 4, 8, 16)
        RUN_E(8, 8, 4, 8, 8)
        RUN_E(8, 8, 4, 8, 12)
        RUN_E(8, 12, 4, 8, 8)
    }

    {
        RUN_E(8, 10, 3, 10, 4)
#define RUN_F(W, UNR)                                                                                          \
        {                                                                                                    \
            const size_t smem = sizeof(float) * (100 * 100 + W * 100 * 28);                                  \
            auto kf = warp_db_kernel<W, UNR>;                                                                \
            CK(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
            double ms = time_ms([&] { kf<<<sms, W * 32, smem>>>(out, steps); });                             \
            printf("(F) TM=8 TN=10 double-buffered, %d warps, unroll %d: %.2f TFLOP/s useful\n", W, UNR,      \
                   2.0 * 100 * 100 * 24.0 * W * sms * steps / (ms * 1e-3) / 1e12);                           \
        }
        RUN_F(4, 1)
        RUN_F(4, 2)
        RUN_F(8, 1)
        RUN_F(8, 2)
    }
    CK(cudaFree(out));
    return 0;
}
