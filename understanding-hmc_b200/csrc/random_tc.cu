// Tensor-core random-trajectory HMC kernel (tcgen05 + TMEM), FP32-grade arithmetic through a bf16x3 split, D = 100.
//
// Follows HMC_sampler.gen_sample_random + leap_frog (/root/reference/samplers.py:387-491, 831-839).
//
// One CTA = 128 chains, 512 threads.  The gradient of all 128 chains,
//       G[128 x N] = Dm[128 x K] * F[N x K]^T          (K = N = 112: D = 100 zero-padded to a multiple of 16)
// runs on the 5th-generation tensor cores: the shifted positions d = q - mu are kept in shared memory as three bf16
// parts (d = d1 + d2 + d3 exactly), the force matrix is split once the same way, and six tcgen05.mma passes
// (1,3) (3,1) (2,2) (1,2) (2,1) (1,1) accumulate in fp32 in tensor memory -- the dropped terms are O(2^-24).
// Shared-memory operands use the canonical no-swizzle K-major layout: 16-byte chunk (kc, row) at (kc*rows + row)*16
// (LBO = rows*16 between K chunks, SBO = 128 between 8-row groups), so per-thread row writes are conflict-free
// 128-bit stores.
//
// The accumulator row of chain c is TMEM lane c.  FOUR threads share a chain: thread (warp w, lane) works on chain
// 32*(w%4) + lane (the TMEM lanes a warp may read) and on the dimension slice w/4 (24, 24, 24, 28 dims): it reads
// its slice of the gradient with tcgen05.ld, keeps the momentum slice in registers, updates the position slice
// (fp32 copy in shared memory) and re-splits it.  Per-chain sums (d.g, p.p) are combined through shared memory by
// the slice-0 thread, which does the chain's bookkeeping (energies, Metropolis accept on a Philox uniform, new
// trajectory length) and posts a command that all four slice threads apply (sample store, restore, momentum
// refresh).  Every chain advances one gradient evaluation per pass; iteration boundaries are per-chain events
// (SURVEY H3); the first point of each trajectory is a gradient-only pass, so an iteration costs L + 1
// evaluations.  Momentum refresh is warp-cooperative (one Philox call per lane, same draws as every other kernel);
// finished chains pull the next chain from a global queue.
#include "hmc_common.cuh"
#include <cuda_bf16.h>

#ifdef HMC_PROFILE_PHASES
__device__ unsigned long long g_tc_cycles[8];
#define TP_T(x) const long long x = clock64()
#define TP_ADD(i, a, b) tph[i] += (b) - (a)
#else
#define TP_T(x)
#define TP_ADD(i, a, b)
#endif

namespace {

constexpr int TC_ND = 100;          // dimensions handled by this instantiation
constexpr int TC_KP = 112;          // padded K = N (multiple of 16)
constexpr int TC_KC = TC_KP / 8;    // 16-byte chunks (8 bf16) per operand row
constexpr int TC_M = 128;           // chains per CTA
constexpr int TC_SPL = 4;           // threads (dimension slices) per chain
constexpr int TC_THREADS = TC_M * TC_SPL;
constexpr int TC_APART = TC_KC * TC_M * 16;     // bytes of one A part
constexpr int TC_BPART = TC_KC * TC_KP * 16;    // bytes of one B part
constexpr int TC_DCH = TC_ND / 4;               // fp32 position chunks (4 dims) per chain
constexpr int TC_SLOTS = 5;                     // momentum staging slots per 32-chain group and round

enum : int { CMD_STORE_Q0 = 1, CMD_STORE_OUT = 2, CMD_RESTORE = 4, CMD_NEW = 8, CMD_PARK = 16, CMD_REFRESH = 32, CMD_INIT0 = 64 };
enum : int { MODE_IDLE = 0, MODE_FIRST = 1, MODE_MID = 2, MODE_LAST = 3 };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version 1 (Blackwell); no swizzle, base offset 0
    return d;
}

// x = b1 + b2 + b3 exactly (three bf16 parts); the parts of two values packed as bf16x2 (lo = x0, hi = x1)
__device__ __forceinline__ void split3(float x0, float x1, uint32_t& h1, uint32_t& h2, uint32_t& h3) {
    __nv_bfloat162 a = __floats2bfloat162_rn(x0, x1);
    h1 = *reinterpret_cast<uint32_t*>(&a);
    float r0 = x0 - __uint_as_float(h1 << 16), r1 = x1 - __uint_as_float(h1 & 0xffff0000u);
    __nv_bfloat162 b = __floats2bfloat162_rn(r0, r1);
    h2 = *reinterpret_cast<uint32_t*>(&b);
    r0 -= __uint_as_float(h2 << 16); r1 -= __uint_as_float(h2 & 0xffff0000u);
    __nv_bfloat162 c = __floats2bfloat162_rn(r0, r1);
    h3 = *reinterpret_cast<uint32_t*>(&c);
}

struct TcGen {
    uint64_t seed;
    const double* p_tape;
    const int32_t* L_tape;
    const double* u_tape;
    int Niter, L_low, L_high;
};

// Warp-cooperative momentum draw for one chain: lane sl < 25 draws the normals of dims 4*sl..4*sl+3 into `stage`,
// lane 25 the scalars of the iteration.  Same arithmetic as hmc_normal4 / hmc_scalar_draws.
__device__ __noinline__ void tc_gen(const TcGen g, long m, uint64_t gid, int iter, int lane, float* stage, float* sumsq,
                                    int* L, float* lnu) {
    constexpr int D = TC_ND, nslot = TC_ND / 4;
    float s = 0.f;
    int Lv = 1;
    float lv = 0.f;
    if (g.p_tape) {
        const double* src = g.p_tape + ((size_t)m * (g.Niter + 1) + iter) * D;
        for (int j = lane; j < D; j += 32) { const float v = (float)src[j]; stage[j] = v; s = fmaf(v, v, s); }
        if (iter >= 1) { Lv = g.L_tape[(size_t)m * g.Niter + iter - 1]; lv = (float)log(g.u_tape[(size_t)m * g.Niter + iter - 1]); }
    } else {
        const bool scalar_lane = (lane == nslot);
        const uint32_t hi = (uint32_t)(gid >> 32) << 8;
        const Philox4 r = philox4x32_10((uint32_t)gid, (uint32_t)iter, scalar_lane ? 0u : (uint32_t)lane,
                                        (scalar_lane ? (uint32_t)HMC_STREAM_SCALAR : (uint32_t)HMC_STREAM_MOMENTUM) | hi,
                                        (uint32_t)g.seed, (uint32_t)(g.seed >> 32));
        const float u1 = ((float)(r.x >> 8) + 0.5f) * 5.9604644775390625e-08f;
        const float u2 = ((float)(r.z >> 8) + 0.5f) * 5.9604644775390625e-08f;
        const float r1 = sqrtf(-2.0f * __logf(u1));
        const float r2 = sqrtf(-2.0f * __logf(u2));
        const float a1 = ((float)(r.y >> 8) * 5.9604644775390625e-08f - 0.5f) * 6.283185307179586f;
        const float a2 = ((float)(r.w >> 8) * 5.9604644775390625e-08f - 0.5f) * 6.283185307179586f;
        const float4 z = make_float4(r1 * __cosf(a1), r1 * __sinf(a1), r2 * __cosf(a2), r2 * __sinf(a2));
        if (lane < nslot) {
            *reinterpret_cast<float4*>(stage + 4 * lane) = z;
            s = z.x * z.x + z.y * z.y + z.z * z.z + z.w * z.w;
        }
        const int Ls = g.L_low + (int)__umulhi(r.x, (uint32_t)(g.L_high - g.L_low));
        const float ls = logf(((float)(r.y >> 8) + 0.5f) * 5.9604644775390625e-08f);
        Lv = __shfl_sync(HMC_FULL_MASK, Ls, nslot);
        lv = __shfl_sync(HMC_FULL_MASK, ls, nslot);
    }
    *sumsq = warp_sum<float>(s);
    *L = Lv;
    *lnu = lv;
    __syncwarp();
}

struct TcShared {                       // small per-chain arrays in shared memory
    float2 red[TC_SPL][TC_M];           // partial (d.g, p.p) per slice
    int mode[TC_M];                     // MODE_* of the gradient being evaluated
    int cmd[TC_M];                      // CMD_* flags posted by the bookkeeping thread
    int cm[TC_M];                       // local chain index the command refers to (row addressing)
    int cidx[TC_M];                     // stored-sample index of the command
    int cit[TC_M];                      // iteration whose momentum is to be drawn
    float gK[TC_M], gK0[TC_M], glnu[TC_M];   // results of the momentum draw
    int gL[TC_M];
};

__global__ void __launch_bounds__(TC_THREADS, 1) hmc_random_tc_kernel(const hmc_random_args a, unsigned int* __restrict__ queue) {
    constexpr int D = TC_ND, KP = TC_KP, KC = TC_KC;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* Ap = smem;                                   // 3 parts [KC][128] 16-byte chunks
    unsigned char* Bp = Ap + 3 * TC_APART;                      // 3 parts [KC][112]
    float4* Df = reinterpret_cast<float4*>(Bp + 3 * TC_BPART);  // fp32 positions [TC_DCH][128] 16-byte chunks
    float* mu_s = reinterpret_cast<float*>(Df + TC_DCH * TC_M); // [KP]
    float* dt_s = mu_s + KP;                                    // [KP]
    float* stage_all = dt_s + KP;                               // [4 groups][TC_SLOTS][128] momentum staging
    TcShared* sh = reinterpret_cast<TcShared*>(stage_all + 4 * TC_SLOTS * 128);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sh + 1);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = warp & 3;                    // 32-chain group == TMEM lane quarter this warp may access
    const int slice = warp >> 2;                 // dimension slice
    const int chain = grp * 32 + lane;           // chain slot of this thread
    const int j0 = 24 * slice;                   // first dimension of the slice
    const bool wide = slice == TC_SPL - 1;       // the last slice has 28 dims, the others 24

    // ---- one-time set-up ---------------------------------------------------------------------------------------------
    for (int t = tid; t < 3 * TC_APART / 16; t += TC_THREADS) reinterpret_cast<uint4*>(Ap)[t] = make_uint4(0u, 0u, 0u, 0u);
    for (int t = tid; t < TC_DCH * TC_M; t += TC_THREADS) Df[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    {
        const float* Ft = (const float*)a.target.Ft;            // Ft[k][n] = F[n][k]; B row n holds F[n][.] (K-major)
        const int Dpad = a.target.D_pad;
        for (int t = tid; t < KC * KP; t += TC_THREADS) {       // one 16-byte chunk (8 k values) of row n per item
            const int kc = t / KP, n = t % KP;
            uint32_t w1[4], w2[4], w3[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k0 = kc * 8 + 2 * e, k1 = k0 + 1;
                const float x0 = (n < D && k0 < D) ? Ft[(size_t)k0 * Dpad + n] : 0.f;
                const float x1 = (n < D && k1 < D) ? Ft[(size_t)k1 * Dpad + n] : 0.f;
                split3(x0, x1, w1[e], w2[e], w3[e]);
            }
            reinterpret_cast<uint4*>(Bp)[t] = make_uint4(w1[0], w1[1], w1[2], w1[3]);
            reinterpret_cast<uint4*>(Bp + TC_BPART)[t] = make_uint4(w2[0], w2[1], w2[2], w2[3]);
            reinterpret_cast<uint4*>(Bp + 2 * TC_BPART)[t] = make_uint4(w3[0], w3[1], w3[2], w3[3]);
        }
        for (int t = tid; t < KP; t += TC_THREADS) {
            mu_s[t] = (t < D) ? ((const float*)a.target.mu)[t] : 0.f;
            dt_s[t] = (t < D) ? ((const float*)a.target.dt)[t] : 0.f;
        }
        for (int t = tid; t < TC_M; t += TC_THREADS) { sh->mode[t] = MODE_IDLE; sh->cmd[t] = 0; }
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmem_row = tmem + ((uint32_t)(grp * 32) << 16) + (uint32_t)j0;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KP >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

    const long Lc = 1 + (a.Niter - a.warm_up_num) / a.thin_rate;      // samplers.py:31
    float* q_chain = (float*)a.q_chain;
    float* q0g = (float*)a.state_q;
    const double vconst = a.target.v_const;
    TcGen ga;
    ga.seed = a.seed; ga.p_tape = a.p_tape; ga.L_tape = a.L_tape; ga.u_tape = a.u_tape;
    ga.Niter = a.Niter; ga.L_low = a.L_low; ga.L_high = a.L_high;

    // ---- per-thread state -----------------------------------------------------------------------------------------------
    float p[28];                     // momentum slice (registers)
#pragma unroll
    for (int j = 0; j < 28; ++j) p[j] = 0.f;
    // bookkeeping state of chain `chain`, used by the slice-0 thread only
    long m = -1;                     // local chain index, -1 = no chain
    int it = 0, l = 0, L = 1;        // iteration, point index of the next gradient, trajectory length
    bool init = false;               // chain start: E_chain[.,0] still to be recorded
    bool want = (slice == 0);        // needs a (new) chain
    bool fetch = false;              // momentum of iteration `it` was requested, results are in sh->g*
    double E_init = 0.0, E_prev = 0.0;
    float K0 = 0.f, Knew = 0.f, lnu = 0.f;
    unsigned int n_acc_warm = 0, n_acc_post = 0, n_sumL = 0, n_sumL2 = 0;
    uint32_t phase = 0;
    bool have_grad = false;          // a gradient pass has been issued and its accumulator is to be consumed
#ifdef HMC_PROFILE_PHASES
    long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif

    // position slice helpers -------------------------------------------------------------------------------------------------
    // write the slice [j0, j0+nj) of the fp32 row and re-split it into the three bf16 parts (8-dim operand chunks)
    auto store_slice = [&](const float (&x)[32], int nch4) {
#pragma unroll
        for (int c = 0; c < 7; ++c)
            if (c < nch4) Df[(j0 / 4 + c) * TC_M + chain] = make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
            if (kc < 3 || wide) {
                uint32_t w1[4], w2[4], w3[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split3(x[8 * kc + 2 * e], x[8 * kc + 2 * e + 1], w1[e], w2[e], w3[e]);
                const int off = ((j0 / 8 + kc) * TC_M + chain) * 16;
                *reinterpret_cast<uint4*>(Ap + off) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
                *reinterpret_cast<uint4*>(Ap + TC_APART + off) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
                *reinterpret_cast<uint4*>(Ap + 2 * TC_APART + off) = make_uint4(w3[0], w3[1], w3[2], w3[3]);
            }
        }
    };

    while (true) {
        TP_T(t0);
        // ===== P1. consume the gradient: thread-local leapfrog update of the slice (samplers.py:835-837) ==============
        if (have_grad) {
            {
                uint32_t done = 0;
                while (!done) {
                    asm volatile("{\n\t.reg .pred pw;\n\tmbarrier.try_wait.parity.shared::cta.b64 pw, [%1], %2;\n\tselp.u32 %0, 1, 0, pw;\n\t}"
                                 : "=r"(done) : "r"(smem_u32(mbar)), "r"(phase) : "memory");
                }
                phase ^= 1u;
            }
            asm volatile("tcgen05.fence::after_thread_sync;");
            TP_T(t1);
            TP_ADD(0, t0, t1);
            const int md = sh->mode[chain];
            // point index of the gradient: first = half kick + drift, last = half kick only, interior = second half
            // kick of step l + first half kick of step l+1, drift
            const float kwt = (md == MODE_IDLE) ? 0.f : (md == MODE_MID ? -1.0f : -0.5f);
            const float dwt = (md == MODE_FIRST || md == MODE_MID) ? 1.f : 0.f;
            uint32_t gv[32];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(gv[0]), "=r"(gv[1]), "=r"(gv[2]), "=r"(gv[3]), "=r"(gv[4]), "=r"(gv[5]), "=r"(gv[6]), "=r"(gv[7]),
                           "=r"(gv[8]), "=r"(gv[9]), "=r"(gv[10]), "=r"(gv[11]), "=r"(gv[12]), "=r"(gv[13]), "=r"(gv[14]), "=r"(gv[15])
                         : "r"(tmem_row));
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(gv[16]), "=r"(gv[17]), "=r"(gv[18]), "=r"(gv[19]), "=r"(gv[20]), "=r"(gv[21]), "=r"(gv[22]), "=r"(gv[23]),
                           "=r"(gv[24]), "=r"(gv[25]), "=r"(gv[26]), "=r"(gv[27]), "=r"(gv[28]), "=r"(gv[29]), "=r"(gv[30]), "=r"(gv[31])
                         : "r"(tmem_row + 16u));
            asm volatile("tcgen05.wait::ld.sync.aligned;");
            asm volatile("tcgen05.fence::before_thread_sync;");      // TMEM reads ordered before the next MMA
            float x[32];
            float hv = 0.f, hk = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const bool on = c < 6 || (wide && c < 7);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (on) v = Df[(j0 / 4 + c) * TC_M + chain];
                x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
            }
            if (slice == 0 && md == MODE_FIRST && a.phi_q && m >= 0 && a.chain_id0 + m == 0 && it <= a.N_save_chain0) {
                double* phi = a.phi_q + (size_t)(it - 1) * a.L_high * 2;      // row 0: the start point (samplers.py:445)
                phi[0] = (double)(x[0] + mu_s[0]); phi[1] = (double)(x[1] + mu_s[1]);
            }
#pragma unroll
            for (int jj = 0; jj < 28; ++jj) {
                if (jj < 24 || wide) {
                    const float gj = __uint_as_float(gv[jj]);
                    const float dtj = dt_s[j0 + jj];
                    hv = fmaf(x[jj], gj, hv);
                    const float pn = fmaf(gj, kwt * dtj, p[jj]);
                    hk = fmaf(pn, pn, hk);
                    p[jj] = pn;
                    x[jj] = fmaf(pn, dwt * dtj, x[jj]);
                }
            }
            store_slice(x, wide ? 7 : 6);
            sh->red[slice][chain] = make_float2(hv, hk);
            if (slice == 0 && a.phi_q && m >= 0 && a.chain_id0 + m == 0 && it <= a.N_save_chain0 && (md == MODE_FIRST || md == MODE_MID)) {
                // chain-0 trajectory capture (samplers.py:442-452): the point reached by this drift is row l+1
                double* phi = a.phi_q + (size_t)(it - 1) * a.L_high * 2;
                const int row = (md == MODE_FIRST) ? 1 : l + 1;
                phi[2 * row] = (double)(x[0] + mu_s[0]); phi[2 * row + 1] = (double)(x[1] + mu_s[1]);
            }
            TP_T(t2);
            TP_ADD(1, t1, t2);
        }
        __syncthreads();
        TP_T(t3);

        // ===== P2. per-chain bookkeeping by the slice-0 thread ==========================================================
        if (slice == 0) {
            int cmd = 0;
            if (fetch && have_grad) {                      // the momentum requested in the previous pass has been drawn
                Knew = 0.5f * sh->gK[chain]; L = sh->gL[chain]; lnu = sh->glnu[chain];
                if (init) K0 = 0.5f * sh->gK0[chain];
                n_sumL += (unsigned int)L; n_sumL2 += (unsigned int)(L * L);
                l = 0;
                fetch = false;
                if (a.phi_q && a.chain_id0 + m == 0 && it <= a.N_save_chain0) a.phi_len[it - 1] = L + 1;   // samplers.py:444
            }
            if (have_grad && m >= 0 && sh->mode[chain] != MODE_IDLE) {
                const int md = sh->mode[chain];
                float sv = 0.f, sk = 0.f;
#pragma unroll
                for (int s2 = 0; s2 < TC_SPL; ++s2) { const float2 r = sh->red[s2][chain]; sv += r.x; sk += r.y; }
                const bool tr = a.phi_q && (a.chain_id0 + m) == 0 && it <= a.N_save_chain0;
                const double V = 0.5 * (double)sv + vconst;                     // V(q) = 0.5 d.P d + const (utils.py:213-218)
                if (md == MODE_FIRST) {
                    if (init) {                                                 // samplers.py:416-420
                        const double E0 = V + (double)K0;
                        a.E_chain[(size_t)m * Lc] = E0;
                        a.dE_chain[(size_t)m * Lc] = 0.0;
                        E_prev = E0;
                        init = false;
                    }
                    E_init = V + (double)Knew;                                  // samplers.py:434-438
                    if (it >= a.warm_up_num) {
                        const long idx = (it - a.warm_up_num) / a.thin_rate;
                        a.E_chain[(size_t)m * Lc + idx] = E_init;
                        a.dE_chain[(size_t)m * Lc + idx] = E_init - E_prev;
                    }
                    l = 1;
                    sh->mode[chain] = (l == L) ? MODE_LAST : MODE_MID;
                } else if (md == MODE_LAST) {
                    // Metropolis accept (samplers.py:455-472)
                    const double E_final = V + 0.5 * (double)sk;
                    const double dE = E_final - E_init;
                    E_prev = E_init;                                            // samplers.py:460
                    const bool accepted = (dE < 0) || ((double)lnu < -dE);      // samplers.py:462
                    const bool keep = it >= a.warm_up_num;
                    if (accepted) { if (keep) n_acc_post++; else n_acc_warm++; cmd |= CMD_STORE_Q0; }
                    else cmd |= CMD_RESTORE;
                    if (keep) cmd |= CMD_STORE_OUT;
                    sh->cm[chain] = (int)m;
                    sh->cidx[chain] = keep ? (int)((it - a.warm_up_num) / a.thin_rate) : 0;
                    if (tr) a.decision_chain[it - 1] = accepted ? 1 : 0;
                    if (it >= a.iter_end) {                                     // chain finished (state_q holds its position)
                        a.state_eprev[m] = E_prev;
                        want = true;
                    } else {
                        it += 1;
                        cmd |= CMD_REFRESH;
                        sh->cit[chain] = it;
                        fetch = true;
                    }
                    // the next gradient of a continuing chain is the first point of its new trajectory (position and
                    // momentum are in place after P3, before the pass is issued)
                    sh->mode[chain] = want ? MODE_IDLE : MODE_FIRST;
                } else {
                    l += 1;
                    sh->mode[chain] = (l == L) ? MODE_LAST : MODE_MID;
                }
            }
            if (want) {
                // NOTE: a finished chain's STORE/RESTORE command is applied first (P3), the new chain is loaded in the
                // next pass -- `want` stays set until then.
                if (cmd == 0) {
                    want = false;
                    const unsigned int nxt = atomicAdd(queue, 1u);
                    if (nxt < (unsigned int)a.Nchain) {
                        m = (long)nxt;
                        it = a.iter_begin + 1;
                        init = (a.iter_begin == 0);
                        if (!init) E_prev = a.state_eprev[m];
                        if (init && a.decision_chain && a.chain_id0 + m == 0) a.decision_chain[a.N_save_chain0] = 0;
                        cmd = CMD_NEW | CMD_REFRESH | (init ? CMD_INIT0 : 0);
                        sh->cm[chain] = (int)m;
                        sh->cit[chain] = it;
                        fetch = true;
                        sh->mode[chain] = MODE_FIRST;
                    } else {
                        m = -1;
                        cmd = CMD_PARK;
                        sh->mode[chain] = MODE_IDLE;
                    }
                }
            }
            sh->cmd[chain] = cmd;
        }
        __syncthreads();
        TP_T(t4);
        TP_ADD(2, t3, t4);

        // ===== P3. apply the commands: sample store / restore / new chain (all four slice threads of a chain) =============
        {
            const int cmd = sh->cmd[chain];
            if (cmd & (CMD_STORE_Q0 | CMD_STORE_OUT | CMD_RESTORE | CMD_NEW | CMD_PARK)) {
                const long mc = sh->cm[chain];
                const int nch4 = wide ? 7 : 6;
                float x[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = 0.f;
                if (cmd & CMD_PARK) {
                    store_slice(x, nch4);
#pragma unroll
                    for (int j = 0; j < 28; ++j) p[j] = 0.f;
                } else if (cmd & CMD_NEW) {
                    const float* src = ((a.iter_begin == 0) ? (const float*)a.q_start : q0g) + (size_t)mc * D + j0;
#pragma unroll
                    for (int c = 0; c < 7; ++c) {
                        if (c < nch4) {
                            const float4 v = *reinterpret_cast<const float4*>(src + 4 * c);
                            if (a.iter_begin == 0) {
                                *reinterpret_cast<float4*>(q0g + (size_t)mc * D + j0 + 4 * c) = v;
                                *reinterpret_cast<float4*>(q_chain + (size_t)mc * Lc * D + j0 + 4 * c) = v;     // samplers.py:413
                            }
                            x[4 * c] = v.x - mu_s[j0 + 4 * c]; x[4 * c + 1] = v.y - mu_s[j0 + 4 * c + 1];
                            x[4 * c + 2] = v.z - mu_s[j0 + 4 * c + 2]; x[4 * c + 3] = v.w - mu_s[j0 + 4 * c + 3];
                        }
                    }
                    store_slice(x, nch4);
                } else {
                    float* dst = q_chain + ((size_t)mc * Lc + sh->cidx[chain]) * D + j0;
                    float* q0 = q0g + (size_t)mc * D + j0;
                    if (cmd & CMD_RESTORE) {
#pragma unroll
                        for (int c = 0; c < 7; ++c) {
                            if (c < nch4) {
                                const float4 v = *reinterpret_cast<const float4*>(q0 + 4 * c);
                                if (cmd & CMD_STORE_OUT) *reinterpret_cast<float4*>(dst + 4 * c) = v;
                                x[4 * c] = v.x - mu_s[j0 + 4 * c]; x[4 * c + 1] = v.y - mu_s[j0 + 4 * c + 1];
                                x[4 * c + 2] = v.z - mu_s[j0 + 4 * c + 2]; x[4 * c + 3] = v.w - mu_s[j0 + 4 * c + 3];
                            }
                        }
                        store_slice(x, nch4);
                    } else {
#pragma unroll
                        for (int c = 0; c < 7; ++c) {
                            if (c < nch4) {
                                const float4 dd = Df[(j0 / 4 + c) * TC_M + chain];
                                const float4 v = make_float4(dd.x + mu_s[j0 + 4 * c], dd.y + mu_s[j0 + 4 * c + 1],
                                                             dd.z + mu_s[j0 + 4 * c + 2], dd.w + mu_s[j0 + 4 * c + 3]);
                                *reinterpret_cast<float4*>(q0 + 4 * c) = v;
                                if (cmd & CMD_STORE_OUT) *reinterpret_cast<float4*>(dst + 4 * c) = v;
                            }
                        }
                    }
                }
            }
        }
        // ---- momentum refresh (samplers.py:431, 441, 461): the flagged chains of a 32-chain group are drawn by the four
        //      warps of the group in turn, TC_SLOTS rows per round, then taken by the slice threads ------------------------
        {
            unsigned pending = __ballot_sync(HMC_FULL_MASK, (sh->cmd[chain] & CMD_REFRESH) != 0);
            while (__syncthreads_or(pending != 0u)) {
                unsigned todo = pending;
                int k = 0;
                unsigned taken = 0;
                while (todo && k < TC_SLOTS) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    if ((k & 3) == slice) {                       // this warp draws the k-th flagged chain of its group
                        const int cs = grp * 32 + src;
                        const long m_s = sh->cm[cs];
                        const int it_s = sh->cit[cs];
                        const uint64_t gid = (uint64_t)(a.chain_id0 + m_s);
                        float* st = stage_all + (grp * TC_SLOTS + k) * 128;
                        float ks, ln; int Lx;
                        if (sh->cmd[cs] & CMD_INIT0) {            // samplers.py:415: chain-start momentum, K only
                            tc_gen(ga, m_s, gid, 0, lane, st, &ks, &Lx, &ln);
                            if (lane == 0) sh->gK0[cs] = ks;
                            __syncwarp();
                        }
                        tc_gen(ga, m_s, gid, it_s, lane, st, &ks, &Lx, &ln);
                        if (lane == 0) { sh->gK[cs] = ks; sh->gL[cs] = Lx; sh->glnu[cs] = ln; }
                    }
                    taken |= 1u << src;
                    ++k;
                }
                __syncthreads();
                if (taken & (1u << lane)) {                       // my chain's row is staged: take my slice
                    const int kk = __popc(taken & ((1u << lane) - 1u));
                    const float* st = stage_all + (grp * TC_SLOTS + kk) * 128 + j0;
#pragma unroll
                    for (int c = 0; c < 7; ++c) {
                        if (c < 6 || wide) {
                            const float4 v = *reinterpret_cast<const float4*>(st + 4 * c);
                            p[4 * c] = v.x; p[4 * c + 1] = v.y; p[4 * c + 2] = v.z; p[4 * c + 3] = v.w;
                        }
                    }
                }
                pending &= ~taken;
                __syncthreads();
            }
        }
        TP_T(t5);
        TP_ADD(3, t4, t5);

        // ===== P4. done?  otherwise issue the next gradient pass ==========================================================
        const bool alive = (slice == 0) && (m >= 0 || want);
        asm volatile("fence.proxy.async.shared::cta;");         // generic-proxy writes of the operand rows -> tensor core
        asm volatile("tcgen05.fence::before_thread_sync;");
        if (__syncthreads_or(alive) == 0) break;
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint32_t a0 = smem_u32(Ap), b0 = smem_u32(Bp);
            // small terms first: (1,3) (3,1) (2,2) (1,2) (2,1) (1,1)
            const int pa[6] = {0, 2, 1, 0, 1, 0}, pb[6] = {2, 0, 1, 1, 0, 0};
            uint32_t acc = 0;
#pragma unroll
            for (int t = 0; t < 6; ++t) {
#pragma unroll
                for (int ks = 0; ks < KP / 16; ++ks) {
                    const uint64_t da = make_desc(a0 + pa[t] * TC_APART + ks * 2 * TC_M * 16, TC_M * 16, 128);
                    const uint64_t db = make_desc(b0 + pb[t] * TC_BPART + ks * 2 * KP * 16, KP * 16, 128);
                    asm volatile(
                        "{\n\t.reg .pred pacc;\n\tsetp.ne.b32 pacc, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, pacc;\n\t}"
                        ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u));
                    acc = 1;
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
        }
        have_grad = true;
        TP_T(t6);
        TP_ADD(4, t5, t6);
#ifdef HMC_PROFILE_PHASES
        tph[5] += 1;
#endif
    }
#ifdef HMC_PROFILE_PHASES
    if (lane == 0) for (int i = 0; i < 6; ++i) atomicAdd(&g_tc_cycles[i], (unsigned long long)tph[i]);
#endif

    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
    if (a.counters) {
        const unsigned long long c0 = warp_sum<unsigned long long>(n_acc_warm), c1 = warp_sum<unsigned long long>(n_acc_post);
        const unsigned long long c2 = warp_sum<unsigned long long>(n_sumL), c3 = warp_sum<unsigned long long>(n_sumL2);
        if (lane == 0) {
            atomicAdd(a.counters + 0, c0); atomicAdd(a.counters + 1, c1);
            atomicAdd(a.counters + 2, c2); atomicAdd(a.counters + 3, c3);
        }
    }
}

}  // namespace

#ifdef HMC_PROFILE_PHASES
extern "C" int hmc_debug_tc_cycles(unsigned long long* out8, int reset) {
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_tc_cycles, z, sizeof(z)); return 0; }
    cudaMemcpyFromSymbol(out8, g_tc_cycles, sizeof(unsigned long long) * 8);
    return 0;
}
#endif

static size_t tc_smem_bytes() {
    return 3 * (size_t)TC_APART + 3 * (size_t)TC_BPART + (size_t)TC_DCH * TC_M * 16 + sizeof(float) * (2 * TC_KP + 4 * TC_SLOTS * 128) +
           sizeof(TcShared) + 64;
}

bool hmc_random_tc_supported(const hmc_random_args& a, const char** why) {
    if (a.dtype != HMC_F32) { *why = "float32 only"; return false; }
    if (a.target.Mit || a.target.Pt || a.target.Lct) { *why = "identity momentum metric only"; return false; }
    if (a.target.D != TC_ND) { *why = "D == 100 in this build"; return false; }
    if (a.iter_end <= a.iter_begin) { *why = "needs at least one iteration"; return false; }
    if (!a.state_g) { *why = "state_g scratch required"; return false; }
    return true;
}

int hmc_random_run_tc(const hmc_random_args& a, cudaStream_t stream) {
    const size_t smem = tc_smem_bytes();
    HMC_CUDA_CHECK(cudaFuncSetAttribute(hmc_random_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0;
    HMC_CUDA_CHECK(cudaGetDevice(&dev));
    HMC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int grid = (a.Nchain + TC_M - 1) / TC_M;
    if (grid > sms) grid = sms;                 // persistent: one CTA per SM, chain slots pull chains from the queue
    unsigned int* queue = (unsigned int*)a.state_g;
    HMC_CUDA_CHECK(cudaMemsetAsync(queue, 0, sizeof(unsigned int), stream));
    hmc_random_tc_kernel<<<grid, TC_THREADS, smem, stream>>>(a, queue);
    HMC_CUDA_CHECK(cudaGetLastError());
    return HMC_OK;
}
