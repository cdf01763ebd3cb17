// Tensor-core random-trajectory HMC kernel (tcgen05 + TMEM), FP32-grade arithmetic through a bf16x3 split, D = 100.
//
// Follows HMC_sampler.gen_sample_random + leap_frog (/root/reference/samplers.py:387-491, 831-839).
//
// One CTA = 128 chains = 128 threads; THREAD t OWNS CHAIN t: its shifted position d = q - mu and its momentum p
// live in that thread's registers, so every per-chain quantity (energies, Metropolis accept, trajectory length,
// bookkeeping) is thread-local -- no cross-lane reductions.  The gradient of all 128 chains,
//       G[128 x N] = Dm[128 x K] * F[N x K]^T          (K = N = 112: D = 100 zero-padded to a multiple of 16)
// runs on the 5th-generation tensor cores: the positions are written to shared memory as three bf16 parts
// (d = d1 + d2 + d3 exactly), the force matrix is split once the same way, and six tcgen05.mma passes
// (1,3) (3,1) (2,2) (1,2) (2,1) (1,1) accumulate in fp32 in tensor memory -- the dropped terms are O(2^-24).  The
// accumulator row of chain t is TMEM lane t, read back with tcgen05.ld by thread t (the "32x32b" shape).
// Shared-memory operands use the canonical no-swizzle K-major layout: 16-byte chunk (kc, row) at (kc*rows + row)*16
// (LBO = rows*16 between K chunks, SBO = 128 between 8-row groups), which makes the per-thread row writes
// conflict-free 128-bit stores.
//
// Every chain advances one gradient evaluation per pass; iteration boundaries are per-thread events:
// the first point of each trajectory is a gradient-only pass (E_initial, first half kick), so an iteration costs
// L + 1 evaluations.  Momentum refresh is warp-cooperative (one Philox call per lane, same draws as every other
// kernel); finished chains pull the next chain from a global queue.
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
constexpr int TC_KC = TC_KP / 8;    // 16-byte chunks (8 bf16) per row
constexpr int TC_M = 128;           // chains per CTA
constexpr int TC_APART = TC_KC * TC_M * 16;     // bytes of one A part
constexpr int TC_BPART = TC_KC * TC_KP * 16;    // bytes of one B part

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version 1 (Blackwell); no swizzle, base offset 0
    return d;
}

// x = b1 + b2 + b3 exactly (three bf16 parts); returns the parts of two values packed as bf16x2 (lo = x0, hi = x1)
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

__global__ void __launch_bounds__(TC_M, 1) hmc_random_tc_kernel(const hmc_random_args a, unsigned int* __restrict__ queue) {
    constexpr int D = TC_ND, KP = TC_KP, KC = TC_KC;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* Ap = smem;                                   // 3 parts [KC][128] 16-byte chunks
    unsigned char* Bp = Ap + 3 * TC_APART;                      // 3 parts [KC][112]
    float* mu_s = reinterpret_cast<float*>(Bp + 3 * TC_BPART);  // [KP]
    float* dt_s = mu_s + KP;                                    // [KP]
    float* stage_all = dt_s + KP;                               // [4 warps][128] momentum staging
    uint64_t* mbar = reinterpret_cast<uint64_t*>(stage_all + 4 * 128);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* stage = stage_all + warp * 128;

    // ---- one-time set-up: zero A (padding chunks stay zero), split the force matrix into bf16 parts ------------
    for (int t = tid; t < 3 * TC_APART / 16; t += TC_M) reinterpret_cast<uint4*>(Ap)[t] = make_uint4(0u, 0u, 0u, 0u);
    {
        const float* Ft = (const float*)a.target.Ft;            // Ft[k][n] = F[n][k]; B row n holds F[n][.] (K-major)
        const int Dpad = a.target.D_pad;
        for (int t = tid; t < KC * KP; t += TC_M) {             // one 16-byte chunk (8 k values) of row n per item
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
        for (int t = tid; t < KP; t += TC_M) {
            mu_s[t] = (t < D) ? ((const float*)a.target.mu)[t] : 0.f;
            dt_s[t] = (t < D) ? ((const float*)a.target.dt)[t] : 0.f;
        }
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
    const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KP >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

    const long Lc = 1 + (a.Niter - a.warm_up_num) / a.thin_rate;      // samplers.py:31
    float* q_chain = (float*)a.q_chain;
    float* q0g = (float*)a.state_q;
    const double vconst = a.target.v_const;
    const float dt0 = dt_s[0];
    const bool udt = (a.flags & 1) != 0;
    TcGen ga;
    ga.seed = a.seed; ga.p_tape = a.p_tape; ga.L_tape = a.L_tape; ga.u_tape = a.u_tape;
    ga.Niter = a.Niter; ga.L_low = a.L_low; ga.L_high = a.L_high;

    // ---- per-thread chain state ------------------------------------------------------------------------------------
    float d[D], p[D];
#pragma unroll
    for (int j = 0; j < D; ++j) { d[j] = 0.f; p[j] = 0.f; }
    long m = -1;                     // local chain index, -1 = no chain
    int it = 0, l = 0, L = 1;        // iteration, point index of the next gradient, trajectory length
    bool init = false;               // chain start: E_chain[.,0] still to be recorded
    bool want = true;                // needs a (new) chain
    bool refresh = false;            // needs the momentum of iteration `it`
    double E_init = 0.0, E_prev = 0.0;
    float K0 = 0.f, Knew = 0.f, lnu = 0.f;
    unsigned int n_acc_warm = 0, n_acc_post = 0, n_sumL = 0, n_sumL2 = 0;
    uint32_t phase = 0;
#ifdef HMC_PROFILE_PHASES
    long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif

    while (true) {
        TP_T(t0);
        // ===== A. chains that need a new chain id / a new momentum (per-thread decisions, warp-cooperative draws) ===
        if (want) {
            want = false;
            const unsigned int nxt = atomicAdd(queue, 1u);
            if (nxt < (unsigned int)a.Nchain) {
                m = (long)nxt;
                it = a.iter_begin + 1;
                const float* src = (a.iter_begin == 0) ? (const float*)a.q_start + (size_t)m * D : q0g + (size_t)m * D;
#pragma unroll
                for (int j4 = 0; j4 < D / 4; ++j4) {
                    const float4 v = *reinterpret_cast<const float4*>(src + 4 * j4);
                    d[4 * j4 + 0] = v.x - mu_s[4 * j4 + 0]; d[4 * j4 + 1] = v.y - mu_s[4 * j4 + 1];
                    d[4 * j4 + 2] = v.z - mu_s[4 * j4 + 2]; d[4 * j4 + 3] = v.w - mu_s[4 * j4 + 3];
                    if (a.iter_begin == 0) {
                        *reinterpret_cast<float4*>(q0g + (size_t)m * D + 4 * j4) = v;
                        *reinterpret_cast<float4*>(q_chain + (size_t)m * Lc * D + 4 * j4) = v;          // samplers.py:413
                    }
                }
                init = (a.iter_begin == 0);
                if (!init) E_prev = a.state_eprev[m];
                if (init && a.decision_chain && a.chain_id0 + m == 0) a.decision_chain[a.N_save_chain0] = 0;
                refresh = true;
            } else {
                m = -1;
#pragma unroll
                for (int j = 0; j < D; ++j) { d[j] = 0.f; p[j] = 0.f; }
            }
        }
        {
            unsigned need = __ballot_sync(HMC_FULL_MASK, refresh);
            while (need) {
                const int src = __ffs(need) - 1;
                need &= need - 1;
                const long m_s = __shfl_sync(HMC_FULL_MASK, m, src);
                const int it_s = __shfl_sync(HMC_FULL_MASK, it, src);
                const int init_s = __shfl_sync(HMC_FULL_MASK, (int)init, src);
                const uint64_t gid = (uint64_t)(a.chain_id0 + m_s);
                float ks, ln; int Lx;
                if (init_s) {                                               // samplers.py:415: chain-start momentum, K only
                    tc_gen(ga, m_s, gid, 0, lane, stage, &ks, &Lx, &ln);
                    if (lane == src) K0 = 0.5f * ks;
                    __syncwarp();
                }
                tc_gen(ga, m_s, gid, it_s, lane, stage, &ks, &Lx, &ln);     // samplers.py:431, 441, 461
                if (lane == src) {
#pragma unroll
                    for (int j4 = 0; j4 < D / 4; ++j4) {
                        const float4 v = *reinterpret_cast<const float4*>(stage + 4 * j4);
                        p[4 * j4 + 0] = v.x; p[4 * j4 + 1] = v.y; p[4 * j4 + 2] = v.z; p[4 * j4 + 3] = v.w;
                    }
                    Knew = 0.5f * ks; L = Lx; lnu = ln; l = 0;
                    n_sumL += (unsigned int)Lx; n_sumL2 += (unsigned int)(Lx * Lx);
                    refresh = false;
                    if (a.phi_q && a.chain_id0 + m == 0 && it <= a.N_save_chain0) {          // samplers.py:442-445
                        double* phi = a.phi_q + (size_t)(it - 1) * a.L_high * 2;
                        phi[0] = (double)(d[0] + mu_s[0]); phi[1] = (double)(d[1] + mu_s[1]);
                        a.phi_len[it - 1] = Lx + 1;
                    }
                }
                __syncwarp();
            }
        }
        TP_T(t1);
        TP_ADD(0, t0, t1);
        // ===== B. done when no chain is left in the CTA ===============================================================
        if (__syncthreads_or(m >= 0) == 0) break;

        // ===== C. positions -> shared memory (three bf16 parts), then the six tensor-core passes =======================
#pragma unroll
        for (int kc = 0; kc < (D + 7) / 8; ++kc) {
            uint32_t w1[4], w2[4], w3[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j0 = kc * 8 + 2 * e;
                const float x0 = (j0 < D) ? d[j0 < D ? j0 : 0] : 0.f;
                const float x1 = (j0 + 1 < D) ? d[j0 + 1 < D ? j0 + 1 : 0] : 0.f;
                split3(x0, x1, w1[e], w2[e], w3[e]);
            }
            const int off = (kc * TC_M + tid) * 16;
            *reinterpret_cast<uint4*>(Ap + off) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
            *reinterpret_cast<uint4*>(Ap + TC_APART + off) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
            *reinterpret_cast<uint4*>(Ap + 2 * TC_APART + off) = make_uint4(w3[0], w3[1], w3[2], w3[3]);
        }
        TP_T(t2);
        TP_ADD(1, t1, t2);
        asm volatile("fence.proxy.async.shared::cta;");         // generic-proxy writes -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
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
        {
            uint32_t done = 0;
            while (!done) {
                asm volatile("{\n\t.reg .pred pw;\n\tmbarrier.try_wait.parity.shared::cta.b64 pw, [%1], %2;\n\tselp.u32 %0, 1, 0, pw;\n\t}"
                             : "=r"(done) : "r"(smem_u32(mbar)), "r"(phase) : "memory");
            }
            phase ^= 1u;
        }
        asm volatile("tcgen05.fence::after_thread_sync;");
        TP_T(t3);
        TP_ADD(2, t2, t3);

        // ===== D. thread-local leapfrog update (samplers.py:835-837) and energies ======================================
        // point index l of the gradient just evaluated: 0 = first point (half kick, drift), L = last (half kick, no
        // drift), otherwise interior (second half kick of step l + first half kick of step l+1, drift).
        const bool running = m >= 0;
        const bool first = running && l == 0, last = running && l == L;
        const float kwt = running ? ((first || last) ? -0.5f : -1.0f) : 0.f;
        const float dwt = (running && !last) ? 1.f : 0.f;
        float hv = 0.f, hk = 0.f;
#pragma unroll
        for (int c0 = 0; c0 < KP; c0 += 16) {
            if (c0 < D) {
                uint32_t v[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(tmem_row + (uint32_t)c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const int j = c0 + c;
                    if (j < D) {
                        const float gj = __uint_as_float(v[c]);
                        const float dtj = udt ? dt0 : dt_s[j];
                        hv = fmaf(d[j], gj, hv);
                        const float pn = fmaf(gj, kwt * dtj, p[j]);
                        hk = fmaf(pn, pn, hk);
                        p[j] = pn;
                        d[j] = fmaf(pn, dwt * dtj, d[j]);
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;");      // TMEM reads ordered before the next pass's MMA

        TP_T(t4);
        TP_ADD(3, t3, t4);
        // ===== E. per-thread bookkeeping ================================================================================
        if (running) {
            const bool tr = a.phi_q && (a.chain_id0 + m) == 0 && it <= a.N_save_chain0;
            const double V = 0.5 * (double)hv + vconst;                     // V(q) = 0.5 d.P d + const (utils.py:213-218)
            if (first) {
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
                if (tr) { double* phi = a.phi_q + (size_t)(it - 1) * a.L_high * 2; phi[2] = (double)(d[0] + mu_s[0]); phi[3] = (double)(d[1] + mu_s[1]); }
            } else if (last) {
                // Metropolis accept (samplers.py:455-472)
                const double E_final = V + 0.5 * (double)hk;
                const double dE = E_final - E_init;
                E_prev = E_init;                                            // samplers.py:460
                const bool accepted = (dE < 0) || ((double)lnu < -dE);      // samplers.py:462
                const bool keep = it >= a.warm_up_num;
                const long idx = keep ? (it - a.warm_up_num) / a.thin_rate : 0;
                float* dst = q_chain + ((size_t)m * Lc + idx) * D;
                float* q0 = q0g + (size_t)m * D;
                if (accepted) {
                    if (keep) n_acc_post++; else n_acc_warm++;
#pragma unroll
                    for (int j4 = 0; j4 < D / 4; ++j4) {
                        const float4 v = make_float4(d[4 * j4] + mu_s[4 * j4], d[4 * j4 + 1] + mu_s[4 * j4 + 1],
                                                     d[4 * j4 + 2] + mu_s[4 * j4 + 2], d[4 * j4 + 3] + mu_s[4 * j4 + 3]);
                        *reinterpret_cast<float4*>(q0 + 4 * j4) = v;
                        if (keep) *reinterpret_cast<float4*>(dst + 4 * j4) = v;
                    }
                } else {
#pragma unroll
                    for (int j4 = 0; j4 < D / 4; ++j4) {
                        const float4 v = *reinterpret_cast<const float4*>(q0 + 4 * j4);
                        if (keep) *reinterpret_cast<float4*>(dst + 4 * j4) = v;
                        d[4 * j4] = v.x - mu_s[4 * j4]; d[4 * j4 + 1] = v.y - mu_s[4 * j4 + 1];
                        d[4 * j4 + 2] = v.z - mu_s[4 * j4 + 2]; d[4 * j4 + 3] = v.w - mu_s[4 * j4 + 3];
                    }
                }
                if (tr) a.decision_chain[it - 1] = accepted ? 1 : 0;
                if (it >= a.iter_end) {                                     // chain finished (state_q holds its position)
                    a.state_eprev[m] = E_prev;
                    want = true;
                } else {
                    it += 1;
                    refresh = true;
                }
            } else {
                l += 1;
                if (tr) { double* phi = a.phi_q + (size_t)(it - 1) * a.L_high * 2; phi[2 * l] = (double)(d[0] + mu_s[0]); phi[2 * l + 1] = (double)(d[1] + mu_s[1]); }
            }
        }
        TP_T(t5);
        TP_ADD(4, t4, t5);
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

bool hmc_random_tc_supported(const hmc_random_args& a, const char** why) {
    if (a.dtype != HMC_F32) { *why = "float32 only"; return false; }
    if (a.target.Mit || a.target.Pt || a.target.Lct) { *why = "identity momentum metric only"; return false; }
    if (a.target.D != TC_ND) { *why = "D == 100 in this build"; return false; }
    if (a.iter_end <= a.iter_begin) { *why = "needs at least one iteration"; return false; }
    if (!a.state_g) { *why = "state_g scratch required"; return false; }
    return true;
}

int hmc_random_run_tc(const hmc_random_args& a, cudaStream_t stream) {
    const size_t smem = 3 * (size_t)TC_APART + 3 * (size_t)TC_BPART + sizeof(float) * (2 * TC_KP + 4 * 128) + 64;
    HMC_CUDA_CHECK(cudaFuncSetAttribute(hmc_random_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0;
    HMC_CUDA_CHECK(cudaGetDevice(&dev));
    HMC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int grid = (a.Nchain + TC_M - 1) / TC_M;
    if (grid > sms) grid = sms;                 // persistent: one CTA per SM, threads pull chains from the queue
    unsigned int* queue = (unsigned int*)a.state_g;
    HMC_CUDA_CHECK(cudaMemsetAsync(queue, 0, sizeof(unsigned int), stream));
    hmc_random_tc_kernel<<<grid, TC_M, smem, stream>>>(a, queue);
    HMC_CUDA_CHECK(cudaGetLastError());
    return HMC_OK;
}
