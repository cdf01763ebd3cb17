// Fused random-trajectory HMC kernel, FP32, 20 < D <= 128, identity momentum metric (the production path).
//
// Follows HMC_sampler.gen_sample_random + leap_frog (/root/reference/samplers.py:387-491, 831-839) with one
// gradient evaluation per leapfrog step (the second gradient of step l is the first of step l+1).
//
// Work decomposition (measured design probes: profiles/microbench_r1_design_probes.txt):
//   * A WARP is autonomous: it owns NSLOT = NCG*8 chain slots and never synchronises with other warps.
//     lane = (cg, dg): chain group cg (8 chains) x dimension group dg (TN dimensions jj*NDG + dg); the lane
//     keeps the gradient accumulators g[8][TN] and the momenta p[8][TN] of its tile in registers.
//   * The positions live in a per-warp shared-memory tile  Ds[k][slot]  (shifted coordinates d = q - mu), the
//     precision matrix in a CTA-wide tile  Ps[k][.]  (columns permuted to the lane order); the gradient
//     g[c][j] = sum_k d[k][c] P[k][j]  is an FFMA2 loop: two chains per packed FMA, P[k][j] enters as the
//     scalar-broadcast operand (SASS: FFMA2 R, R.F32x2.HI_LO, R.F32, R).
//   * Chains advance asynchronously: every pass of the main loop is one gradient + leapfrog update
//     for every slot; a slot whose trajectory ends is serviced (energy, Metropolis accept on a Philox uniform,
//     sample store, momentum refresh, new L) without stalling the others, and a slot whose chain is finished
//     pulls the next chain from a global queue.
// HBM sees only the stored sample / energy stream and the final chain state.
#include "hmc_common.cuh"

// per-phase cycle accounting for profiles/phase_cycles.py (debug build only)
#ifdef HMC_PROFILE_PHASES
__device__ unsigned long long g_phase_cycles[8];
#define PH_T(x) const long long x = clock64()
#define PH_ADD(i, a, b) ph[i] += (b) - (a)
#else
#define PH_T(x)
#define PH_ADD(i, a, b)
#endif

namespace {


__device__ __forceinline__ float2 fma2(const float2 a, const float2 b, const float2 c) {
    unsigned long long D;
    const unsigned long long A = *reinterpret_cast<const unsigned long long*>(&a);
    const unsigned long long B = *reinterpret_cast<const unsigned long long*>(&b);
    const unsigned long long Cc = *reinterpret_cast<const unsigned long long*>(&c);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(D) : "l"(A), "l"(B), "l"(Cc));
    return *reinterpret_cast<float2*>(&D);
}
__device__ __forceinline__ float2 bc2(const float x) { return make_float2(x, x); }

enum SlotState : int { ST_IDLE = 0, ST_RUN = 1, ST_NEED_CHAIN = 2, ST_ACC = 3, ST_REJ = 4 };

template <int TM, int TN, int NDG, int NCG>
struct Geo {
    static constexpr int NSLOT = NCG * TM;       // chain slots per warp
    static constexpr int DP = NDG * TN;          // padded dimension
    // row stride of the position tile: a multiple of 4 floats with QS/4 odd, so that the lanes of a quarter warp
    // hit distinct banks in the update pass
    static constexpr int QS = ((NSLOT + 3) / 4) % 2 ? (NSLOT + 3) / 4 * 4 : (NSLOT + 3) / 4 * 4 + 4;
    static constexpr int RW = 2 * TM;            // floats per lane in the reduction scratch
    static constexpr int BK = 32 * 16;           // per-slot bookkeeping block (16 words per slot)
    static constexpr int STAGE = (DP + 31) / 32 * 32 + 32;
    static constexpr int RED = 32 * 2 * TM;
};

// Per-slot bookkeeping, kept in shared memory (one 64-byte record per slot) so that it costs no registers in
// the gradient loop.  Only lane s touches record s, except for the warp-uniform reads in the service loop.
struct __align__(16) SlotBk {
    int state, m, it, l;
    int L, init;
    float K0, Knew;
    double Einit, Eprev, V;
    float lnu;
    int pad;
};
static_assert(sizeof(SlotBk) == 64, "SlotBk layout");

struct GenArgs {
    uint64_t seed;
    const double* p_tape;
    const int32_t* L_tape;
    const double* u_tape;
    int D, Niter, L_low, L_high;
};

// Momentum refresh for one chain, warp-cooperative: lane sl draws the 4 normals of dims 4*sl .. 4*sl+3 into
// `stage`; the lane after the last momentum slot draws the scalars of the iteration (trajectory length and
// acceptance uniform, samplers.py:441, 461) in the same Philox pass.  Returns sum p^2 to every lane.
__device__ __noinline__ void gen_momentum(const GenArgs a, long m, uint64_t gid, int iter, int lane, float* stage,
                                          float* sumsq, int* L, float* lnu) {
    const int D = a.D;
    float s = 0.f;
    int Lv = 1;
    float lv = 0.f;
    if (a.p_tape) {
        const double* src = a.p_tape + ((size_t)m * (a.Niter + 1) + iter) * D;
        for (int j = lane; j < D; j += 32) { const float v = (float)src[j]; stage[j] = v; s = fmaf(v, v, s); }
        if (iter >= 1) {
            Lv = a.L_tape[(size_t)m * a.Niter + iter - 1];
            lv = (float)log(a.u_tape[(size_t)m * a.Niter + iter - 1]);
        }
    } else {
        const int nslot = (D + 3) >> 2;
        const bool scalar_lane = (lane == nslot);        // a free lane exists while nslot < 32 (D <= 124)
        const uint32_t hi = (uint32_t)(gid >> 32) << 8;
        const Philox4 r = philox4x32_10((uint32_t)gid, (uint32_t)iter, scalar_lane ? 0u : (uint32_t)lane,
                                        (scalar_lane ? (uint32_t)HMC_STREAM_SCALAR : (uint32_t)HMC_STREAM_MOMENTUM) | hi,
                                        (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
        // same arithmetic as hmc_normal4 / hmc_scalar_draws (hmc_common.cuh)
        const float u1 = ((float)(r.x >> 8) + 0.5f) * 5.9604644775390625e-08f;
        const float u2 = ((float)(r.z >> 8) + 0.5f) * 5.9604644775390625e-08f;
        const float r1 = sqrtf(-2.0f * __logf(u1));
        const float r2 = sqrtf(-2.0f * __logf(u2));
        const float a1 = ((float)(r.y >> 8) * 5.9604644775390625e-08f - 0.5f) * 6.283185307179586f;
        const float a2 = ((float)(r.w >> 8) * 5.9604644775390625e-08f - 0.5f) * 6.283185307179586f;
        const float zz[4] = {r1 * __cosf(a1), r1 * __sinf(a1), r2 * __cosf(a2), r2 * __sinf(a2)};
        if (lane < nslot) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = 4 * lane + q;
                if (j < D) { stage[j] = zz[q]; s = fmaf(zz[q], zz[q], s); }
            }
        }
        const int Ls = a.L_low + (int)__umulhi(r.x, (uint32_t)(a.L_high - a.L_low));
        const float ls = logf(((float)(r.y >> 8) + 0.5f) * 5.9604644775390625e-08f);
        if (nslot < 32) {
            Lv = __shfl_sync(HMC_FULL_MASK, Ls, nslot);
            lv = __shfl_sync(HMC_FULL_MASK, ls, nslot);
        } else if (iter >= 1) {                              // all 32 lanes carry momentum slots: separate scalar draw
            double uu;
            hmc_scalar_draws(a.seed, gid, (uint32_t)iter, a.L_low, a.L_high, &Lv, &uu);
            lv = logf((float)uu);
        }
    }
    *sumsq = warp_sum<float>(s);
    *L = Lv;
    *lnu = lv;
    __syncwarp();
}

template <int TM, int TN, int NDG, int NCG, int WARPS, bool UDT, bool FULL, bool PINGPONG>
__global__ void __launch_bounds__(WARPS * 32, 1) hmc_random_fast_kernel(const hmc_random_args a, unsigned int* __restrict__ queue) {
    using G = Geo<TM, TN, NDG, NCG>;
    static_assert(TM % 4 == 0 && G::NSLOT <= 32, "tile");
    constexpr int NSLOT = G::NSLOT, DP = G::DP, QS = G::QS;
    extern __shared__ __align__(16) float sm[];
    const int D = FULL ? Geo<TM, TN, NDG, NCG>::DP : a.target.D;          // FULL: the tile has no padded dimensions
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* Ps = sm;                       // [D][DP]   Ps[k][dg*TN + jj] = P[k][jj*NDG + dg]
    float* mu_s = Ps + D * DP;            // [DP]
    float* dt_s = mu_s + DP;              // [DP]
    int* done_flags = reinterpret_cast<int*>(dt_s + DP);   // [32] per-warp 'no chain left' flags (ping-pong pairs)
    float* wbase = dt_s + DP + 32 + (size_t)warp * (D * QS + G::RED + G::STAGE + G::BK);
    float* Ds = wbase;                    // [D][QS]   live positions (shifted by mu)
    float* red = Ds + D * QS;             // [32][16]  per-lane partial sums
    float* stage = red + G::RED;          // momentum staging
    SlotBk* bk = reinterpret_cast<SlotBk*>(stage + G::STAGE);   // [32] per-slot bookkeeping
    // The position at the start of the running iteration (what a rejection restores) lives in HBM/L2: state_q.
    float* q0g = (float*)a.state_q;
    {
        const float* Ft = (const float*)a.target.Ft;
        const int Dpad = a.target.D_pad;
        for (int t = threadIdx.x; t < D * DP; t += blockDim.x) {
            const int k = t / DP, c = t - k * DP;
            const int j = (c % TN) * NDG + c / TN;          // lane-order column -> dimension
            Ps[t] = (j < D) ? Ft[(size_t)k * Dpad + j] : 0.f;
        }
        if (threadIdx.x < 32) done_flags[threadIdx.x] = 0;
        for (int t = threadIdx.x; t < DP; t += blockDim.x) {
            mu_s[t] = (t < D) ? ((const float*)a.target.mu)[t] : 0.f;
            dt_s[t] = (t < D) ? ((const float*)a.target.dt)[t] : 0.f;
        }
        for (int t = lane; t < D * QS; t += 32) Ds[t] = 0.f;
        __syncthreads();
    }
    const bool active = lane < NCG * NDG;
    const int cg = active ? lane / NDG : 0;
    const int dg = active ? lane % NDG : 0;
    const long Lc = 1 + (a.Niter - a.warm_up_num) / a.thin_rate;      // samplers.py:31
    const long Lrow = a.store_ring > 0 ? a.store_ring : Lc;         // rows allocated per chain (ring of the last store_ring stored samples)
    float* q_chain = (float*)a.q_chain;
    const float dt0 = dt_s[0];

    float2 g[TM / 2][TN], p[TM / 2][TN];
#pragma unroll
    for (int c = 0; c < TM / 2; ++c)
#pragma unroll
        for (int j = 0; j < TN; ++j) { g[c][j] = make_float2(0.f, 0.f); p[c][j] = make_float2(0.f, 0.f); }

    // ---- per-slot bookkeeping lives in shared memory (SlotBk), record s is owned by lane s < NSLOT -------------
    if (lane < 32) {
        SlotBk z;
        z.state = (lane < NSLOT) ? ST_NEED_CHAIN : ST_IDLE; z.m = -1; z.it = 0; z.l = 0; z.L = 1; z.init = 0;
        z.K0 = 0.f; z.Knew = 0.f; z.Einit = 0.0; z.Eprev = 0.0; z.V = 0.0; z.lnu = 0.f; z.pad = 0;
        bk[lane] = z;
    }
    unsigned long long n_acc_warm = 0, n_acc_post = 0, n_sumL = 0, n_sumL2 = 0;
    GenArgs ga;
    ga.seed = a.seed; ga.p_tape = a.p_tape; ga.L_tape = a.L_tape; ga.u_tape = a.u_tape;
    ga.D = D; ga.Niter = a.Niter; ga.L_low = a.L_low; ga.L_high = a.L_high;
    __syncwarp();

#ifdef HMC_PROFILE_PHASES
    long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    // ---------------------------------------------------------------------------------------------------------
    // The three phases of one pass over the warp's slots.
    // ---------------------------------------------------------------------------------------------------------
    auto do_service = [&]() {
        PH_T(tA);
        // ===== A. service every slot that finished a trajectory or needs a chain (warp-uniform loop) ==========
        unsigned need = __ballot_sync(HMC_FULL_MASK, bk[lane].state >= ST_NEED_CHAIN);
        while (need) {
            PH_T(tS0);
            const int s = __ffs(need) - 1;
            need &= need - 1;
            const int scg = s / TM, sc = s % TM;
            int kind = bk[s].state;                 // uniform (broadcast) reads
            long m = bk[s].m;
            int it = bk[s].it;
            const bool owner = active && (cg == scg);
            float dv[TN];
#pragma unroll
            for (int jj = 0; jj < TN; ++jj) dv[jj] = 0.f;
            if (kind == ST_ACC || kind == ST_REJ) {
                // ---- the trajectory of iteration `it` ended: keep or restore the position, store the sample
                //      (samplers.py:462-472)
                const bool keep = it >= a.warm_up_num;
                const long idx = keep ? hmc_store_index(it, a.warm_up_num, a.thin_rate, Lrow, a.store_ring > 0) : 0;
                if (owner) {
                    float* dst = q_chain + ((size_t)m * Lrow + idx) * D;
                    float* q0 = q0g + (size_t)m * D;
#pragma unroll
                    for (int jj = 0; jj < TN; ++jj) {
                        const int j = jj * NDG + dg;
                        if (FULL || j < D) {
                            float qv;
                            if (kind == ST_ACC) { dv[jj] = Ds[j * QS + s]; qv = dv[jj] + mu_s[j]; q0[j] = qv; }
                            else { qv = q0[j]; dv[jj] = qv - mu_s[j]; Ds[j * QS + s] = dv[jj]; }
                            if (keep) dst[j] = qv;
                        }
                    }
                }
                if (it >= a.iter_end) {
                    // chain finished: state_q already holds its position; the slot asks for the next chain
                    if (lane == s) a.state_eprev[m] = bk[s].Eprev;
                    kind = ST_NEED_CHAIN;
                }
            }
            if (kind == ST_NEED_CHAIN) {
                unsigned int nxt = 0;
                if (lane == s) nxt = atomicAdd(queue, 1u);
                nxt = __shfl_sync(HMC_FULL_MASK, nxt, s);
                if (nxt >= (unsigned int)a.Nchain) {
                    // queue empty: park the slot with finite numbers
                    if (owner) {
#pragma unroll
                        for (int jj = 0; jj < TN; ++jj) { const int j = jj * NDG + dg; if (FULL || j < D) Ds[j * QS + s] = 0.f; }
                    }
                    if (lane == s) { bk[s].state = ST_IDLE; bk[s].m = -1; }
#pragma unroll
                    for (int c2 = 0; c2 < TM / 2; ++c2) {
                        const bool hx = owner && (sc == 2 * c2), hy = owner && (sc == 2 * c2 + 1);
#pragma unroll
                        for (int jj = 0; jj < TN; ++jj) {
                            if (hx) { p[c2][jj].x = 0.f; g[c2][jj].x = 0.f; }
                            if (hy) { p[c2][jj].y = 0.f; g[c2][jj].y = 0.f; }
                        }
                    }
                    __syncwarp();
                    continue;
                }
                m = (long)nxt;
                it = a.iter_begin;
                if (owner) {
                    float* q0 = q0g + (size_t)m * D;
#pragma unroll
                    for (int jj = 0; jj < TN; ++jj) {
                        const int j = jj * NDG + dg;
                        if (FULL || j < D) {
                            float qv;
                            if (a.iter_begin == 0) {
                                qv = ((const float*)a.q_start)[(size_t)m * D + j];
                                q0[j] = qv;
                                q_chain[(size_t)m * Lrow * D + j] = qv;                       // samplers.py:413
                            } else {
                                qv = q0[j];
                            }
                            dv[jj] = qv - mu_s[j];
                            Ds[j * QS + s] = dv[jj];
                        }
                    }
                }
                if (a.iter_begin == 0) {                                                   // samplers.py:415 (K only)
                    float k0, ld; int Ld;
                    gen_momentum(ga, m, (uint64_t)(a.chain_id0 + m), 0, lane, stage, &k0, &Ld, &ld);
                    if (lane == s) { bk[s].K0 = 0.5f * k0; bk[s].init = 1; }
                    if (a.decision_chain && a.chain_id0 + m == 0 && lane == s) a.decision_chain[a.N_save_chain0] = 0;
                } else if (lane == s) {
                    bk[s].Eprev = a.state_eprev[m];
                    bk[s].init = 0;
                }
            }
            PH_T(tS1);
            PH_ADD(6, tS0, tS1);
            // ---- start iteration it+1: momentum refresh (samplers.py:431), trajectory length (:441), uniform (:461)
            const int itn = it + 1;
            const uint64_t gid = (uint64_t)(a.chain_id0 + m);
            float ksum, lnun; int Ln;
            gen_momentum(ga, m, gid, itn, lane, stage, &ksum, &Ln, &lnun);
            PH_T(tS2);
            PH_ADD(7, tS1, tS2);
            const bool go = (kind == ST_ACC);      // gradient at the accepted point is still in g: kick and drift now
            const bool tr = a.phi_q && gid == 0 && itn <= a.N_save_chain0;
            if (lane == s) {
                SlotBk b = bk[s];
                b.m = (int)m; b.it = itn; b.L = Ln; b.lnu = lnun; b.Knew = 0.5f * ksum; b.state = ST_RUN;
                n_sumL += (unsigned long long)Ln; n_sumL2 += (unsigned long long)(Ln * Ln);
                if (go) {
                    // E_initial of the new iteration (samplers.py:434-438): V at the accepted point + new kinetic energy
                    b.Einit = b.V + (double)b.Knew;
                    if (itn >= a.warm_up_num) {
                        const long idx = hmc_store_index(itn, a.warm_up_num, a.thin_rate, Lrow, a.store_ring > 0);
                        a.E_chain[(size_t)m * Lrow + idx] = b.Einit;
                        a.dE_chain[(size_t)m * Lrow + idx] = b.Einit - b.Eprev;
                    }
                    b.l = 1;
                } else {
                    b.l = 0;
                }
                bk[s] = b;
            }
            // ---- owners: new momentum (+ first half kick and drift when the gradient is at hand) -----------------
            {
                const bool odd = sc & 1;
                const int sp = sc >> 1;
#pragma unroll
                for (int jj = 0; jj < TN; ++jj) {
                    const int j = jj * NDG + dg;
                    float pj = (FULL || j < D) ? stage[j] : 0.f;
                    if (go) {
                        float2 g01 = g[0][jj];
#pragma unroll
                        for (int c2 = 1; c2 < TM / 2; ++c2) if (sp == c2) g01 = g[c2][jj];
                        const float gj = odd ? g01.y : g01.x;
                        const float dtj = UDT ? dt0 : dt_s[j];
                        pj = fmaf(gj, -0.5f * dtj, pj);                                        // first half kick (samplers.py:835)
                        if (owner && (FULL || j < D)) Ds[j * QS + s] = fmaf(pj, dtj, dv[jj]);   // drift (samplers.py:836)
                    }
#pragma unroll
                    for (int c2 = 0; c2 < TM / 2; ++c2) {
                        if (owner && sc == 2 * c2) p[c2][jj].x = pj;
                        if (owner && sc == 2 * c2 + 1) p[c2][jj].y = pj;
                    }
                }
            }
            __syncwarp();
            if (tr && lane == s) {
                // chain-0 trajectory capture (samplers.py:442-452): row 0 is the start point, row 1 the first step
                double* phi = a.phi_q + (size_t)(itn - 1) * a.L_high * 2;
                const float* q0 = q0g + (size_t)m * D;
                phi[0] = (double)q0[0];
                if (D > 1) phi[1] = (double)q0[1];
                a.phi_len[itn - 1] = Ln + 1;
                if (go) {
                    phi[2] = (double)(Ds[s] + mu_s[0]);
                    if (D > 1) phi[3] = (double)(Ds[QS + s] + mu_s[1]);
                }
            }
        }

        PH_T(tB);
        PH_ADD(0, tA, tB);
    };
    auto do_gradient = [&]() {
        PH_T(tB);
        // ===== C. gradient  g[c][j] = sum_k d[k][c] P[k][j]  ==================================================
#pragma unroll
        for (int c = 0; c < TM / 2; ++c)
#pragma unroll
            for (int j = 0; j < TN; ++j) g[c][j] = make_float2(0.f, 0.f);
        {
            const float* qp = Ds + cg * TM;
            const float* pp = Ps + dg * TN;
            auto load = [&](int k, float (&qv)[TM], float (&pv)[TN]) {
#pragma unroll
                for (int i = 0; i < TM / 4; ++i) *reinterpret_cast<float4*>(&qv[4 * i]) = *reinterpret_cast<const float4*>(qp + k * QS + 4 * i);
                if constexpr (TN % 4 == 0) {
#pragma unroll
                    for (int i = 0; i < TN / 4; ++i) *reinterpret_cast<float4*>(&pv[4 * i]) = *reinterpret_cast<const float4*>(pp + k * DP + 4 * i);
                } else {
#pragma unroll
                    for (int i = 0; i < TN / 2; ++i) *reinterpret_cast<float2*>(&pv[2 * i]) = *reinterpret_cast<const float2*>(pp + k * DP + 2 * i);
                }
            };
            auto fmas = [&](const float (&qv)[TM], const float (&pv)[TN]) {
#pragma unroll
                for (int c = 0; c < TM / 2; ++c)
#pragma unroll
                    for (int j = 0; j < TN; ++j) g[c][j] = fma2(make_float2(qv[2 * c], qv[2 * c + 1]), bc2(pv[j]), g[c][j]);
            };
            if constexpr (PINGPONG && FULL) {
                // explicit register double buffering: with one warp per scheduler in the loop (ping-pong) the
                // shared-memory latency is otherwise exposed at every iteration (probe F: 78 % vs 68 % of peak)
                float qa[TM], pa[TN], qb[TM], pb[TN];
                load(0, qa, pa);
#pragma unroll 1
                for (int k = 0; k < D; k += 2) {
                    load(k + 1, qb, pb);
                    fmas(qa, pa);
                    load((k + 2 < D) ? k + 2 : 0, qa, pa);
                    fmas(qb, pb);
                }
            } else {
#pragma unroll 2
                for (int k = 0; k < D; ++k) {
                    float qv[TM], pv[TN];
                    load(k, qv, pv);
                    fmas(qv, pv);
                }
            }
        }
        __syncwarp();
        PH_T(tD);
        PH_ADD(1, tB, tD);
    };
    auto do_update = [&]() {
        PH_T(tD);
        // ===== D. leapfrog update of the lane's tile (samplers.py:835-837) + energy partial sums ================
        // point index of the gradient just evaluated: 0 = first point, L = last point of the trajectory.
        // Interior points take the second half kick of step l and the first half kick of step l+1 (kick weight
        // -1), the first and last point one half kick (-1/2); the position moves at every point but the last.
        float2 kw[TM / 2], dw[TM / 2], hv[TM / 2], hk[TM / 2];
        {
            const SlotBk& me = bk[lane];
            const bool running = me.state == ST_RUN;
            const unsigned m_l0 = __ballot_sync(HMC_FULL_MASK, running && me.l == 0);
            const unsigned m_last = __ballot_sync(HMC_FULL_MASK, running && me.l == me.L);
            const unsigned m_mid = __ballot_sync(HMC_FULL_MASK, running && me.l != 0 && me.l != me.L);
#pragma unroll
            for (int c = 0; c < TM / 2; ++c) {
                const int b0 = cg * TM + 2 * c;
                const float mid0 = (float)((m_mid >> b0) & 1u), mid1 = (float)((m_mid >> (b0 + 1)) & 1u);
                const float mv0 = (float)(((m_mid | m_l0) >> b0) & 1u), mv1 = (float)(((m_mid | m_l0) >> (b0 + 1)) & 1u);
                kw[c] = make_float2(-0.5f - 0.5f * mid0, -0.5f - 0.5f * mid1);
                dw[c] = make_float2(mv0, mv1);
                if (UDT) { kw[c].x *= dt0; kw[c].y *= dt0; dw[c].x *= dt0; dw[c].y *= dt0; }
                hv[c] = make_float2(0.f, 0.f);
                hk[c] = make_float2(0.f, 0.f);
            }
            (void)m_last;
        }
#pragma unroll
        for (int jj = 0; jj < TN; ++jj) {
            const int j = jj * NDG + dg;
            const bool jv = FULL || (j < D);
            float dq[TM];
#pragma unroll
            for (int i = 0; i < TM / 4; ++i)
                *reinterpret_cast<float4*>(&dq[4 * i]) = jv ? *reinterpret_cast<const float4*>(Ds + j * QS + cg * TM + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float dtj = UDT ? 1.f : dt_s[j];
#pragma unroll
            for (int c = 0; c < TM / 2; ++c) {
                const float2 dd = make_float2(dq[2 * c], dq[2 * c + 1]);
                const float2 gg = g[c][jj];
                hv[c] = fma2(dd, gg, hv[c]);
                const float2 kc = UDT ? kw[c] : make_float2(kw[c].x * dtj, kw[c].y * dtj);
                const float2 pnw = fma2(gg, kc, p[c][jj]);
                hk[c] = fma2(pnw, pnw, hk[c]);
                p[c][jj] = pnw;
                const float2 dc = UDT ? dw[c] : make_float2(dw[c].x * dtj, dw[c].y * dtj);
                const float2 dn = fma2(pnw, dc, dd);
                dq[2 * c] = dn.x; dq[2 * c + 1] = dn.y;
            }
            if (active && jv) {
#pragma unroll
                for (int i = 0; i < TM / 4; ++i) *reinterpret_cast<float4*>(Ds + j * QS + cg * TM + 4 * i) = *reinterpret_cast<const float4*>(&dq[4 * i]);
            }
        }
        // ---- reduce the partial sums over the dimension groups (shared memory, fixed order => deterministic)
        if (active) {
            float* r = red + (cg * NDG + dg) * G::RW;
#pragma unroll
            for (int c = 0; c < TM / 4; ++c) {
                *reinterpret_cast<float4*>(r + 4 * c) = make_float4(hv[2 * c].x, hv[2 * c].y, hv[2 * c + 1].x, hv[2 * c + 1].y);
                *reinterpret_cast<float4*>(r + TM + 4 * c) = make_float4(hk[2 * c].x, hk[2 * c].y, hk[2 * c + 1].x, hk[2 * c + 1].y);
            }
        }
        __syncwarp();
        PH_T(tE);
        PH_ADD(2, tD, tE);
        // ===== E. per-slot bookkeeping (lane s < NSLOT) ========================================================
        if (lane < NSLOT && bk[lane].state == ST_RUN) {
            const int s = lane, scg = s / TM, sc = s % TM;
            SlotBk b = bk[s];
            const bool tr = a.phi_q && (a.chain_id0 + b.m) == 0 && b.it <= a.N_save_chain0;
            if (b.l != 0 && b.l != b.L) {
                b.l += 1;                                              // interior point: nothing to record
                bk[s].l = b.l;
                if (tr) {
                    double* phi = a.phi_q + (size_t)(b.it - 1) * a.L_high * 2;
                    phi[2 * b.l] = (double)(Ds[s] + mu_s[0]);
                    if (D > 1) phi[2 * b.l + 1] = (double)(Ds[QS + s] + mu_s[1]);
                }
            } else {
                float sv = 0.f, sk = 0.f;
#pragma unroll
                for (int d2 = 0; d2 < NDG; ++d2) {
                    sv += red[(scg * NDG + d2) * G::RW + sc];
                    sk += red[(scg * NDG + d2) * G::RW + TM + sc];
                }
                const double V = 0.5 * (double)sv + a.target.v_const;    // V(q) = 0.5 d.P d + const  (utils.py:213-218)
                if (b.l == 0) {
                    // first point of a trajectory reached through a fresh gradient (chain start or after a rejection)
                    if (b.init) {                                      // samplers.py:416-420
                        const double E0 = V + (double)b.K0;
                        a.E_chain[(size_t)b.m * Lrow] = E0;
                        a.dE_chain[(size_t)b.m * Lrow] = 0.0;
                        b.Eprev = E0;
                        b.init = 0;
                    }
                    b.Einit = V + (double)b.Knew;                      // samplers.py:434-438
                    if (b.it >= a.warm_up_num) {
                        const long idx = hmc_store_index(b.it, a.warm_up_num, a.thin_rate, Lrow, a.store_ring > 0);
                        a.E_chain[(size_t)b.m * Lrow + idx] = b.Einit;
                        a.dE_chain[(size_t)b.m * Lrow + idx] = b.Einit - b.Eprev;
                    }
                    b.l = 1;
                    if (tr) {
                        double* phi = a.phi_q + (size_t)(b.it - 1) * a.L_high * 2;
                        phi[2] = (double)(Ds[s] + mu_s[0]);
                        if (D > 1) phi[3] = (double)(Ds[QS + s] + mu_s[1]);
                    }
                } else {
                    // last point: Metropolis accept (samplers.py:455-472)
                    const double E_final = V + 0.5 * (double)sk;
                    const double dE = E_final - b.Einit;
                    b.Eprev = b.Einit;                                 // samplers.py:460
                    const bool accepted = (dE < 0) || ((double)b.lnu < -dE);      // samplers.py:462
                    if (accepted) {
                        if (b.it >= a.warm_up_num) n_acc_post++; else n_acc_warm++;
                        b.V = V;
                        b.state = ST_ACC;
                    } else {
                        b.state = ST_REJ;
                    }
                    if (tr) a.decision_chain[b.it - 1] = accepted ? 1 : 0;
                }
                bk[s] = b;
            }
        }
        __syncwarp();
        PH_T(tF);
        PH_ADD(3, tE, tF);
#ifdef HMC_PROFILE_PHASES
        ph[5] += 1;
#endif
    };
    auto any_running = [&]() -> bool { return __ballot_sync(HMC_FULL_MASK, bk[lane].state == ST_RUN) != 0u; };

    do_service();                              // initial fill of the slots
    bool running = any_running();
    if constexpr (PINGPONG) {
        // Warps w and w+4 share a scheduler (warp id mod 4).  Left alone they lock in phase -- both in the
        // gradient loop (sharing the FMA pipe), then both in the latency-bound update/service code (FMA pipe idle;
        // measured: profiles/phase_cycles.py).  A 64-thread named barrier per pair makes them alternate instead:
        // on every tick one warp of the pair runs its gradient loop while the other runs update + service.
        static_assert(WARPS == 8, "ping-pong pairs warps w and w+4");
        const int role = warp >> 2, pair = warp & 3;
        volatile int* done = reinterpret_cast<volatile int*>(done_flags);
        bool have_grad = false;
        for (int tick = 0;; ++tick) {
            PH_T(tT0);
            if (running) {
                if ((tick & 1) == role) { do_gradient(); have_grad = true; }
                else if (have_grad) { do_update(); do_service(); have_grad = false; running = any_running(); }
            }
            // flags double-buffered by tick parity: the partner may already be writing the flag of tick + 1 while this warp
            // still reads the flags of this tick; both warps of a pair decide on the same snapshot
            volatile int* dn = done + (tick & 1) * WARPS;
            if (lane == 0) dn[warp] = (running || have_grad) ? 0 : 1;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");
            PH_T(tT2);
            PH_ADD(4, tT0, tT2);
            if (dn[warp] && dn[warp ^ 4]) break;
        }
    } else {
        while (running) {
            PH_T(tL0);
            do_gradient();
            do_update();
            do_service();
            running = any_running();
            PH_T(tL1);
            PH_ADD(4, tL0, tL1);
        }
    }
#ifdef HMC_PROFILE_PHASES
    if (lane == 0) for (int i = 0; i < 8; ++i) atomicAdd(&g_phase_cycles[i], (unsigned long long)ph[i]);
#endif

    // ---- counters (samplers.py:484-488 numerators; sum L, sum L^2 for N_total_steps) ---------------------------
    if (a.counters) {
        const unsigned long long c0 = warp_sum<unsigned long long>(n_acc_warm), c1 = warp_sum<unsigned long long>(n_acc_post);
        const unsigned long long c2 = warp_sum<unsigned long long>(n_sumL), c3 = warp_sum<unsigned long long>(n_sumL2);
        if (lane == 0) {
            atomicAdd(a.counters + 0, c0); atomicAdd(a.counters + 1, c1);
            atomicAdd(a.counters + 2, c2); atomicAdd(a.counters + 3, c3);
        }
    }
}

template <int TM, int TN, int NDG, int NCG, int WARPS>
size_t fast_smem_bytes(int D) {
    using G = Geo<TM, TN, NDG, NCG>;
    return sizeof(float) * ((size_t)D * G::DP + 2 * G::DP + 32 + (size_t)WARPS * ((size_t)D * G::QS + G::RED + G::STAGE + G::BK));
}

template <int TM, int TN, int NDG, int NCG, int WARPS, bool UDT, bool FULL, bool PINGPONG = false>
int launch_fast(const hmc_random_args& a, cudaStream_t stream) {
    auto kern = hmc_random_fast_kernel<TM, TN, NDG, NCG, WARPS, UDT, FULL, PINGPONG>;
    const size_t smem = fast_smem_bytes<TM, TN, NDG, NCG, WARPS>(a.target.D);
    HMC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0;
    HMC_CUDA_CHECK(cudaGetDevice(&dev));
    HMC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int per_cta = WARPS * Geo<TM, TN, NDG, NCG>::NSLOT;
    int grid = (a.Nchain + per_cta - 1) / per_cta;
    if (grid > sms) grid = sms;               // persistent: one CTA per SM, slots refill from the queue
    unsigned int* queue = (unsigned int*)a.state_g;   // scratch: work-queue head
    HMC_CUDA_CHECK(cudaMemsetAsync(queue, 0, sizeof(unsigned int), stream));
    kern<<<grid, WARPS * 32, smem, stream>>>(a, queue);
    HMC_CUDA_CHECK(cudaGetLastError());
    return HMC_OK;
}

}  // namespace

#ifdef HMC_PROFILE_PHASES
extern "C" int hmc_debug_phase_cycles(unsigned long long* out8, int reset) {
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z)); return 0; }
    cudaMemcpyFromSymbol(out8, g_phase_cycles, sizeof(unsigned long long) * 8);
    return 0;
}
#endif

bool hmc_random_fast_supported(const hmc_random_args& a, const char** why) {
    if (a.dtype != HMC_F32) { *why = "float32 only"; return false; }
    if (a.target.Mit || a.target.Pt || a.target.Lct) { *why = "identity momentum metric only"; return false; }
    if (a.target.D > 128 || a.target.D <= 20) { *why = "20 < D <= 128 in this build"; return false; }
    if (a.iter_end <= a.iter_begin) { *why = "needs at least one iteration"; return false; }
    if (!a.state_g) { *why = "state_g scratch required"; return false; }
    return true;
}

// `uniform_dt`: all entries of target.dt are equal (decided by the host mirror; flags bit 0).
int hmc_random_run_fast(const hmc_random_args& a, cudaStream_t stream) {
    const int D = a.target.D;
    const bool udt = (a.flags & 1) != 0;
    const int variant = (a.flags >> 8) & 0xff;        // tuning knob (0 = default tile)
    // measured on B200, Case 3c, 65,536 chains: 8x10 tile / 8 warps 1.39e9, 4x10 tile / 16 warps 1.38e9,
    // 4x10 tile / 12 warps 1.49e9 gradient evals/s  =>  the 4x10 tile with 12 warps is the default.
    if (D == 100 && udt && variant == 1) return launch_fast<8, 10, 10, 3, 8, true, true>(a, stream);
    if (D == 100 && udt && variant == 2) return launch_fast<8, 10, 10, 3, 8, true, true, true>(a, stream);
    if (D == 100 && udt && variant == 3) return launch_fast<8, 10, 10, 3, 4, true, true>(a, stream);   // probe: 1 warp per scheduler
    if (D == 100 && udt) return launch_fast<4, 10, 10, 3, 12, true, true>(a, stream);
    if (D == 100) return udt ? launch_fast<8, 10, 10, 3, 8, true, true>(a, stream) : launch_fast<8, 10, 10, 3, 8, false, true>(a, stream);
    if (D > 100) return udt ? launch_fast<4, 8, 16, 2, 12, true, false>(a, stream) : launch_fast<4, 8, 16, 2, 12, false, false>(a, stream);
    if (D > 80) return udt ? launch_fast<4, 10, 10, 3, 12, true, false>(a, stream) : launch_fast<4, 10, 10, 3, 12, false, false>(a, stream);
    if (D > 40) return udt ? launch_fast<4, 10, 8, 4, 12, true, false>(a, stream) : launch_fast<4, 10, 8, 4, 12, false, false>(a, stream);
    return udt ? launch_fast<4, 10, 4, 8, 12, true, false>(a, stream) : launch_fast<4, 10, 4, 8, 12, false, false>(a, stream);
}
