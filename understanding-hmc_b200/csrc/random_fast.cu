// Fused random-trajectory HMC kernel, FP32, D <= 128, identity momentum metric (the production path).
//
// Follows HMC_sampler.gen_sample_random + leap_frog (/root/reference/samplers.py:387-491, 831-839) with one
// gradient evaluation per leapfrog step (the second gradient of step l is the first of step l+1).
//
// Work decomposition (measured design probes: profiles/microbench_r1_design_probes.txt):
//   * A WARP is autonomous: it owns NSLOT = NCG*8 chain slots and never synchronises with other warps.
//     lane = (cg, dg): chain group cg (8 chains) x dimension group dg (TN dimensions); the lane keeps the
//     gradient accumulators g[8][TN] and the momenta p[8][TN] of its tile in registers.
//   * The positions live in a per-warp shared-memory tile  Ds[k][slot]  (shifted coordinates d = q - mu), the
//     precision matrix in a CTA-wide tile  Ps[k][j]; the gradient  g[c][j] = sum_k d[k][c] P[k][j]  is an
//     FFMA2 loop: two chains per packed FMA, P[k][j] as the scalar-broadcast operand.
//   * Chains advance asynchronously: every pass of the loop is one gradient + leapfrog update for
//     every slot; a slot whose trajectory ends is serviced (energy, Metropolis accept on a Philox uniform,
//     sample store, momentum refresh, new L) without stalling the others, and a slot whose chain is finished
//     pulls the next chain from a global queue.
// HBM sees only the stored sample / energy stream and the final chain state.
#include "hmc_common.cuh"

namespace {

constexpr int TM = 8;  // chains per lane (4 FFMA2 pairs)

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

template <int TN, int NDG, int NCG>
struct Geo {
    static constexpr int NSLOT = NCG * TM;   // chain slots per warp
    static constexpr int DP = NDG * TN;      // padded dimension
    static constexpr int QS = NSLOT;         // row stride of the position tiles (floats)
    static constexpr int STAGE = (DP + 31) / 32 * 32 + 32;
    static constexpr int RED = 32 * 16;
    static constexpr int WARP_FLOATS = 0;    // computed at run time (depends on D)
};

struct WarpCtx {
    float* Ds;      // [D][QS] current positions (shifted)
    float* D0s;     // [D][QS] positions at the start of the running iteration
    float* red;     // [32][16] per-lane partial sums
    float* stage;   // momentum staging
};

// Momentum refresh for one chain, warp-cooperative: 4 normals per lane (dims 4*lane .. 4*lane+3) into `stage`,
// returns sum p^2 (all lanes) and the scalar draws (L, u) of the iteration (samplers.py:431, 441, 461).
struct GenArgs {
    uint64_t seed;
    const double* p_tape;
    const int32_t* L_tape;
    const double* u_tape;
    int D, Niter, L_low, L_high;
};

__device__ __noinline__ void gen_momentum(const GenArgs a, long m, uint64_t gid, int iter, int lane, float* stage,
                                          double* sumsq, int* L, double* u) {
    const int D = a.D;
    float s = 0.f;
    if (a.p_tape) {
        const double* src = a.p_tape + ((size_t)m * (a.Niter + 1) + iter) * D;
        for (int j = lane; j < D; j += 32) { const float v = (float)src[j]; stage[j] = v; s = fmaf(v, v, s); }
        if (iter >= 1) { *L = a.L_tape[(size_t)m * a.Niter + iter - 1]; *u = a.u_tape[(size_t)m * a.Niter + iter - 1]; }
    } else {
        const int nslot = (D + 3) >> 2;
        for (int sl = lane; sl < nslot; sl += 32) {
            const float4 z = hmc_normal4(a.seed, gid, (uint32_t)iter, (uint32_t)sl);
            const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int j = 4 * sl + r;
                if (j < D) { stage[j] = zz[r]; s = fmaf(zz[r], zz[r], s); }
            }
        }
        if (iter >= 1) hmc_scalar_draws(a.seed, gid, (uint32_t)iter, a.L_low, a.L_high, L, u);
    }
    *sumsq = warp_sum<double>((double)s);
    __syncwarp();
}

template <int TN, int NDG, int NCG, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) hmc_random_fast_kernel(const hmc_random_args a, unsigned int* __restrict__ queue) {
    using G = Geo<TN, NDG, NCG>;
    constexpr int NSLOT = G::NSLOT, DP = G::DP, QS = G::QS;
    extern __shared__ __align__(16) float sm[];
    const int D = a.target.D;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* Ps = sm;                       // [D][DP]
    float* mu_s = Ps + D * DP;            // [DP]
    float* dt_s = mu_s + DP;              // [DP]
    float* wbase = dt_s + DP + (size_t)warp * (2 * D * QS + G::RED + G::STAGE);
    float* Ds = wbase;
    float* D0s = Ds + D * QS;
    float* red = D0s + D * QS;
    float* stage = red + G::RED;
    {   // stage P (transposed force matrix == precision matrix for M = I), mu, dt; zero the padding
        const float* Ft = (const float*)a.target.Ft;
        const int Dpad = a.target.D_pad;
        for (int t = threadIdx.x; t < D * DP; t += blockDim.x) {
            const int k = t / DP, j = t - k * DP;
            Ps[t] = (j < D) ? Ft[(size_t)k * Dpad + j] : 0.f;
        }
        for (int t = threadIdx.x; t < DP; t += blockDim.x) {
            mu_s[t] = (t < D) ? ((const float*)a.target.mu)[t] : 0.f;
            dt_s[t] = (t < D) ? ((const float*)a.target.dt)[t] : 0.f;
        }
        for (int t = lane; t < 2 * D * QS; t += 32) Ds[t] = 0.f;
        __syncthreads();
    }
    const bool active = lane < NCG * NDG;
    const int cg = active ? lane / NDG : 0;
    const int dg = active ? lane % NDG : 0;
    const int j0 = dg * TN;
    const long Lc = 1 + (a.Niter - a.warm_up_num) / a.thin_rate;      // samplers.py:31
    float* q_chain = (float*)a.q_chain;

    float2 g[TM / 2][TN], p[TM / 2][TN];
#pragma unroll
    for (int c = 0; c < TM / 2; ++c)
#pragma unroll
        for (int j = 0; j < TN; ++j) { g[c][j] = make_float2(0.f, 0.f); p[c][j] = make_float2(0.f, 0.f); }

    // ---- per-slot bookkeeping, held by lane s < NSLOT --------------------------------------------------------
    int bk_state = (lane < NSLOT) ? ST_NEED_CHAIN : ST_IDLE;
    long bk_m = -1;
    int bk_it = 0, bk_l = 0, bk_L = 1;
    bool bk_init = false;
    double bk_Einit = 0.0, bk_Eprev = 0.0, bk_K0 = 0.0, bk_Knew = 0.0, bk_u = 0.5, bk_V = 0.0;
    unsigned long long n_acc_warm = 0, n_acc_post = 0, n_sumL = 0, n_sumL2 = 0;
    const double vconst = a.target.v_const;
    GenArgs ga;
    ga.seed = a.seed; ga.p_tape = a.p_tape; ga.L_tape = a.L_tape; ga.u_tape = a.u_tape;
    ga.D = D; ga.Niter = a.Niter; ga.L_low = a.L_low; ga.L_high = a.L_high;

    while (true) {
        // ===== A. service every slot that finished a trajectory or needs a chain (warp-uniform loop) ==========
        unsigned need = __ballot_sync(HMC_FULL_MASK, bk_state >= ST_NEED_CHAIN);
        while (need) {
            const int s = __ffs(need) - 1;
            need &= need - 1;
            const int scg = s / TM, sc = s % TM;
            int kind = __shfl_sync(HMC_FULL_MASK, bk_state, s);
            long m = __shfl_sync(HMC_FULL_MASK, bk_m, s);
            int it = __shfl_sync(HMC_FULL_MASK, bk_it, s);
            const bool owner = active && (cg == scg);
            float dv[TN];
#pragma unroll
            for (int jj = 0; jj < TN; ++jj) dv[jj] = 0.f;
            if (kind == ST_ACC || kind == ST_REJ) {
                // ---- the trajectory of iteration `it` ended: store the sample (samplers.py:462-472)
                const bool keep = it >= a.warm_up_num;
                const long idx = keep ? (it - a.warm_up_num) / a.thin_rate : 0;
                if (owner) {
#pragma unroll
                    for (int jj = 0; jj < TN; ++jj) {
                        const int j = j0 + jj;
                        if (j < D) {
                            if (kind == ST_ACC) { dv[jj] = Ds[j * QS + s]; D0s[j * QS + s] = dv[jj]; }
                            else { dv[jj] = D0s[j * QS + s]; Ds[j * QS + s] = dv[jj]; }
                            if (keep) q_chain[((size_t)m * Lc + idx) * D + j] = dv[jj] + mu_s[j];
                        }
                    }
                }
                if (it >= a.iter_end) {
                    // ---- chain finished: final state out, slot asks for the next chain
                    if (owner) {
#pragma unroll
                        for (int jj = 0; jj < TN; ++jj) { const int j = j0 + jj; if (j < D) ((float*)a.state_q)[(size_t)m * D + j] = dv[jj] + mu_s[j]; }
                    }
                    if (lane == s) a.state_eprev[m] = bk_Eprev;
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
                        for (int jj = 0; jj < TN; ++jj) { const int j = j0 + jj; if (j < D) { Ds[j * QS + s] = 0.f; D0s[j * QS + s] = 0.f; } }
                    }
                    if (lane == s) { bk_state = ST_IDLE; bk_m = -1; }
                    // zero the slot's momenta (static register index)
#pragma unroll
                    for (int c = 0; c < TM; ++c) {
                        if (c == sc && owner) {
#pragma unroll
                            for (int jj = 0; jj < TN; ++jj) { if (c & 1) { p[c / 2][jj].y = 0.f; g[c / 2][jj].y = 0.f; } else { p[c / 2][jj].x = 0.f; g[c / 2][jj].x = 0.f; } }
                        }
                    }
                    __syncwarp();
                    continue;
                }
                m = (long)nxt;
                it = a.iter_begin;
                const float* src = (a.iter_begin == 0) ? (const float*)a.q_start : (const float*)a.state_q;
                if (owner) {
#pragma unroll
                    for (int jj = 0; jj < TN; ++jj) {
                        const int j = j0 + jj;
                        if (j < D) {
                            const float qv = src[(size_t)m * D + j];
                            dv[jj] = qv - mu_s[j];
                            Ds[j * QS + s] = dv[jj];
                            D0s[j * QS + s] = dv[jj];
                            if (a.iter_begin == 0) q_chain[(size_t)m * Lc * D + j] = qv;      // samplers.py:413
                        }
                    }
                }
                if (a.iter_begin == 0) {                                                   // samplers.py:415 (K only)
                    double k0; int Ld; double ud;
                    gen_momentum(ga, m, (uint64_t)(a.chain_id0 + m), 0, lane, stage, &k0, &Ld, &ud);
                    if (lane == s) { bk_K0 = 0.5 * k0; bk_init = true; }
                    if (a.decision_chain && a.chain_id0 + m == 0 && lane == s) a.decision_chain[a.N_save_chain0] = 0;
                } else if (lane == s) {
                    bk_Eprev = a.state_eprev[m];
                    bk_init = false;
                }
            }
            // ---- start iteration it+1: momentum refresh (samplers.py:431), trajectory length (:441), uniform (:461)
            const int itn = it + 1;
            const uint64_t gid = (uint64_t)(a.chain_id0 + m);
            double ksum; int Ln = 1; double un = 0.5;
            gen_momentum(ga, m, gid, itn, lane, stage, &ksum, &Ln, &un);
            const bool go = (kind == ST_ACC);      // gradient at the accepted point is still in g: kick and drift now
            const bool tr = a.phi_q && gid == 0 && itn <= a.N_save_chain0;
            if (lane == s) {
                bk_m = m; bk_it = itn; bk_L = Ln; bk_u = un; bk_Knew = 0.5 * ksum; bk_state = ST_RUN;
                n_sumL += (unsigned long long)Ln; n_sumL2 += (unsigned long long)Ln * Ln;
                if (go) {
                    // E_initial of the new iteration (samplers.py:434-438): V at the accepted point + new kinetic energy
                    bk_Einit = bk_V + bk_Knew;
                    if (itn >= a.warm_up_num) {
                        const long idx = (itn - a.warm_up_num) / a.thin_rate;
                        a.E_chain[(size_t)m * Lc + idx] = bk_Einit;
                        a.dE_chain[(size_t)m * Lc + idx] = bk_Einit - bk_Eprev;
                    }
                    bk_l = 1;
                } else {
                    bk_l = 0;
                }
                if (tr) {
                    double* phi = a.phi_q + (size_t)(itn - 1) * a.L_high * 2;
                    phi[0] = (double)(D0s[s] + mu_s[0]);
                    if (D > 1) phi[1] = (double)(D0s[QS + s] + mu_s[1]);
                    a.phi_len[itn - 1] = Ln + 1;
                }
            }
            // owners take the new momentum into registers (static register index c == sc)
#pragma unroll
            for (int c = 0; c < TM; ++c) {
                if (c == sc && owner) {
#pragma unroll
                    for (int jj = 0; jj < TN; ++jj) {
                        const int j = j0 + jj;
                        float pn = (j < D) ? stage[j] : 0.f;
                        if (go && j < D) {
                            const float gj = (c & 1) ? g[c / 2][jj].y : g[c / 2][jj].x;
                            const float dtj = dt_s[j];
                            pn = fmaf(gj, -0.5f * dtj, pn);                    // first half kick (samplers.py:835)
                            const float dn = fmaf(pn, dtj, dv[jj]);            // drift (samplers.py:836)
                            Ds[j * QS + s] = dn;
                        }
                        if (c & 1) p[c / 2][jj].y = pn; else p[c / 2][jj].x = pn;
                        if (!go) { if (c & 1) g[c / 2][jj].y = 0.f; else g[c / 2][jj].x = 0.f; }
                    }
                }
            }
            __syncwarp();
            if (tr && go && lane == s) {
                double* phi = a.phi_q + (size_t)(itn - 1) * a.L_high * 2;
                phi[2] = (double)(Ds[s] + mu_s[0]);
                if (D > 1) phi[3] = (double)(Ds[QS + s] + mu_s[1]);
            }
        }

        // ===== B. done when no slot runs ====================================================================
        const unsigned run = __ballot_sync(HMC_FULL_MASK, bk_state == ST_RUN);
        if (run == 0u) break;
        // point index of the gradient about to be evaluated: 0 = first point, L = last point of the trajectory
        const unsigned m_l0 = __ballot_sync(HMC_FULL_MASK, bk_state == ST_RUN && bk_l == 0);
        const unsigned m_last = __ballot_sync(HMC_FULL_MASK, bk_state == ST_RUN && bk_l == bk_L);
        const unsigned m_mid = run & ~m_l0 & ~m_last;

        // ===== C. gradient  g[c][j] = sum_k d[k][c] P[k][j]  ==================================================
#pragma unroll
        for (int c = 0; c < TM / 2; ++c)
#pragma unroll
            for (int j = 0; j < TN; ++j) g[c][j] = make_float2(0.f, 0.f);
        {
            const float* qp = Ds + cg * TM;
            const float* pp = Ps + j0;
#pragma unroll 2
            for (int k = 0; k < D; ++k) {
                float qv[TM], pv[TN];
#pragma unroll
                for (int i = 0; i < TM / 4; ++i) *reinterpret_cast<float4*>(&qv[4 * i]) = *reinterpret_cast<const float4*>(qp + k * QS + 4 * i);
                if constexpr (TN % 4 == 0) {
#pragma unroll
                    for (int i = 0; i < TN / 4; ++i) *reinterpret_cast<float4*>(&pv[4 * i]) = *reinterpret_cast<const float4*>(pp + k * DP + 4 * i);
                } else {
#pragma unroll
                    for (int i = 0; i < TN / 2; ++i) *reinterpret_cast<float2*>(&pv[2 * i]) = *reinterpret_cast<const float2*>(pp + k * DP + 2 * i);
                }
#pragma unroll
                for (int c = 0; c < TM / 2; ++c)
#pragma unroll
                    for (int j = 0; j < TN; ++j) g[c][j] = fma2(make_float2(qv[2 * c], qv[2 * c + 1]), bc2(pv[j]), g[c][j]);
            }
        }
        __syncwarp();

        // ===== D. leapfrog update of the lane's tile (samplers.py:835-837) + energy partial sums ================
        // per chain: w2 = 1 for interior points (second half kick of step l and first half kick of step l+1),
        //            wd = 1 where the position moves (every point but the last of a trajectory).
        float2 kw[TM / 2], dw[TM / 2], hv[TM / 2], hk[TM / 2];
#pragma unroll
        for (int c = 0; c < TM / 2; ++c) {
            const int b0 = cg * TM + 2 * c;
            const float mid0 = (float)((m_mid >> b0) & 1u), mid1 = (float)((m_mid >> (b0 + 1)) & 1u);
            const float mv0 = (float)(((m_mid | m_l0) >> b0) & 1u), mv1 = (float)(((m_mid | m_l0) >> (b0 + 1)) & 1u);
            kw[c] = make_float2(-0.5f - 0.5f * mid0, -0.5f - 0.5f * mid1);
            dw[c] = make_float2(mv0, mv1);
            hv[c] = make_float2(0.f, 0.f);
            hk[c] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int jj = 0; jj < TN; ++jj) {
            const int j = j0 + jj;
            const float dtj = dt_s[j];     // 0 for padded dimensions
            float dq[TM];
#pragma unroll
            for (int i = 0; i < TM / 4; ++i)
                *reinterpret_cast<float4*>(&dq[4 * i]) = (j < D) ? *reinterpret_cast<const float4*>(Ds + j * QS + cg * TM + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int c = 0; c < TM / 2; ++c) {
                const float2 dd = make_float2(dq[2 * c], dq[2 * c + 1]);
                const float2 gg = g[c][jj];
                hv[c] = fma2(dd, gg, hv[c]);
                const float2 kc = make_float2(kw[c].x * dtj, kw[c].y * dtj);
                const float2 pn = fma2(gg, kc, p[c][jj]);
                hk[c] = fma2(pn, pn, hk[c]);
                p[c][jj] = pn;
                const float2 dc = make_float2(dw[c].x * dtj, dw[c].y * dtj);
                const float2 dn = fma2(pn, dc, dd);
                dq[2 * c] = dn.x; dq[2 * c + 1] = dn.y;
            }
            if (active && j < D) {
#pragma unroll
                for (int i = 0; i < TM / 4; ++i) *reinterpret_cast<float4*>(Ds + j * QS + cg * TM + 4 * i) = *reinterpret_cast<const float4*>(&dq[4 * i]);
            }
        }
        // ---- reduce the partial sums over the dimension groups (shared memory, fixed order => deterministic)
        if (active) {
            float* r = red + (cg * NDG + dg) * 16;
            *reinterpret_cast<float4*>(r + 0) = make_float4(hv[0].x, hv[0].y, hv[1].x, hv[1].y);
            *reinterpret_cast<float4*>(r + 4) = make_float4(hv[2].x, hv[2].y, hv[3].x, hv[3].y);
            *reinterpret_cast<float4*>(r + 8) = make_float4(hk[0].x, hk[0].y, hk[1].x, hk[1].y);
            *reinterpret_cast<float4*>(r + 12) = make_float4(hk[2].x, hk[2].y, hk[3].x, hk[3].y);
        }
        __syncwarp();

        // ===== E. per-slot bookkeeping (lane s < NSLOT) ========================================================
        if (lane < NSLOT && bk_state == ST_RUN) {
            const int s = lane, scg = s / TM, sc = s % TM;
            double sv = 0.0, sk = 0.0;
#pragma unroll
            for (int d2 = 0; d2 < NDG; ++d2) {
                sv += (double)red[(scg * NDG + d2) * 16 + sc];
                sk += (double)red[(scg * NDG + d2) * 16 + 8 + sc];
            }
            const double V = 0.5 * sv + vconst;                        // V(q) = 0.5 d.P d + const  (utils.py:213-218)
            const bool tr = a.phi_q && (a.chain_id0 + bk_m) == 0 && bk_it <= a.N_save_chain0;
            if (bk_l == 0) {
                // first point of a trajectory reached through a fresh gradient (chain start or after a rejection)
                if (bk_init) {                                         // samplers.py:416-420
                    const double E0 = V + bk_K0;
                    a.E_chain[(size_t)bk_m * Lc] = E0;
                    a.dE_chain[(size_t)bk_m * Lc] = 0.0;
                    bk_Eprev = E0;
                    bk_init = false;
                }
                bk_Einit = V + bk_Knew;                                // samplers.py:434-438
                if (bk_it >= a.warm_up_num) {
                    const long idx = (bk_it - a.warm_up_num) / a.thin_rate;
                    a.E_chain[(size_t)bk_m * Lc + idx] = bk_Einit;
                    a.dE_chain[(size_t)bk_m * Lc + idx] = bk_Einit - bk_Eprev;
                }
                bk_l = 1;
                if (tr) {
                    double* phi = a.phi_q + (size_t)(bk_it - 1) * a.L_high * 2;
                    phi[2] = (double)(Ds[s] + mu_s[0]);
                    if (D > 1) phi[3] = (double)(Ds[QS + s] + mu_s[1]);
                }
            } else if (bk_l == bk_L) {
                // last point: Metropolis accept (samplers.py:455-472)
                const double E_final = V + 0.5 * sk;
                const double dE = E_final - bk_Einit;
                bk_Eprev = bk_Einit;                                   // samplers.py:460
                const double lnu = log(bk_u);
                const bool accepted = (dE < 0) || (lnu < -dE);          // samplers.py:462
                if (accepted) {
                    if (bk_it >= a.warm_up_num) n_acc_post++; else n_acc_warm++;
                    bk_V = V;
                    bk_state = ST_ACC;
                } else {
                    bk_state = ST_REJ;
                }
                if (tr) a.decision_chain[bk_it - 1] = accepted ? 1 : 0;
            } else {
                bk_l += 1;
                if (tr) {
                    double* phi = a.phi_q + (size_t)(bk_it - 1) * a.L_high * 2;
                    phi[2 * bk_l] = (double)(Ds[s] + mu_s[0]);
                    if (D > 1) phi[2 * bk_l + 1] = (double)(Ds[QS + s] + mu_s[1]);
                }
            }
        }
        __syncwarp();
    }

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

template <int TN, int NDG, int NCG, int WARPS>
size_t fast_smem_bytes(int D) {
    using G = Geo<TN, NDG, NCG>;
    return sizeof(float) * ((size_t)D * G::DP + 2 * G::DP + (size_t)WARPS * (2 * (size_t)D * G::QS + G::RED + G::STAGE));
}

template <int TN, int NDG, int NCG, int WARPS>
int launch_fast(const hmc_random_args& a, cudaStream_t stream) {
    auto kern = hmc_random_fast_kernel<TN, NDG, NCG, WARPS>;
    const size_t smem = fast_smem_bytes<TN, NDG, NCG, WARPS>(a.target.D);
    HMC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0;
    HMC_CUDA_CHECK(cudaGetDevice(&dev));
    HMC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int per_cta = WARPS * Geo<TN, NDG, NCG>::NSLOT;
    int grid = (a.Nchain + per_cta - 1) / per_cta;
    if (grid > sms) grid = sms;               // persistent: one CTA per SM, slots refill from the queue
    unsigned int* queue = (unsigned int*)a.state_g;   // scratch: work-queue head
    HMC_CUDA_CHECK(cudaMemsetAsync(queue, 0, sizeof(unsigned int), stream));
    kern<<<grid, WARPS * 32, smem, stream>>>(a, queue);
    HMC_CUDA_CHECK(cudaGetLastError());
    return HMC_OK;
}

}  // namespace

bool hmc_random_fast_supported(const hmc_random_args& a, const char** why) {
    if (a.dtype != HMC_F32) { *why = "float32 only"; return false; }
    if (a.target.Mit || a.target.Pt || a.target.Lct) { *why = "identity momentum metric only"; return false; }
    if (a.target.D > 100 || a.target.D <= 40) { *why = "40 < D <= 100 in this build"; return false; }
    if (a.iter_end <= a.iter_begin) { *why = "needs at least one iteration"; return false; }
    if (!a.state_g) { *why = "state_g scratch required"; return false; }
    return true;
}

int hmc_random_run_fast(const hmc_random_args& a, cudaStream_t stream) {
    const int D = a.target.D;
    if (D > 80) return launch_fast<10, 10, 3, 8>(a, stream);      // 24 chains x 100 dims per warp
    return launch_fast<10, 8, 4, 8>(a, stream);                   // 32 chains x 80 dims per warp
}
