// Chain-tiled tensor-core NUTS kernel (tcgen05 + TMEM), D = 100, float32 state, identity momentum metric.
//
// Follows HMC_sampler.gen_sample_NUTS (/root/reference/samplers.py:495-808): iterative tree doubling with a direction
// coin per doubling, uniform progressive sampling inside the new sub-trajectory, sub-tree U-turn checks against saved odd
// points, biased old/new choice, termination when BOTH ends have turned (Q6), d_max overflow (Q7), the |dE| > 1000 guard.
//
// One CTA = 128 chains sharing the shared-memory resident split force matrix; every PASS evaluates the gradient of all 128
// chains as one tcgen05 GEMM (the machinery of random_tc.cu: A operand = split positions in tensor memory, B = split F in
// shared memory, fp32 accumulator in tensor memory).  FOUR threads share a chain (24/24/24/28 dimensions); position and
// momentum slices stay in registers.  Each chain is a state machine that advances by one leapfrog step per pass:
//     phase STEP        the gradient is that of the new point: second half kick, energy, the reference's per-point logic
//     phase ITER_START  gradient at the start point of an iteration (fresh momentum): E_initial, first doubling
//     phase CHAIN_START the same for a new chain (also records E_chain[., 0])
//     phase GRAD        gradient at the trajectory end a doubling switched to
//     phase FETCH       the slot takes its next chain from the global queue
// The four slice threads of a chain run the SAME scalar state machine redundantly (per-chain scalars in float64),
// so there is no command traffic between them: they only exchange partial dot products through shared memory, at two fixed
// points of every pass (a 32-chain group = four warps = one named barrier):
//     exchange 1: q.g, p.p and, at even points, (q - q_chk).p and (q - q_chk).p_chk for every check point of the point
//     exchange 2: the whole-trajectory test dots at the end of a doubling, or |p|^2 of a new chain's first momentum.
// The reference's slot table is replaced by the closed forms of nuts_generic.cu (checks at even m against l = m - 2^j + 1,
// saved point l in slot popcount((l-1) >> 1)); the check-point stack and the live / boundary points are rows of the per-chain
// scratch in HBM / L2, every thread reading and writing only its own dimension slice of a row.
// Draws: injected tapes consumed in order per chain, or Philox keyed by position (iteration, depth, step) -- the same keys as
// nuts_generic.cu, so both kernels see the same draws.
#include "tc_common.cuh"
#include <cstdlib>
#include <cstdio>

namespace {

enum : int { NP_IDLE = 0, NP_FETCH = 1, NP_CHAIN_START = 2, NP_ITER_START = 3, NP_STEP = 4, NP_GRAD = 5 };
enum { HMC_STREAM_NUTS_INNER = 3 };
constexpr int NT_MAXCHK = 12;        // check points per even point <= depth of the sub-trajectory <= d_max - 1
constexpr int NT_S1 = TC_THREADS + 32;     // S1: the 512 workers arrive, the issuing warp waits (its three sibling warps only gave their registers)
// Barriers in their non-".aligned" PTX forms (barrier.sync / barrier.arrive count threads and do not require a converged
// warp).  The worker loop is full of per-lane branches (32 chains in different phases per warp); with the aligned forms
// (bar.sync) a warp was observed to get one group barrier out of step with its three sibling warps as soon as one of its
// chains ended while the others kept running -- a deadlock between the group barrier and the wait for the gradient pass
// (found with the stage markers below; profiles/r2_nuts_tc_barrier_note.md).
__device__ __forceinline__ void s1_sync() { asm volatile("barrier.sync 5, %0;" ::"n"(NT_S1) : "memory"); }
__device__ __forceinline__ void s1_arrive() { asm volatile("barrier.arrive 5, %0;" ::"n"(NT_S1) : "memory"); }
__device__ __forceinline__ void grp_sync(int grp) { asm volatile("barrier.sync %0, 128;" ::"r"(1 + grp) : "memory"); }

// exp() of an energy difference: float32 expf (2 ulp) inside its range, float64 beyond (the weights overflow float32
// at |dE| > 88; the guard allows 1000)
__device__ __forceinline__ double nt_exp(double v) { return fabs(v) < 80.0 ? (double)expf((float)v) : exp(v); }

struct NutsTcShared {
    float2 ex[TC_SPL][TC_M];                    // exchange 1: (q.g, p.p) per slice
    float2 chk[TC_SPL][NT_MAXCHK][TC_M];        // exchange 1: ((q - q_chk).p, (q - q_chk).p_chk) per slice and check point
    float2 ex2[TC_SPL][TC_M];                   // exchange 2
    int newm[TC_M];                             // chain handed to the slot by its slice-0 thread (-1: none left)
    int galive[2][4];                           // by pass parity and group: some slot still has (or wants) a chain
    int stop;
    unsigned int fmax_bits;
    int mark[20];                               // HMC_NUTS_TC_DEBUG: last stage marker of every warp
};

// -DHMC_NUTS_TC_DEBUG: stage markers of every warp in shared memory, printed by the wait watchdog (which warp stopped where)
#ifdef HMC_NUTS_TC_DEBUG
#define NT_MARK(v) do { if (lane == 0) *reinterpret_cast<volatile int*>(&sh->mark[warp]) = (v); } while (0)
#else
#define NT_MARK(v)
#endif
// -DHMC_PROFILE_PHASES: cycles per phase of the worker loop, summed over warps (hmc_debug_nuts_tc_cycles)
#ifdef HMC_PROFILE_PHASES
__device__ unsigned long long g_nt_cycles[8];
#define NP_T(x) const unsigned int x = (unsigned int)clock()
#define NP_ADD(i, a, b) tph[i] += (b) - (a)
#else
#define NP_T(x)
#define NP_ADD(i, a, b)
#endif

template <int PREC>
__global__ void __launch_bounds__(TC_NT, 1) hmc_nuts_tc_kernel(const hmc_nuts_args a, unsigned int* __restrict__ queue) {
    constexpr int D = TC_ND, KP = TC_KP, KC = TC_KC;
    constexpr int NPART = TcPrec<PREC>::NPART;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* Bp = smem;
    float* mu_s = reinterpret_cast<float*>(Bp + NPART * TC_BPART);
    float* dt_s = mu_s + KP;
    NutsTcShared* sh = reinterpret_cast<NutsTcShared*>(dt_s + KP);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sh + 1);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = warp & 3, slice = warp >> 2, chain = grp * 32 + lane;
    const int j0 = 24 * (slice & 3);
    const bool wide = slice == TC_SPL - 1;
    const int nch4 = wide ? 7 : 6;

    // ---- one-time set-up (as random_tc.cu) -----------------------------------------------------------------------------
    float binv = 1.f, bscale = 1.f;
    const float* Ftg = (const float*)a.target.Ft;
    const int Dpad = a.target.D_pad;
    if constexpr (TcPrec<PREC>::F16) {
        if (tid == 0) sh->fmax_bits = 0u;
        __syncthreads();
        unsigned int mx = 0u;
        for (int t = tid; t < D * D; t += TC_NT) mx = max(mx, __float_as_uint(fabsf(Ftg[(size_t)(t / D) * Dpad + (t % D)])));
        mx = __reduce_max_sync(HMC_FULL_MASK, mx);
        if (lane == 0) atomicMax(&sh->fmax_bits, mx);
        __syncthreads();
        const int e = (int)(sh->fmax_bits >> 23) - 127;
        int sft = (sh->fmax_bits == 0u || e < -100) ? 0 : 14 - e;
        sft = sft > 100 ? 100 : (sft < -100 ? -100 : sft);
        bscale = __uint_as_float((unsigned int)(127 + sft) << 23);
        binv = __uint_as_float((unsigned int)(127 - sft) << 23);
    }
    for (int t = tid; t < KC * KP; t += TC_NT) {
        const int kc = t / KP, n = t % KP;
        uint32_t w1[4], w2[4], w3[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k0 = kc * 8 + 2 * e, k1 = k0 + 1;
            const float x0 = (n < D && k0 < D) ? Ftg[(size_t)k0 * Dpad + n] * bscale : 0.f;
            const float x1 = (n < D && k1 < D) ? Ftg[(size_t)k1 * Dpad + n] * bscale : 0.f;
            split_pair<PREC>(x0, x1, w1[e], w2[e], w3[e]);
            w2[e] ^= 0x80008000u;
        }
        reinterpret_cast<uint4*>(Bp)[t] = make_uint4(w1[0], w1[1], w1[2], w1[3]);
        reinterpret_cast<uint4*>(Bp + TC_BPART)[t] = make_uint4(w2[0], w2[1], w2[2], w2[3]);
        if constexpr (NPART == 3) reinterpret_cast<uint4*>(Bp + 2 * TC_BPART)[t] = make_uint4(w3[0], w3[1], w3[2], w3[3]);
    }
    for (int t = tid; t < KP; t += TC_NT) {
        mu_s[t] = (t < D) ? ((const float*)a.target.mu)[t] : 0.f;
        dt_s[t] = (t < D) ? ((const float*)a.target.dt)[t] : 0.f;
    }
    if (tid < 8) sh->galive[tid >> 2][tid & 3] = 1;
    if (tid == 0) {
        sh->stop = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmem_row = tmem + ((uint32_t)(grp * 32) << 16) + (uint32_t)j0;
    const uint32_t acol = tmem + ((uint32_t)(grp * 32) << 16) + TC_ACOL + 12u * (uint32_t)(slice & 3);
    constexpr uint32_t fmt = TcPrec<PREC>::F16 ? 0u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(KP >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

    if (slice >= TC_SPL) {
        // ===== issuing warpgroup: after every S1 barrier its first warp launches the gradient pass of the rows as they stand =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (warp == TC_THREADS / 32) {
            int pn = 0;
            while (true) {
                asm volatile("tcgen05.fence::before_thread_sync;");
                s1_sync();
                const volatile int* ga_ = sh->galive[pn & 1];        // posted by the workers before they arrived at this S1
                if ((ga_[0] | ga_[1] | ga_[2] | ga_[3]) == 0) {
                    if (lane == 0) {
                        *reinterpret_cast<volatile int*>(&sh->stop) = 1;
                        __threadfence_block();
                        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(mbar)) : "memory");
                    }
                    break;
                }
                ++pn;
                if (lane == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint64_t dsc = make_desc(smem_u32(Bp), KP * 16, 128);
                    tc_mma_all<PREC, 0>(tmem, (uint32_t)dsc, (uint32_t)(dsc >> 32), idesc);
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
                }
                __syncwarp();
            }
        }
        __syncthreads();
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");

    // ---- per-thread state --------------------------------------------------------------------------------------------------
    float p[28], x[28];
#pragma unroll
    for (int j = 0; j < 28; ++j) { p[j] = 0.f; x[j] = 0.f; }
    // per-chain scalars, identical in the four slice threads of the chain
    int ph = NP_FETCH;               // phase of the gradient that the next pass delivers
    int m = -1, it = 0, depth = 0, k = 0, u_dir = 0, L_sub = 1;
    bool left_term = false, right_term = false, fetch_posted = false;
    double E_initial = 0.0, E_prev = 0.0, E_max_old = 0.0, E_max_new = 0.0, pi_old = 1.0, pi_new = 1.0, u_biased = 0.0;
    long n_dir = 0, n_u = 0;
    unsigned long long c_leap = 0, c_doubling = 0, c_instab = 0, c_dmax = 0;     // accumulated by the slice-0 thread only
    long chain_leap = 0;
    uint32_t mphase = 0;
    bool have_grad = false;
    int pn = 0;
    const long Lc = 1 + (a.Niter - a.warm_up_num) / a.thin_rate;
    const int R = 2 * (a.d_max + 1);
    const double vconst = a.target.v_const;
    float* q_chain = (float*)a.q_chain;

    // check-point stack, live and boundary points: rows of the scratch block of this SLOT (not of the chain: the resident slots'
    // 148 x 128 x 11.6 KB = 220 MB are re-used by every chain a slot runs, which keeps more of them in the 126 MB L2 than
    // Nchain x 11.6 KB would); the cursor row R + 6 stays per chain (resume state of the injected draw streams)
    const size_t slot_id = (size_t)blockIdx.x * TC_M + chain;
    auto scr_row = [&](int row) -> float* { return (float*)a.scratch + (slot_id * (R + 7) + row) * Dpad + j0; };
    auto row_store = [&](int row, const float* v, float sgn) {
        float* dst = scr_row(row);
#pragma unroll
        for (int c = 0; c < 7; ++c)
            if (c < nch4) *reinterpret_cast<float4*>(dst + 4 * c) = make_float4(sgn * v[4 * c], sgn * v[4 * c + 1], sgn * v[4 * c + 2], sgn * v[4 * c + 3]);
    };
    auto row_load = [&](int row, float* v) {
        const float* src = scr_row(row);
#pragma unroll
        for (int c = 0; c < 7; ++c)
            if (c < nch4) { const float4 t = __ldcg(reinterpret_cast<const float4*>(src + 4 * c)); v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w; }
    };
    auto draw_p = [&](int iter) -> float {       // my slice of the momentum of (chain, iteration); returns the slice's |p|^2
        float s = 0.f;
        if (a.p_tape) {
            const double* src = a.p_tape + ((size_t)m * (a.Niter + 1) + iter) * D + j0;
#pragma unroll
            for (int j = 0; j < 28; ++j) if (j < 24 || wide) { p[j] = (float)src[j]; s = fmaf(p[j], p[j], s); }
        } else {
            const uint64_t gid = (uint64_t)(a.chain_id0 + m);
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                if (c < nch4) {
                    const float4 z = hmc_normal4(a.seed, gid, (uint32_t)iter, (uint32_t)((j0 >> 2) + c));
                    p[4 * c] = z.x; p[4 * c + 1] = z.y; p[4 * c + 2] = z.z; p[4 * c + 3] = z.w;
                    s += z.x * z.x + z.y * z.y + z.z * z.z + z.w * z.w;
                }
            }
        }
        return s;
    };
    auto draw_dir = [&](int iter, int dep, double* ub) -> int {
        if (a.dir_tape) return a.dir_tape[(size_t)m * a.tape_dir_stride + n_dir++];
        const uint64_t gid = (uint64_t)(a.chain_id0 + m);
        const Philox4 r = philox4x32_10((uint32_t)gid, (uint32_t)iter, (uint32_t)dep, HMC_STREAM_NUTS | ((uint32_t)(gid >> 32) << 8),
                                        (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
        *ub = ((double)(r.y >> 8) + 0.5) * 5.9604644775390625e-08;
        return (int)(r.x >> 31);
    };
    auto draw_u_inner = [&](int iter, int dep, int kk) -> double {
        if (a.u_tape) return a.u_tape[(size_t)m * a.tape_u_stride + n_u++];
        const uint64_t gid = (uint64_t)(a.chain_id0 + m);
        const Philox4 r = philox4x32_10((uint32_t)gid, (uint32_t)iter, ((1u << dep) + (uint32_t)kk) >> 2,
                                        HMC_STREAM_NUTS_INNER | ((uint32_t)(gid >> 32) << 8), (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
        const uint32_t w = (kk & 3) == 0 ? r.x : (kk & 3) == 1 ? r.y : (kk & 3) == 2 ? r.z : r.w;
        return ((double)(w >> 8) + 0.5) * 5.9604644775390625e-08;
    };
    // first half kick + drift of the next leapfrog step from the gradient in tensor memory (samplers.py:835-836).  The
    // tensor-memory load is a warp-wide instruction: every lane executes it, `on` says whose chain actually steps.
    auto advance = [&](bool on) {
        uint32_t gv[16];
        tmem_ld16(tmem_row, gv);
        if (on) {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const float dtj = dt_s[j0 + jj];
                p[jj] = fmaf(__uint_as_float(gv[jj]), -0.5f * binv * dtj, p[jj]);
                x[jj] = fmaf(p[jj], dtj, x[jj]);
            }
        }
        __syncwarp();
        tmem_ld16(tmem_row + 16u, gv);
        if (on) {
#pragma unroll
            for (int jj = 16; jj < 28; ++jj) {
                if (jj < 24 || wide) {
                    const float dtj = dt_s[j0 + jj];
                    p[jj] = fmaf(__uint_as_float(gv[jj - 16]), -0.5f * binv * dtj, p[jj]);
                    x[jj] = fmaf(p[jj], dtj, x[jj]);
                }
            }
        }
        __syncwarp();
    };

    {   // defined operand rows before the first pass
        put_half0<PREC>(acol, x);
        put_half1<PREC>(acol, x, true);
        if (wide) {
            const uint32_t z[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int pt = 0; pt < NPART; ++pt) { tmem_st4(acol + 14 + pt * TC_APITCH, z); tmem_st2(acol + 18 + pt * TC_APITCH, z); }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }

#ifdef HMC_PROFILE_PHASES
    unsigned int tph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    while (true) {
        NP_T(tA);
        // ===== A. the gradient pass issued after the previous S1 ===========================================================
        if (have_grad) {
            uint32_t done = 0;
            const long long t0 = clock64();
            while (!done) {
                asm volatile("{\n\t.reg .pred pw;\n\tmbarrier.try_wait.parity.shared::cta.b64 pw, [%1], %2;\n\tselp.u32 %0, 1, 0, pw;\n\t}"
                             : "=r"(done) : "r"(smem_u32(mbar)), "r"(mphase) : "memory");
                if (!done && clock64() - t0 > 8000000000ll) {              // a lost completion: abort instead of hanging the GPU
#ifdef HMC_NUTS_TC_DEBUG
                    if (lane == 0) {
                        printf("[nuts_tc watchdog] block %d warp %d pass %d; markers:", blockIdx.x, warp, pn);
                        for (int i = 0; i < 20; ++i) printf(" %d", *reinterpret_cast<volatile int*>(&sh->mark[i]));
                        printf("\n");
                    }
#endif
                    __trap();
                }
            }
            mphase ^= 1u;
            __syncwarp();
            asm volatile("tcgen05.fence::after_thread_sync;");
            if (*reinterpret_cast<volatile int*>(&sh->stop)) break;
        }
        NT_MARK(pn * 100 + 10);
        NP_T(tB);
        NP_ADD(0, tA, tB);
        // ===== B. consume the gradient: second half kick (phase STEP), partial energies, partial sub-tree check dots ==========
        const int pt = k + 1;                                   // number of the point the step reached (phase STEP)
        int nchk = 0;
        {
            float hv = 0.f, hk = 0.f;
            const bool stepping = ph == NP_STEP;                // (tensor memory is undefined before the first pass: no 0 * garbage)
            const float kw = -0.5f * binv;                      // samplers.py:837
            uint32_t gv[16];
            tmem_ld16(tmem_row, gv);
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const float gj = __uint_as_float(gv[jj]);
                hv = fmaf(x[jj], gj, hv);
                if (stepping) p[jj] = fmaf(gj, kw * dt_s[j0 + jj], p[jj]);
                hk = fmaf(p[jj], p[jj], hk);
            }
            __syncwarp();
            tmem_ld16(tmem_row + 16u, gv);
#pragma unroll
            for (int jj = 16; jj < 28; ++jj) {
                if (jj < 24 || wide) {
                    const float gj = __uint_as_float(gv[jj - 16]);
                    hv = fmaf(x[jj], gj, hv);
                    if (stepping) p[jj] = fmaf(gj, kw * dt_s[j0 + jj], p[jj]);
                    hk = fmaf(p[jj], p[jj], hk);
                }
            }
            sh->ex[slice][chain] = make_float2(hv, hk);
            if (ph == NP_STEP) {
                if (pt & 1) {                                   // odd point (1 included): save (samplers.py:623-626, 654-658)
                    const int slot = __popc((unsigned)(pt - 1) >> 1);
                    row_store(slot, x, 1.f);
                    row_store(a.d_max + 1 + slot, p, 1.f);
                } else {                                        // even point: sub-tree U-turn checks (samplers.py:699-736)
                    const int tz = __ffs(pt) - 1;
                    for (int j = tz; j >= 1; --j, ++nchk) {
                        const int l = pt - (1 << j) + 1;
                        const int slot = __popc((unsigned)(l - 1) >> 1);
                        float s_cur = 0.f, s_chk = 0.f;
                        const float* qs = scr_row(slot);
                        const float* ps = scr_row(a.d_max + 1 + slot);
#pragma unroll
                        for (int c = 0; c < 7; ++c) {
                            if (c < nch4) {
                                const float4 qc = __ldcg(reinterpret_cast<const float4*>(qs + 4 * c));
                                const float4 pc = __ldcg(reinterpret_cast<const float4*>(ps + 4 * c));
                                const float d0 = x[4 * c] - qc.x, d1 = x[4 * c + 1] - qc.y, d2 = x[4 * c + 2] - qc.z, d3 = x[4 * c + 3] - qc.w;
                                s_cur = fmaf(d0, p[4 * c], fmaf(d1, p[4 * c + 1], fmaf(d2, p[4 * c + 2], fmaf(d3, p[4 * c + 3], s_cur))));
                                s_chk = fmaf(d0, pc.x, fmaf(d1, pc.y, fmaf(d2, pc.z, fmaf(d3, pc.w, s_chk))));
                            }
                        }
                        sh->chk[slice][nchk][chain] = make_float2(s_cur, s_chk);
                    }
                }
            }
        }
        NT_MARK(pn * 100 + 20);
        NP_T(tB2);
        NP_ADD(1, tB, tB2);
        __syncwarp();
        grp_sync(grp);
        NT_MARK(pn * 100 + 30);
        NP_T(tC);
        NP_ADD(2, tB2, tC);
        // ===== C. the chain's state machine, run identically by its four slice threads ===================================
        float r2a = 0.f, r2b = 0.f;                             // my partials for exchange 2
        bool finish = false, new_doubling = false, wait_k = false, traj_test = false;
        double V = 0.0, Kc = 0.0;
        if (ph >= NP_CHAIN_START) {
            float sv = 0.f, sk = 0.f;
#pragma unroll
            for (int s2 = 0; s2 < TC_SPL; ++s2) { const float2 r = sh->ex[s2][chain]; sv += r.x; sk += r.y; }
            V = 0.5 * (double)(sv * binv) + vconst;             // utils.py:213-218
            Kc = 0.5 * (double)sk;
        }
        if (ph == NP_FETCH && fetch_posted) {
            // the slice-0 thread took a chain from the queue at the end of the previous pass and posted it
            const int nm = sh->newm[chain];
            fetch_posted = false;
            if (nm < 0) { m = -1; ph = NP_IDLE; }
            else {
                m = nm;
                it = a.iter_begin + 1;
                chain_leap = 0;
                const bool fresh = a.iter_begin == 0;
                const float* src = (fresh ? (const float*)a.q_start : (const float*)a.state_q) + (size_t)m * D + j0;
#pragma unroll
                for (int c = 0; c < 7; ++c) {
                    if (c < nch4) {
                        const float4 v = __ldcg(reinterpret_cast<const float4*>(src + 4 * c));
                        if (fresh) *reinterpret_cast<float4*>(q_chain + (size_t)m * Lc * D + j0 + 4 * c) = v;     // samplers.py:548
                        const float4 mu4 = *reinterpret_cast<const float4*>(mu_s + j0 + 4 * c);
                        x[4 * c] = v.x - mu4.x; x[4 * c + 1] = v.y - mu4.y; x[4 * c + 2] = v.z - mu4.z; x[4 * c + 3] = v.w - mu4.w;
                    }
                }
                n_dir = 0; n_u = 0;
                if (fresh) { draw_p(0); ph = NP_CHAIN_START; }   // samplers.py:549-555
                else {
                    E_prev = a.state_eprev[m];
                    if (a.dir_tape) {
                        const long long* cur = reinterpret_cast<const long long*>((float*)a.scratch + ((size_t)m * (R + 7) + R + 6) * Dpad);
                        n_dir = (long)cur[0]; n_u = (long)cur[1];
                    }
                    draw_p(it);
                    ph = NP_ITER_START;
                }
            }
        } else if (ph == NP_CHAIN_START) {
            const double E0 = V + Kc;                           // samplers.py:551-555
            if (slice == 0) { a.E_chain[(size_t)m * Lc] = E0; a.dE_chain[(size_t)m * Lc] = 0.0; }
            E_prev = E0;
            r2a = draw_p(it);                                   // samplers.py:565: the momentum of the first iteration
            wait_k = true;                                      // its |p|^2 comes back through exchange 2
        } else if (ph == NP_ITER_START) {
            E_initial = V + Kc;                                 // samplers.py:569
            new_doubling = true;
        } else if (ph == NP_STEP) {
            if (slice == 0) { c_leap++; chain_leap++; }
            const double E_tmp = V + Kc;                        // samplers.py:618, 643
            bool reject = false;
            if (k == 0) {                                       // first point of the new sub-trajectory (samplers.py:611-626)
                row_store(R + 0, x, 1.f);
                E_max_new = E_tmp;
                pi_new = 1.0;
            } else {
                if (fabs(E_tmp - E_initial) > 1000.0) {         // samplers.py:647-651
                    reject = true;
                    if (slice == 0) c_instab++;
                } else {
                    for (int c = 0; c < nchk; ++c) {            // in the reference's order: j = tz(pt) .. 1
                        float a_cur = 0.f, a_chk = 0.f;
#pragma unroll
                        for (int s2 = 0; s2 < TC_SPL; ++s2) { const float2 r = sh->chk[s2][c][chain]; a_cur += r.x; a_chk += r.y; }
                        if (a_cur < 0.f && a_chk < 0.f) { reject = true; break; }        // samplers.py:722-732 (Q6)
                    }
                    if (!reject) {                              // uniform progressive sampling (samplers.py:743-751)
                        const double E_prev_max = E_max_new;
                        E_max_new = fmax(E_prev_max, E_tmp);
                        const double ex = nt_exp(fabs(E_tmp - E_prev_max));
                        const bool new_max = E_tmp > E_prev_max;
                        const double numer = new_max ? 1.0 : ex;
                        pi_new = numer + (new_max ? ex : 1.0) * pi_new;
                        const double r = numer / pi_new;
                        const double u = draw_u_inner(it, depth, k);
                        if (u < r) row_store(R + 0, x, 1.f);
                    }
                }
            }
            if (reject) finish = true;                          // samplers.py:754-755: the sample stays live_old
            else if (k == L_sub - 1) {
                // the new sub-trajectory is complete: extend the end (samplers.py:758-761), biased choice (:766-776, Q8)
                if (u_dir == 0) { row_store(R + 4, x, 1.f); row_store(R + 5, p, 1.f); }
                else { row_store(R + 2, x, 1.f); row_store(R + 3, p, 1.f); }
                const double rb = nt_exp(-(E_max_new - E_max_old)) * pi_old / pi_new;
                const double E_max_old_prev = E_max_old;
                E_max_old = fmax(E_max_old_prev, E_max_new);
                pi_old = nt_exp(-(E_max_new - E_max_old)) * pi_new + nt_exp(-(E_max_old_prev - E_max_old)) * pi_old;
                const double A = fmin(1.0, rb);
                const double ub = a.u_tape ? a.u_tape[(size_t)m * a.tape_u_stride + n_u++] : u_biased;
                if (ub < A) {                                   // live_old <- live_new (each thread moves its own slice)
                    const float* src = scr_row(R + 0);
                    float* dst = scr_row(R + 1);
#pragma unroll
                    for (int c = 0; c < 7; ++c)
                        if (c < nch4) *reinterpret_cast<float4*>(dst + 4 * c) = __ldcg(reinterpret_cast<const float4*>(src + 4 * c));
                }
                // whole-trajectory U-turn test (samplers.py:779-781): the other end comes from the scratch rows
                {
                    const float* oq = scr_row(u_dir == 0 ? R + 2 : R + 4);
                    const float* op = scr_row(u_dir == 0 ? R + 3 : R + 5);
                    float s_r = 0.f, s_l = 0.f;
#pragma unroll
                    for (int c = 0; c < 7; ++c) {
                        if (c < nch4) {
                            const float4 q4 = __ldcg(reinterpret_cast<const float4*>(oq + 4 * c));
                            const float4 p4 = __ldcg(reinterpret_cast<const float4*>(op + 4 * c));
                            const float oqv[4] = {q4.x, q4.y, q4.z, q4.w}, opv[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                // dq = right_q - left_q; s_r = dq . right_p; s_l = -dq . left_p
                                const float dq = (u_dir == 0) ? x[4 * c + e] - oqv[e] : oqv[e] - x[4 * c + e];
                                const float rp = (u_dir == 0) ? p[4 * c + e] : opv[e];
                                const float lp = (u_dir == 0) ? opv[e] : p[4 * c + e];
                                s_r = fmaf(dq, rp, s_r);
                                s_l = fmaf(-dq, lp, s_l);
                            }
                        }
                    }
                    r2a = s_r; r2b = s_l;
                }
                traj_test = true;
            } else {
                k += 1;                                         // next point of the sub-trajectory: first half kick + drift below
            }
        }
        sh->ex2[slice][chain] = make_float2(r2a, r2b);
        NT_MARK(pn * 100 + 40);
        NP_T(tC2);
        NP_ADD(3, tC, tC2);
        __syncwarp();
        grp_sync(grp);
        NT_MARK(pn * 100 + 50);
        NP_T(tD);
        NP_ADD(4, tC2, tD);
        // ===== D. second half of the state machine: results of exchange 2, trajectory ends, doublings, iteration ends ======
        {
            float e2a = 0.f, e2b = 0.f;
#pragma unroll
            for (int s2 = 0; s2 < TC_SPL; ++s2) { const float2 r = sh->ex2[s2][chain]; e2a += r.x; e2b += r.y; }
            if (wait_k) {                                       // chain start: E_initial with the first iteration's momentum
                E_initial = V + 0.5 * (double)e2a;
                new_doubling = true;
            }
            if (traj_test) {
                right_term = e2a < 0.f;
                left_term = e2b < 0.f;
                depth += 1;                                     // samplers.py:784
                if (left_term && right_term) finish = true;     // samplers.py:595 (Q6: both ends)
                else new_doubling = true;
            }
            bool start_iter = (ph == NP_ITER_START) || wait_k;
            if (start_iter) {
                // samplers.py:571-594: stored energies, live point, both boundary points, running maxima
                const bool keep = it >= a.warm_up_num;
                if (keep && slice == 0) {
                    const long idx = (it - a.warm_up_num) / a.thin_rate;
                    a.E_chain[(size_t)m * Lc + idx] = E_initial;
                    a.dE_chain[(size_t)m * Lc + idx] = E_initial - E_prev;
                }
                row_store(R + 1, x, 1.f);
                row_store(R + 2, x, 1.f);
                row_store(R + 3, p, -1.f);
                row_store(R + 4, x, 1.f);
                row_store(R + 5, p, 1.f);
                E_max_old = E_initial; pi_old = 1.0;
                left_term = false; right_term = false;
                depth = 0;
            }
            int next_ph = ph;
            bool do_advance = false;
            if (new_doubling && !finish) {
                if (depth > a.d_max - 1) {                      // samplers.py:596-598 (Q7)
                    if (slice == 0) { c_dmax++; if (a.status) a.status[m] |= 1; }
                    finish = true;
                } else {
                    if (slice == 0) c_doubling++;
                    L_sub = 1 << depth;                         // samplers.py:604
                    const int prev_dir = u_dir;
                    u_dir = draw_dir(it, depth, &u_biased);     // samplers.py:608
                    k = 0;
                    if (start_iter) {
                        // both ends are the current point: the left end carries -p (samplers.py:577-594, 611-614)
                        if (u_dir == 1) {
#pragma unroll
                            for (int j = 0; j < 28; ++j) p[j] = -p[j];
                        }
                        do_advance = true; next_ph = NP_STEP;
                    } else if (u_dir == prev_dir) {
                        do_advance = true; next_ph = NP_STEP;   // the registers hold this end and its gradient is the one in tensor memory
                    } else {
                        row_load(u_dir == 0 ? R + 4 : R + 2, x);
                        row_load(u_dir == 0 ? R + 5 : R + 3, p);
                        next_ph = NP_GRAD;                      // its gradient first
                    }
                }
            } else if (ph == NP_GRAD) {
                do_advance = true; next_ph = NP_STEP;
            } else if (ph == NP_STEP && !finish && !traj_test) {
                do_advance = true;                              // k was advanced in C
            }
            if (finish) {
                // samplers.py:787-791: E_previous, the stored sample is live_old
                E_prev = E_initial;
                row_load(R + 1, x);
                const bool keep = it >= a.warm_up_num;
                if (keep) {
                    float* dst = q_chain + ((size_t)m * Lc + (it - a.warm_up_num) / a.thin_rate) * D + j0;
#pragma unroll
                    for (int c = 0; c < 7; ++c) {
                        if (c < nch4) {
                            const float4 mu4 = *reinterpret_cast<const float4*>(mu_s + j0 + 4 * c);
                            *reinterpret_cast<float4*>(dst + 4 * c) = make_float4(x[4 * c] + mu4.x, x[4 * c + 1] + mu4.y, x[4 * c + 2] + mu4.z, x[4 * c + 3] + mu4.w);
                        }
                    }
                }
                if (it >= a.iter_end) {
                    // chain done: resume state, per-chain records; the slot asks for its next chain
                    float* dst = (float*)a.state_q + (size_t)m * D + j0;
#pragma unroll
                    for (int c = 0; c < 7; ++c) {
                        if (c < nch4) {
                            const float4 mu4 = *reinterpret_cast<const float4*>(mu_s + j0 + 4 * c);
                            *reinterpret_cast<float4*>(dst + 4 * c) = make_float4(x[4 * c] + mu4.x, x[4 * c + 1] + mu4.y, x[4 * c + 2] + mu4.z, x[4 * c + 3] + mu4.w);
                        }
                    }
                    if (slice == 0) {
                        a.state_eprev[m] = E_prev;
                        long long* cur = reinterpret_cast<long long*>((float*)a.scratch + ((size_t)m * (R + 7) + R + 6) * Dpad);
                        cur[0] = n_dir; cur[1] = n_u;
                        if (a.n_leapfrog) a.n_leapfrog[m] += (int64_t)chain_leap;
                    }
                    m = -1;
                    next_ph = NP_FETCH;
#pragma unroll
                    for (int j = 0; j < 28; ++j) { x[j] = 0.f; p[j] = 0.f; }
                } else {
                    it += 1;
                    draw_p(it);                                 // samplers.py:565
                    next_ph = NP_ITER_START;
                }
            }
            ph = next_ph;
            NT_MARK(pn * 100 + 61);
            __syncwarp();
            NT_MARK(pn * 100 + 62);
            advance(do_advance);
        }
        NT_MARK(pn * 100 + 63);
        // a slot in phase FETCH: its slice-0 thread takes the next chain from the queue and posts it for the next pass
        if (ph == NP_FETCH && !fetch_posted) {
            if (slice == 0) {
                const unsigned int nxt = atomicAdd(queue, 1u);
                sh->newm[chain] = (nxt < (unsigned int)a.Nchain) ? (int)nxt : -1;
            }
            fetch_posted = true;
        }
        NT_MARK(pn * 100 + 64);
        NP_T(tE);
        NP_ADD(5, tD, tE);
        // ===== E. operand rows of the next gradient pass ====================================================================
        __syncwarp();
        put_half0<PREC>(acol, x);
        put_half1<PREC>(acol, x, wide);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if (slice == 0) {
            const int alive = __any_sync(HMC_FULL_MASK, ph != NP_IDLE);
            if (lane == 0) sh->galive[pn & 1][grp] = alive;
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        NT_MARK(pn * 100 + 70);
        NP_T(tF);
        NP_ADD(6, tE, tF);
#ifdef HMC_PROFILE_PHASES
        tph[7] += 1;
#endif
        s1_arrive();                                            // S1: rows written, alive flags posted (only the issuing warp waits there)
        have_grad = true;
        ++pn;
    }

#ifdef HMC_PROFILE_PHASES
    if (lane == 0) for (int i = 0; i < 8; ++i) atomicAdd(&g_nt_cycles[i], (unsigned long long)tph[i]);
#endif
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    if (slice == 0) {
        const unsigned long long c0 = warp_sum<unsigned long long>(c_leap), c1 = warp_sum<unsigned long long>(c_doubling);
        const unsigned long long c2 = warp_sum<unsigned long long>(c_instab), c3 = warp_sum<unsigned long long>(c_dmax);
        if (lane == 0) {
            atomicAdd(a.counters + 0, c0); atomicAdd(a.counters + 1, c1);
            atomicAdd(a.counters + 2, c2); atomicAdd(a.counters + 3, c3);
        }
    }
}

#ifdef HMC_PROFILE_PHASES
}  // namespace
extern "C" int hmc_debug_nuts_tc_cycles(unsigned long long* out8, int reset) {
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_nt_cycles, z, sizeof(z)); return 0; }
    cudaMemcpyFromSymbol(out8, g_nt_cycles, sizeof(unsigned long long) * 8);
    return 0;
}
namespace {
#endif

constexpr size_t nuts_tc_smem_bytes(int npart) {
    return (size_t)npart * TC_BPART + sizeof(float) * 2 * TC_KP + sizeof(NutsTcShared) + 64;
}
static_assert(nuts_tc_smem_bytes(3) <= 232448, "shared memory of the tensor-core NUTS kernel exceeds 227 KB");

template <int PREC>
int nuts_tc_launch(const hmc_nuts_args& a, int grid, unsigned int* queue, cudaStream_t stream) {
    const size_t smem = nuts_tc_smem_bytes(TcPrec<PREC>::NPART);
    HMC_CUDA_CHECK(cudaFuncSetAttribute(hmc_nuts_tc_kernel<PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hmc_nuts_tc_kernel<PREC><<<grid, TC_NT, smem, stream>>>(a, queue);
    HMC_CUDA_CHECK(cudaGetLastError());
    return HMC_OK;
}

}  // namespace

bool hmc_nuts_tc_supported(const hmc_nuts_args& a, const char** why) {
    if (a.dtype != HMC_F32) { *why = "float32 only"; return false; }
    if (a.target.Mit || a.target.Pt || a.target.Lct) { *why = "identity momentum metric only"; return false; }
    if (a.target.D != TC_ND) { *why = "D == 100 in this build"; return false; }
    if (a.d_max > NT_MAXCHK) { *why = "d_max <= 12"; return false; }
    if (a.iter_end <= a.iter_begin) { *why = "needs at least one iteration"; return false; }
    return true;
}

// the work-queue head lives in the last row of chain 0's scratch (the cursor row, whose first 16 bytes hold chain 0's tape
// cursors): bytes 64..67
int hmc_nuts_run_tc(const hmc_nuts_args& a, cudaStream_t stream) {
    bool fp16 = false;
    if (const char* e = getenv("HMC_B200_TC_PREC")) fp16 = (e[0] == 'f');
    int dev = 0, sms = 0;
    HMC_CUDA_CHECK(cudaGetDevice(&dev));
    HMC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int grid = (a.Nchain + TC_M - 1) / TC_M;
    if (grid > sms) grid = sms;
    const int R = 2 * (a.d_max + 1);
    unsigned int* queue = reinterpret_cast<unsigned int*>((float*)a.scratch + (size_t)(R + 6) * a.target.D_pad) + 16;
    HMC_CUDA_CHECK(cudaMemsetAsync(queue, 0, sizeof(unsigned int), stream));
    return fp16 ? nuts_tc_launch<PREC_FP16X2>(a, grid, queue, stream) : nuts_tc_launch<PREC_BF16X3>(a, grid, queue, stream);
}
