// Tensor-core random-trajectory HMC kernel (tcgen05 + TMEM), FP32-grade arithmetic through a bf16x3 split, D = 100.
//
// Follows HMC_sampler.gen_sample_random + leap_frog (/root/reference/samplers.py:387-491, 831-839).
//
// One CTA = 128 chains: 512 worker threads + one warpgroup with the MMA-issuing warp and three copying warps.  The
// gradient of all 128 chains,
//       G[128 x N] = Dm[128 x K] * F[N x K]^T          (K = N = 112: D = 100 zero-padded to a multiple of 16)
// runs on the 5th-generation tensor cores.  The shifted positions d = q - mu are split into three bf16 parts
// (d = d1 + d2 + d3 exactly) that live in TENSOR MEMORY (A operand from TMEM, written with tcgen05.st: row = TMEM lane
// = chain, two bf16 per 32-bit column); the force matrix is split once the same way into shared memory (canonical
// no-swizzle K-major layout: 16-byte chunk (kc, row) at (kc*rows + row)*16, LBO = rows*16, SBO = 128).  Six
// tcgen05.mma passes (1,3) (3,1) (2,2) (1,2) (2,1) (1,1) accumulate in fp32 in tensor memory -- the dropped terms
// are O(2^-24).  Keeping A out of shared memory halves the tensor pipe's shared-memory traffic (an SS-mode pass reads
// 7.5 KB per MMA = the whole shared-memory bandwidth, which starves everything that runs beside it).
//
// FOUR threads share a chain: thread (warp w, lane) works on chain 32*(w%4) + lane (the TMEM lanes a warp may
// access) and on the dimension slice w/4 (24, 24, 24, 28 dims).  Position and momentum slices stay in registers; a
// pass is: read the slice of the gradient (tcgen05.ld), leapfrog update, re-split, tcgen05.st.  Per-chain sums
// (d.g, p.p) are combined through shared memory by the slice-0 thread, which does the chain's bookkeeping
// (energies, Metropolis accept on a Philox uniform, new trajectory length) and posts a command that all four slice
// threads apply at the top of the next pass (new start point / restore / take the next momentum: all through
// per-chain rows in shared memory).  Every chain advances one gradient evaluation per pass; iteration boundaries are
// per-chain events; the first point of each trajectory is a gradient-only pass, so an iteration costs
// L + 1 evaluations.
//
// Schedule of pass n:   workers: P1a apply the commands of P2(n-1), P1b wait for MMA(n-1), update, re-split | S1 |
// issuing warp: MMA(n) from the rows as they stand; under it, per 32-chain group, the slice-0 warp runs the bookkeeping
// P2(n) while the other three warps draw momenta (D), then a group barrier.  The workers only ARRIVE at S1 (the
// issuing and copying warps wait there): all synchronisation between workers is per 32-chain group, and what a worker
// needs from the rest of the CTA is the MMA, which it gets through the MMA's mbarrier.  An accepted chain's in-flight gradient is
// the first gradient of its next trajectory; a rejected chain's row is restored in P1a and used one pass later.
// Momenta are drawn ahead (the draw of iteration i+1 does not depend on the accept decision of iteration i:
// samplers.py:431, 441 draw p, L, u at the top of every iteration), requested when a trajectory starts, warp-
// cooperatively (one Philox call per lane, the same draws as every other kernel), one draw per warp and pass, into
// the chain's staging row; P2 hands a momentum over only when drawn[] says it is there.  The copying warps write every
// stored sample (and a unit's final state) as one coalesced 400-byte row taken from the chain's start-point row.
// Finished slots pull the next (chain, sub-block of the iteration block) unit from a global queue.
#include "tc_common.cuh"
#include <cstdlib>
#include <cmath>

#ifdef HMC_TC_DEBUG
// progress markers in mapped host memory (readable while a kernel hangs): g_tc_dbg[warp] = pass * 100 + stage, block 0 only
__device__ volatile int* g_tc_dbg = nullptr;
#define TC_MARK(stage) do { if (g_tc_dbg && blockIdx.x == 0 && (threadIdx.x & 31) == 0) { g_tc_dbg[threadIdx.x >> 5] = dbg_pass * 100 + (stage); __threadfence_system(); } } while (0)
#else
#define TC_MARK(stage)
#endif

#ifdef HMC_PROFILE_PHASES
__device__ unsigned long long g_tc_cycles[16];     // [0..8) bookkeeping warps, [8..16) the others; slot 4 = issuing warp
__device__ unsigned long long g_tc_apply[8];       // inside 'apply commands': [0] command words + new / parked chains, [1] start-point rows, [2] restore / take loads, [3] warp sync
#define TP_T(x) const unsigned int x = (unsigned int)clock()
#define TP_ADD(i, a, b) tph[i] += (b) - (a)
#else
#define TP_T(x)
#define TP_ADD(i, a, b)
#endif

namespace {

constexpr int TC_SROW = TC_ND;                  // floats per momentum staging row (one row per chain)

// commands of the bookkeeping thread to the four slice threads of its chain (applied at the top of the next pass)
enum : int {
    CMD_STORE_Q0 = 1,   // accepted: the proposal becomes the chain's start-point row
    CMD_RESTORE = 4,    // rejected: positions <- start-point row
    CMD_NEW = 8,        // load a chain (CMD_NEW0: from q_start, and record it as stored sample 0; else from state_q)
    CMD_PARK = 16,      // no chain left for this slot: zero rows
    CMD_NEW0 = 64,
    CMD_TAKE = 128      // momentum <- the chain's staging row
};
enum : int { MODE_IDLE = 0, MODE_FIRST = 1, MODE_MID = 2, MODE_LAST = 3 };

struct TcGen {
    uint64_t seed;
    const double* p_tape;
    const int32_t* L_tape;
    const double* u_tape;
    int Niter, L_low, L_high;
    int D;                              // dimensions of the target (<= TC_ND; the rest of the 100-wide rows is zero padding)
};

// Warp-cooperative momentum draw for one chain: lane sl < 25 draws the normals of dims 4*sl..4*sl+3 into `stage`,
// lane 25 the scalars of the iteration.  Same arithmetic as hmc_normal4 / hmc_scalar_draws.  A target with D < 100 uses the
// first D / 4 lanes; the padded dimensions keep momentum zero (their force rows are zero too), so they never move.
template <bool DFULL>
__device__ __forceinline__ void tc_gen(const TcGen& g, long m, uint64_t gid, int iter, int lane, float* stage, float* sumsq,
                                       int* L, float* lnu) {
    constexpr int nslot = TC_ND / 4;
    const int D = DFULL ? TC_ND : g.D;
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
        if (4 * lane < D) {
            *reinterpret_cast<float4*>(stage + 4 * lane) = z;
            s = z.x * z.x + z.y * z.y + z.z * z.z + z.w * z.w;
        }
        const int Ls = g.L_low + (int)__umulhi(r.x, (uint32_t)(g.L_high - g.L_low));
        const float ls = logf(((float)(r.y >> 8) + 0.5f) * 5.9604644775390625e-08f);
        Lv = __shfl_sync(HMC_FULL_MASK, Ls, nslot);
        lv = __shfl_sync(HMC_FULL_MASK, ls, nslot);
    }
    // |p|^2 over the warp in one REDUX instead of five dependent shuffle-add rounds (the draw is a latency chain on the
    // group's critical path): fixed point with 2^-18 resolution -- a lane's four squares stay below 2^13 (|z| < 45), 25
    // lanes below 2^18, so the sum fits 32 bits; the rounding (4e-6 absolute on |p|^2 ~ 100) is far below the float32
    // resolution of the energies it enters.
    *sumsq = (float)__reduce_add_sync(HMC_FULL_MASK, __float2uint_rn(s * 262144.0f)) * (1.0f / 262144.0f);
    *L = Lv;
    *lnu = lv;
    __syncwarp();
}

struct TcShared {                       // small per-chain arrays in shared memory
    float2 red[TC_SPL][TC_M];           // partial (d.g, p.p) per slice
    int mode[TC_M];                     // MODE_* of the gradient in flight
    int cmd[TC_M];                      // CMD_* flags posted by the bookkeeping thread, applied at the top of the next P1
    int cm[TC_M];                       // local chain index of the chain in this slot
    int out_req[4][TC_M];               // by pass number & 3: row copies for the copying warps: (stored-sample index + 1) | OUT_* flags, 0 = none
    int out_m[4][TC_M];                 // local chain index of the copy
    int req[TC_M];                      // pending momentum draw request: iteration (| REQ_INIT0), 0 = none (cleared by the drawer)
    int drawn[TC_M];                    // iteration whose momentum is staged (row and scalars complete)
    float gK[TC_M], gK0[TC_M], glnu[TC_M];   // results of the momentum draw
    int gL[TC_M];
    int galive[4][4];                   // per pass number & 3 and group: some slot still has (or wants) a chain
    unsigned long long cnt[4][TC_M];    // per-slot counters (accepted warm-up / kept iterations, sum L, sum L^2): shared memory, not
                                        // registers -- every register of the workers counts (a build with 17 more spilled words ran 6 % slower)
    int stop;                           // set by the issuing warp when the CTA is done (the workers see it after the MMA wait)
    unsigned int fmax_bits;             // bit pattern of max |F| (set-up only: scale of the fp16 B parts)
};

constexpr int OUT_SAMPLE = 1 << 29, OUT_STATE = 1 << 30;   // copy the chain's start-point row to q_chain[m][idx] / to state_q[m]
constexpr int REQ_INIT0 = 1 << 30;      // request flag: also draw the chain-start momentum (iteration 0, K only)

template <bool UDT, int PREC, bool DFULL>
__global__ void __launch_bounds__(TC_NT, 1) hmc_random_tc_kernel(const hmc_random_args a, unsigned int* __restrict__ queue, int* progress, int nsb, int SB) {
    constexpr int KP = TC_KP, KC = TC_KC;
    // DFULL: D == TC_ND at compile time (the benchmark path keeps its constant addressing); otherwise D <= TC_ND, D % 4 == 0
    // and dimensions D..111 are zero padding (force, mu, dt, rows)
    const int D = DFULL ? TC_ND : a.target.D;
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int NPART = TcPrec<PREC>::NPART;
    unsigned char* Bp = smem;                                   // NPART parts [KC][112] 16-byte chunks
    float* mu_s = reinterpret_cast<float*>(Bp + NPART * TC_BPART);  // [KP]
    float* dt_s = mu_s + KP;                                    // [KP]
    float* stage_all = dt_s + KP;                               // [128][TC_SROW] momentum staging, one row per chain
    float* q0_s = stage_all + TC_M * TC_SROW;                   // [128][TC_SROW] shifted start position of the trajectory
    TcShared* sh = reinterpret_cast<TcShared*>(q0_s + TC_M * TC_SROW);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sh + 1);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef HMC_TC_DEBUG
    int dbg_pass = 0;
#endif
    TC_MARK(1);
    const int grp = warp & 3;                    // 32-chain group == TMEM lane quarter this warp may access
    const int slice = warp >> 2;                 // dimension slice (4 = the issuing warpgroup)
    const int chain = grp * 32 + lane;           // chain slot of this thread
    const int j0 = 24 * (slice & 3);             // first dimension of the slice
    const bool wide = slice == TC_SPL - 1;       // the last slice has 28 dims, the others 24

    // ---- one-time set-up ---------------------------------------------------------------------------------------------
    // fp16 split: the force matrix is scaled by a power of two so that its largest entry sits just below the top of the fp16
    // range (no part of an entry of interest is an fp16 subnormal, whatever the magnitude of F).  The accumulator then holds
    // 2^s g; the factor 2^-s is folded into the kick weight and the V sum, exactly.
    float binv = 1.f, bscale = 1.f;
    if constexpr (TcPrec<PREC>::F16) {
        const float* Ft = (const float*)a.target.Ft;
        const int Dpad = a.target.D_pad;
        if (tid == 0) sh->fmax_bits = 0u;
        __syncthreads();
        unsigned int mx = 0u;
        for (int t = tid; t < D * D; t += TC_NT) mx = max(mx, __float_as_uint(fabsf(Ft[(size_t)(t / D) * Dpad + (t % D)])));
        mx = __reduce_max_sync(HMC_FULL_MASK, mx);
        if (lane == 0) atomicMax(&sh->fmax_bits, mx);
        __syncthreads();
        const int e = (int)(sh->fmax_bits >> 23) - 127;         // max |F| in [2^e, 2^(e+1))
        int sft = (sh->fmax_bits == 0u || e < -100) ? 0 : 14 - e; // scaled maximum in [2^14, 2^15)
        sft = sft > 100 ? 100 : (sft < -100 ? -100 : sft);
        bscale = __uint_as_float((unsigned int)(127 + sft) << 23);
        binv = __uint_as_float((unsigned int)(127 - sft) << 23);
    }
    {
        const float* Ft = (const float*)a.target.Ft;            // Ft[k][n] = F[n][k]; B row n holds F[n][.] (K-major)
        const int Dpad = a.target.D_pad;
        for (int t = tid; t < KC * KP; t += TC_NT) {            // one 16-byte chunk (8 k values) of row n per item
            const int kc = t / KP, n = t % KP;
            uint32_t w1[4], w2[4], w3[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k0 = kc * 8 + 2 * e, k1 = k0 + 1;
                const float x0 = (n < D && k0 < D) ? Ft[(size_t)k0 * Dpad + n] * bscale : 0.f;
                const float x1 = (n < D && k1 < D) ? Ft[(size_t)k1 * Dpad + n] * bscale : 0.f;
                split_pair<PREC>(x0, x1, w1[e], w2[e], w3[e]);
                w2[e] ^= 0x80008000u;                           // split_pair returns the second part negated; B holds it as is
            }
            reinterpret_cast<uint4*>(Bp)[t] = make_uint4(w1[0], w1[1], w1[2], w1[3]);
            reinterpret_cast<uint4*>(Bp + TC_BPART)[t] = make_uint4(w2[0], w2[1], w2[2], w2[3]);
            if constexpr (NPART == 3) reinterpret_cast<uint4*>(Bp + 2 * TC_BPART)[t] = make_uint4(w3[0], w3[1], w3[2], w3[3]);
        }
        for (int t = tid; t < KP; t += TC_NT) {
            mu_s[t] = (t < D) ? ((const float*)a.target.mu)[t] : 0.f;
            dt_s[t] = (t < D) ? ((const float*)a.target.dt)[t] : 0.f;
        }
        for (int t = tid; t < 2 * TC_M * TC_SROW; t += TC_NT) stage_all[t] = 0.f;      // staging and start-point rows (padding stays zero)
        for (int t = tid; t < TC_M; t += TC_NT) { sh->mode[t] = MODE_IDLE; sh->cmd[t] = 0; sh->req[t] = 0; sh->cm[t] = 0; sh->drawn[t] = -1; sh->out_req[0][t] = 0; sh->out_req[1][t] = 0; sh->out_req[2][t] = 0; sh->out_req[3][t] = 0;
                                                 sh->cnt[0][t] = 0ull; sh->cnt[1][t] = 0ull; sh->cnt[2][t] = 0ull; sh->cnt[3][t] = 0ull; }
        if (tid < 16) sh->galive[tid >> 2][tid & 3] = 1;
        if (tid == 0) sh->stop = 0;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    TC_MARK(2);
    asm volatile("fence.proxy.async.shared::cta;");             // B parts (generic-proxy writes) -> tensor core
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;
    TC_MARK(3);
    const uint32_t tmem_row = tmem + ((uint32_t)(grp * 32) << 16) + (uint32_t)j0;                   // my accumulator slice
    const uint32_t acol = tmem + ((uint32_t)(grp * 32) << 16) + TC_ACOL + 12u * (uint32_t)(slice & 3);  // my A-part columns
    // instruction descriptor: D = f32 (bits 4-5), A / B format (bits 7-9 / 10-12: 0 = f16, 1 = bf16), N >> 3 (17-22), M >> 4 (24-28)
    constexpr uint32_t fmt = TcPrec<PREC>::F16 ? 0u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(KP >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

    const long Lc = 1 + (a.Niter - a.warm_up_num) / a.thin_rate;      // samplers.py:31
    const long Lrow = a.store_ring > 0 ? a.store_ring : Lc;         // rows allocated per chain (ring of the last store_ring stored samples)
    float* q_chain = (float*)a.q_chain;
    float* q0g = (float*)a.state_q;
    const float vconst = (float)a.target.v_const;
    TcGen ga;
    ga.seed = a.seed; ga.p_tape = a.p_tape; ga.L_tape = a.L_tape; ga.u_tape = a.u_tape;
    ga.Niter = a.Niter; ga.L_low = a.L_low; ga.L_high = a.L_high; ga.D = D;

    if (slice >= TC_SPL) {
        // ===== the issuing warpgroup: hands its registers to the workers (per scheduler: 4 x 112 + 24 registers x 32 lanes
        //       = 16 K).  Its first warp launches, after every S1 barrier, the gradient pass of the operand rows as they
        //       stand; the workers' bookkeeping, commands and momentum draws run under it. =====================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        TC_MARK(4);
        if (warp == TC_THREADS / 32) {                        // the other three warps of the group only gave their registers
            int pn = 0;                                         // pass number
#ifdef HMC_PROFILE_PHASES
            long long tph4 = 0;
#endif
            while (true) {
                TC_MARK(10);
                asm volatile("tcgen05.fence::before_thread_sync;");
                bar_all();
                TC_MARK(11);
                const volatile int* ga_ = sh->galive[(pn - 1) & 3];     // posted by P2 of the previous pass
                if ((ga_[0] | ga_[1] | ga_[2] | ga_[3]) == 0) {
                    // done: the workers are (or will be) waiting for the pass that is not coming -- tell them and release them
                    if (lane == 0) {
                        *reinterpret_cast<volatile int*>(&sh->stop) = 1;
                        __threadfence_block();
                        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(mbar)) : "memory");
                    }
                    break;
                }
                ++pn;
                if (lane == 0) {
    #ifdef HMC_PROFILE_PHASES
                    const long long ti0 = clock64();
    #endif
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint64_t dsc = make_desc(smem_u32(Bp), KP * 16, 128);
                    const uint32_t dlo = (uint32_t)dsc, dhi = (uint32_t)(dsc >> 32);
                    tc_mma_all<PREC, 0>(tmem, dlo, dhi, idesc);
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
    #ifdef HMC_PROFILE_PHASES
                    tph4 += clock64() - ti0;
    #endif
                }
                TC_MARK(12);
#ifdef HMC_TC_DEBUG
                ++dbg_pass;
#endif
                __syncwarp();
            }
#ifdef HMC_PROFILE_PHASES
            if (lane == 0) atomicAdd(&g_tc_cycles[4], (unsigned long long)tph4);
#endif
        } else {
            // ===== the copying warps: a finished trajectory's stored sample (samplers.py:462-472) -- and, at the end of a
            //       unit, the chain state -- is the chain's start-point row in shared memory (accepted proposal or restored
            //       point) + mu: one coalesced 400-byte row per warp instruction, instead of 16-byte pieces from the four
            //       slice threads.  Rows posted by P2(n-1), complete after the workers' P1(n), are copied during pass n. ==========
            const int cw = warp - TC_THREADS / 32 - 1;          // 0..2
            int pn = 0;                                         // pass number
            while (true) {
                bar_all();
                const int rp = (pn - 1) & 3;                    // ring slot of the previous pass's P2
                const volatile int* ga_ = sh->galive[rp];
                if ((ga_[0] | ga_[1] | ga_[2] | ga_[3]) == 0) break;
                int k = 0;
                bool state_row = false;
                for (int g = 0; g < 4; ++g) {
                    unsigned todo = __ballot_sync(HMC_FULL_MASK, sh->out_req[rp][g * 32 + lane] != 0);
                    while (todo) {
                        const int cs = g * 32 + __ffs(todo) - 1;
                        todo &= todo - 1;
                        if (k == cw && 4 * lane < D) {
                            const int r = sh->out_req[rp][cs];
                            const size_t mc = (size_t)sh->out_m[rp][cs];
                            const float4 d4 = *reinterpret_cast<const float4*>(q0_s + cs * TC_SROW + 4 * lane);
                            const float4 mu4 = *reinterpret_cast<const float4*>(mu_s + 4 * lane);
                            const float4 v = make_float4(d4.x + mu4.x, d4.y + mu4.y, d4.z + mu4.z, d4.w + mu4.w);
                            if (r & OUT_SAMPLE)
                                *reinterpret_cast<float4*>(q_chain + (mc * Lrow + (size_t)((r & 0xfffffff) - 1)) * D + 4 * lane) = v;
                            if (r & OUT_STATE) { *reinterpret_cast<float4*>(q0g + mc * D + 4 * lane) = v; state_row = true; }
                        }
                        k = (k == 2) ? 0 : k + 1;
                    }
                }
                if (state_row) __threadfence();                 // the unit's progress flag is raised after the next S1
                ++pn;
            }
        }
        __syncthreads();
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    TC_MARK(5);

    // ---- per-thread state -----------------------------------------------------------------------------------------------
    float p[28], x[28];              // momentum and shifted-position (q - mu) slices, registers
#pragma unroll
    for (int j = 0; j < 28; ++j) { p[j] = 0.f; x[j] = 0.f; }
    // bookkeeping state of chain `chain`, used by the slice-0 thread only
    int m = -1;                      // local chain index, -1 = no chain
    int it = 0, l = 0, L = 1;        // iteration, point index of the next gradient, trajectory length
    bool init = false;               // chain start: E_chain[.,0] still to be recorded
    bool want = (slice == 0);        // needs a (new) chain
    bool delayed = false;            // operand row (re)written by the pending command: first gradient one pass later
    bool need_take = false;          // new chain: its first momentum is drawn in the pass after the load
    int it_end = 0, sub = 0;         // last iteration and sub-block index of the unit being run
    int wait_unit = -1;              // dequeued unit waiting for its predecessor (-1 none, -2 queue empty)
    const bool thin1 = a.thin_rate == 1;
    const bool ring = a.store_ring > 0;
    int publish = 0;                 // passes until a finished unit's state (copied by a copying warp) is announced
    int pn = 0;                      // pass number
    unsigned pend = 0;               // chains of my group with a pending momentum request (snapshot at the group barrier)
    float E_init = 0.f, E_prev = 0.f, lnu = 0.f;
    uint32_t phase = 0;
    bool have_grad = false;          // a gradient pass has been issued and its accumulator is to be consumed
#ifdef HMC_PROFILE_PHASES
    unsigned int tph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned int tpa[4] = {0, 0, 0, 0};
#endif

    {                                // defined operand rows before the first pass: all parts zero (K padding included)
        put_half0<PREC>(acol, x);
        put_half1<PREC>(acol, x, true);    // the wide form also clears columns 12, 13 of the slice; harmless for the others
        if (wide) {                  // K padding: dims 100..111 = columns 50..55 of each part
            const uint32_t z[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int pt = 0; pt < NPART; ++pt) {
                tmem_st4(acol + 14 + pt * TC_APITCH, z);
                tmem_st2(acol + 18 + pt * TC_APITCH, z);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }

    // Pass n of a worker:  P1 (apply the commands of P2(n-1), consume G(n-1))  |  S1  |  P2(n) on the slice-0 warp, momentum
    // draws D(n) on the other three warps of the group  |  group barrier.  The issuing warp launches MMA(n) right after S1.
    TC_MARK(6);
    while (true) {
        TP_T(t0);
        TC_MARK(20);
        // ===== P1a. commands posted by P2 of the previous pass (visible through the group barrier).  The register
        //       slices are only touched by predicated in-place loads (no per-path copies of the arrays): a new or parked
        //       chain first writes its shared-memory rows, then takes them like a restored chain. =============================
        const int md = sh->mode[chain];
        {
            const int cmd = sh->cmd[chain];
            const int nch4 = wide ? 7 : 6;
            float* q0r = q0_s + chain * TC_SROW + j0;
            float* str = stage_all + chain * TC_SROW + j0;
            if (cmd & (CMD_NEW | CMD_PARK)) {
                if (cmd & CMD_NEW) {                                  // samplers.py:411-413
                    const size_t mc = (size_t)sh->cm[chain];
                    const bool fresh = (cmd & CMD_NEW0) != 0;           // chain start: q_start; otherwise the state another
                    const float* src = (fresh ? (const float*)a.q_start : q0g) + mc * D + j0;       // unit / launch left in state_q
#pragma unroll
                    for (int c = 0; c < 7; ++c) {
                        if (c < nch4 && j0 + 4 * c < D) {
                            const float4 v = __ldcg(reinterpret_cast<const float4*>(src + 4 * c));
                            const float4 mu4 = *reinterpret_cast<const float4*>(mu_s + j0 + 4 * c);
                            if (fresh) *reinterpret_cast<float4*>(q_chain + mc * Lrow * D + j0 + 4 * c) = v;
                            const float4 d4 = make_float4(v.x - mu4.x, v.y - mu4.y, v.z - mu4.z, v.w - mu4.w);
                            *reinterpret_cast<float4*>(q0r + 4 * c) = d4;
                            if constexpr (TcPrec<PREC>::F16) {       // start point outside the range of the fp16 split: tell the host
                                if (fresh && !(fmaxf(fmaxf(fabsf(d4.x), fabsf(d4.y)), fmaxf(fabsf(d4.z), fabsf(d4.w))) < 16384.f))
                                    atomicOr(progress + a.Nchain, 1);
                            }
                        }
                    }
                } else {                                              // no chain left for this slot: zero rows
#pragma unroll
                    for (int c = 0; c < 7; ++c) {
                        if (c < nch4) {
                            *reinterpret_cast<float4*>(q0r + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
                            *reinterpret_cast<float4*>(str + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                }
            }
#ifdef HMC_PROFILE_PHASES
            const unsigned int ta1 = (unsigned int)clock(); tpa[0] += ta1 - t0;
#endif
            if (cmd & CMD_STORE_Q0) {                                 // accepted: the proposal is the new start point
#pragma unroll
                for (int c = 0; c < 7; ++c)
                    if (c < nch4) *reinterpret_cast<float4*>(q0r + 4 * c) = make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
            }
#ifdef HMC_PROFILE_PHASES
            const unsigned int ta2 = (unsigned int)clock(); tpa[1] += ta2 - ta1;
#endif
            {
                const int rs = (cmd & (CMD_RESTORE | CMD_NEW | CMD_PARK)) != 0, tk = (cmd & (CMD_TAKE | CMD_PARK)) != 0;
                const uint32_t qa = smem_u32(q0r), sa = smem_u32(str);
#pragma unroll
                for (int c = 0; c < 7; ++c) {
                    if (c < 6 || wide) {
                        lds4_if(rs, qa + 16 * c, x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
                        lds4_if(tk, sa + 16 * c, p[4 * c], p[4 * c + 1], p[4 * c + 2], p[4 * c + 3]);
                    }
                }
            }
#ifdef HMC_PROFILE_PHASES
            const unsigned int ta3 = (unsigned int)clock(); tpa[2] += ta3 - ta2;
#endif
        }
#ifdef HMC_PROFILE_PHASES
        const unsigned int ta4 = (unsigned int)clock();
#endif
        __syncwarp();
        TP_T(t0b);
        TP_ADD(7, t0, t0b);
#ifdef HMC_PROFILE_PHASES
        tpa[3] += t0b - ta4;
#endif
        // ===== P1b. consume the gradient: thread-local leapfrog update of the slice (samplers.py:835-837), re-split ======
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
            if (*reinterpret_cast<volatile int*>(&sh->stop)) break;    // released by the issuing warp: nothing left to do
            TC_MARK(21);
            TP_T(t1);
            TP_ADD(0, t0b, t1);
            // point index of the gradient: first = half kick + drift, last = half kick only, interior = second half
            // kick of step l + first half kick of step l+1, drift
            const float kwt = (md == MODE_IDLE) ? 0.f : (md == MODE_MID ? -binv : -0.5f * binv);    // (binv: the accumulator holds g / binv)
            const float dwt = (md == MODE_FIRST || md == MODE_MID) ? 1.f : 0.f;
            const float dt0 = dt_s[0], kdt = kwt * dt0, ddt = dwt * dt0;      // uniform step size (UDT)
            const bool tr = slice == 0 && a.phi_q && m >= 0 && a.chain_id0 + m == 0 && it <= a.N_save_chain0;
            if (tr && md == MODE_FIRST) {
                double* phi = a.phi_q + (size_t)(it - 1) * a.L_high * 2;      // row 0: the start point (samplers.py:445)
                phi[0] = (double)(x[0] + mu_s[0]); phi[1] = (double)(x[1] + mu_s[1]);
            }
            float hv = 0.f, hk = 0.f;
            uint32_t gv[16];
            tmem_ld16(tmem_row, gv);
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const float gj = __uint_as_float(gv[jj]);
                const float dtj = UDT ? dt0 : dt_s[j0 + jj];
                hv = fmaf(x[jj], gj, hv);
                const float pn = fmaf(gj, UDT ? kdt : kwt * dtj, p[jj]);
                hk = fmaf(pn, pn, hk);
                p[jj] = pn;
                x[jj] = fmaf(pn, UDT ? ddt : dwt * dtj, x[jj]);
            }
            put_half0<PREC>(acol, x);
            tmem_ld16(tmem_row + 16u, gv);
#pragma unroll
            for (int jj = 16; jj < 28; ++jj) {
                if (jj < 24 || wide) {
                    const float gj = __uint_as_float(gv[jj - 16]);
                    const float dtj = UDT ? dt0 : dt_s[j0 + jj];
                    hv = fmaf(x[jj], gj, hv);
                    const float pn = fmaf(gj, UDT ? kdt : kwt * dtj, p[jj]);
                    hk = fmaf(pn, pn, hk);
                    p[jj] = pn;
                    x[jj] = fmaf(pn, UDT ? ddt : dwt * dtj, x[jj]);
                }
            }
            put_half1<PREC>(acol, x, wide);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            sh->red[slice][chain] = make_float2(hv, hk);
            if (tr && (md == MODE_FIRST || md == MODE_MID)) {
                // chain-0 trajectory capture (samplers.py:442-452): the point reached by this drift is row l+1
                double* phi = a.phi_q + (size_t)(it - 1) * a.L_high * 2;
                const int row = (md == MODE_FIRST) ? 1 : l + 1;
                phi[2 * row] = (double)(x[0] + mu_s[0]); phi[2 * row + 1] = (double)(x[1] + mu_s[1]);
            }
            TP_T(t2);
            TP_ADD(1, t1, t2);
        } else {                                                // first pass: rows of the chains loaded by nobody yet
            put_half0<PREC>(acol, x);
            put_half1<PREC>(acol, x, wide);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        TC_MARK(22);
        TP_T(t2b);
        asm volatile("tcgen05.fence::before_thread_sync;");     // TMEM reads / writes ordered before the next MMA
        bar_all_arrive();                                       // S1: my rows are written (the issuing warp waits for everybody)
        bar_group(grp);                                         // the group's partial sums are in shared memory
        have_grad = true;
        TC_MARK(23);
        TP_T(t3);
        TP_ADD(2, t2b, t3);

        if (slice == 0) {
            // ===== P2. per-chain bookkeeping by the slice-0 thread (under the running pass) ===============================
            int cmd = 0, req = 0, oreq = 0;
            const int mdp = sh->mode[chain];                    // mode of the gradient P1 has just consumed
            if (m >= 0 && mdp != MODE_IDLE) {
                float sv = 0.f, sk = 0.f;
#pragma unroll
                for (int s2 = 0; s2 < TC_SPL; ++s2) { const float2 r = sh->red[s2][chain]; sv += r.x; sk += r.y; }
                const bool tr = a.phi_q && (a.chain_id0 + m) == 0;
                const float V = fmaf(0.5f * binv, sv, vconst);                  // V(q) = 0.5 d.P d + const (utils.py:213-218)
                if (mdp == MODE_FIRST) {
                    // the draws of this iteration (requested at least two passes ago): samplers.py:431, 441
                    const float Knew = 0.5f * sh->gK[chain];
                    L = sh->gL[chain]; lnu = sh->glnu[chain];
                    sh->cnt[2][chain] += (unsigned long long)L; sh->cnt[3][chain] += (unsigned long long)(L * L);   // 64-bit: sum L^2 of a long launch exceeds 2^32
                    if (tr && it <= a.N_save_chain0) a.phi_len[it - 1] = L + 1;                     // samplers.py:444
                    if (init) {                                                 // samplers.py:416-420
                        const float E0 = V + 0.5f * sh->gK0[chain];
                        a.E_chain[(size_t)m * Lrow] = (double)E0;
                        a.dE_chain[(size_t)m * Lrow] = 0.0;
                        E_prev = E0;
                        init = false;
                    }
                    E_init = V + Knew;                                          // samplers.py:434-438
                    if (it >= a.warm_up_num) {
                        // (stored index: below L_chain by construction; only a ring of the last store_ring samples wraps.  The 64-bit
                        //  `% Lrow` here and at the trajectory end was a software division in the bookkeeping's latency chain.)
                        int idx = thin1 ? it - a.warm_up_num : (it - a.warm_up_num) / a.thin_rate;
                        if (ring) idx = (int)((unsigned int)idx % (unsigned int)Lrow);
                        a.E_chain[(size_t)m * Lrow + idx] = (double)E_init;
                        a.dE_chain[(size_t)m * Lrow + idx] = (double)(E_init - E_prev);
                    }
                    l = 1;
                    sh->mode[chain] = (l == L) ? MODE_LAST : MODE_MID;
                    if (it < it_end) req = it + 1;                              // next momentum: drawn during the next pass
                } else if (mdp == MODE_LAST) {
                    // Metropolis accept (samplers.py:455-472)
                    const float dE = (V + 0.5f * sk) - E_init;
                    E_prev = E_init;                                            // samplers.py:460
                    const bool accepted = (dE < 0.f) || (lnu < -dE);            // samplers.py:462
                    const bool keep = it >= a.warm_up_num;
                    if (accepted) { sh->cnt[keep ? 1 : 0][chain] += 1ull; cmd |= CMD_STORE_Q0; }
                    else cmd |= CMD_RESTORE;
                    if (keep) {
                        int idx = thin1 ? it - a.warm_up_num : (it - a.warm_up_num) / a.thin_rate;
                        if (ring) idx = (int)((unsigned int)idx % (unsigned int)Lrow);
                        oreq = (idx + 1) | OUT_SAMPLE;
                    }
                    if (tr && it <= a.N_save_chain0) a.decision_chain[it - 1] = accepted ? 1 : 0;
                    if (it >= it_end) {                                         // unit finished: its position goes to state_q
                        a.state_eprev[m] = (double)E_prev;
                        publish = 2;
                        oreq |= OUT_STATE;
                        want = true;
                        sh->mode[chain] = MODE_IDLE;
                    } else {
                        it += 1;
                        // the pass in flight was issued from the proposal: for an accepted chain it is the first gradient
                        // of the new trajectory, a rejected chain's row is restored in P1 and used one pass later; a chain
                        // whose next momentum is not staged yet (draws are rationed to one per warp and pass) waits for it
                        const bool ready = sh->drawn[chain] == it;
                        if (ready) cmd |= CMD_TAKE;
                        need_take = !ready;
                        delayed = !accepted || !ready;
                        sh->mode[chain] = delayed ? MODE_IDLE : MODE_FIRST;
                    }
                } else {
                    l += 1;
                    if (l == L) sh->mode[chain] = MODE_LAST;
                }
            } else if (delayed) {                               // restored / loaded row: the pass in flight is its first
                if (!need_take || sh->drawn[chain] == it) {
                    delayed = false;
                    sh->mode[chain] = MODE_FIRST;
                    if (need_take) { cmd |= CMD_TAKE; need_take = false; }
                }
            } else if (want) {
                // The work queue hands out (chain, sub-block of the iteration block) units: unit u = chain u % Nchain,
                // iterations iter_begin + (u / Nchain) * SB + 1 ... .  Splitting the block evens out the last wave of a
                // launch; a chain's units pass its state through state_q / state_eprev and a progress counter.
                if (publish > 0 && --publish == 0 && sub + 1 < nsb) {     // (a copying warp stored state_q during the previous
                    __threadfence();                                      //  pass; the last unit of a chain has no successor)
                    reinterpret_cast<volatile int*>(progress)[m] = sub + 1;
                }
                if (publish == 0) {                             // (while a finished unit's rows are being copied the slot waits)
                if (wait_unit == -1) {
                    const unsigned int nxt = atomicAdd(queue, 1u);
                    wait_unit = (nxt < (unsigned int)a.Nchain * (unsigned int)nsb) ? (int)nxt : -2;
                }
                if (wait_unit == -2) {
                    want = false;
                    m = -1;
                    cmd = CMD_PARK;
                } else {
                    const int c = wait_unit % a.Nchain, sb = wait_unit / a.Nchain;
                    if (sb == 0 || reinterpret_cast<volatile int*>(progress)[c] >= sb) {     // predecessor unit done?
                        if (sb > 0) __threadfence();
                        want = false;
                        wait_unit = -1;
                        m = c;
                        sub = sb;
                        it = a.iter_begin + sb * SB + 1;
                        it_end = min(a.iter_end, a.iter_begin + (sb + 1) * SB);
                        init = (a.iter_begin == 0 && sb == 0);
                        if (!init) E_prev = (float)__ldcg(a.state_eprev + m);
                        if (init && a.decision_chain && a.chain_id0 + m == 0) a.decision_chain[a.N_save_chain0] = 0;
                        cmd = CMD_NEW | (init ? CMD_NEW0 : 0);
                        req = it | (init ? REQ_INIT0 : 0);
                        sh->cm[chain] = m;
                        sh->drawn[chain] = -1;                  // (the slot's previous chain may have staged this iteration number)
                        delayed = true;                         // row loaded in the next P1
                        need_take = true;
                    }
                }
                }
                sh->mode[chain] = MODE_IDLE;
            }
            sh->cmd[chain] = cmd;
            if (req) sh->req[chain] = req;
            sh->out_req[pn & 3][chain] = oreq;
            if (oreq) sh->out_m[pn & 3][chain] = m;
            const int alive = __any_sync(HMC_FULL_MASK, m >= 0 || want);
            if (lane == 0) sh->galive[pn & 3][grp] = alive;
            TP_T(t4);
            TP_ADD(3, t3, t4);
        } else {
            // ===== D. momentum draws (samplers.py:415, 431, 441) into the chain's own staging row, warp-cooperatively.  The
            //      requests pending at the last group barrier (same snapshot in all warps of the group) are served lowest
            //      chain first, ONE per warp and pass: a warp with two draws would hold the whole CTA at S1, and a draw has
            //      several passes of slack (it is requested when its trajectory starts).  P2 hands a momentum to the workers
            //      only once drawn[] says it is staged. ==========================================================================
            unsigned todo = pend;
            for (int k = 0; k < slice - 1 && todo; ++k) todo &= todo - 1;
            if (todo) {
                const int cs = grp * 32 + (__ffs(todo) - 1);
                const long m_s = sh->cm[cs];
                const int rq = sh->req[cs];
                const uint64_t gid = (uint64_t)(a.chain_id0 + m_s);
                float* st = stage_all + cs * TC_SROW;
                float ks = 0.f, ln = 0.f; int Lx = 1;
                // a chain start also needs the momentum of iteration 0 (samplers.py:415, K only): one call site, two turns
                for (int turn = (rq & REQ_INIT0) ? 0 : 1; turn < 2; ++turn) {
                    tc_gen<DFULL>(ga, m_s, gid, turn ? (rq & ~REQ_INIT0) : 0, lane, st, &ks, &Lx, &ln);
                    if (turn == 0) { if (lane == 0) sh->gK0[cs] = ks; __syncwarp(); }
                }
                if (lane == 0) {
                    sh->gK[cs] = ks; sh->gL[cs] = Lx; sh->glnu[cs] = ln;
                    sh->drawn[cs] = rq & ~REQ_INIT0;
                    sh->req[cs] = 0;
                }
            }
            TP_T(t4);
            TP_ADD(3, t3, t4);
        }
        TC_MARK(24);
        TP_T(t5);
        bar_group(grp);
        pend = __ballot_sync(HMC_FULL_MASK, sh->req[chain] != 0);   // same snapshot in the four warps of the group
        TP_T(t6);
        TP_ADD(6, t5, t6);
        ++pn;
        TC_MARK(28);
#ifdef HMC_TC_DEBUG
        ++dbg_pass;
#endif
#ifdef HMC_PROFILE_PHASES
        tph[5] += 1;
#endif
    }
#ifdef HMC_PROFILE_PHASES
    if (lane == 0) for (int i = 0; i < 8; ++i) if (i != 4) atomicAdd(&g_tc_cycles[(slice == 0 ? 0 : 8) + i], (unsigned long long)tph[i]);
    if (lane == 0) for (int i = 0; i < 4; ++i) atomicAdd(&g_tc_apply[i], (unsigned long long)tpa[i]);
#endif

    TC_MARK(30);
    // every issued pass has been consumed (the exit test follows S1, before the issuing warp launches the next one)
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    if (a.counters) {
        const bool mine = slice == 0;                           // (the bookkeeping threads' own slots; nobody writes them any more)
        const unsigned long long c0 = warp_sum<unsigned long long>(mine ? sh->cnt[0][chain] : 0ull), c1 = warp_sum<unsigned long long>(mine ? sh->cnt[1][chain] : 0ull);
        const unsigned long long c2 = warp_sum<unsigned long long>(mine ? sh->cnt[2][chain] : 0ull), c3 = warp_sum<unsigned long long>(mine ? sh->cnt[3][chain] : 0ull);
        if (lane == 0) {
            atomicAdd(a.counters + 0, c0); atomicAdd(a.counters + 1, c1);
            atomicAdd(a.counters + 2, c2); atomicAdd(a.counters + 3, c3);
        }
    }
}

}  // namespace

#ifdef HMC_TC_DEBUG
extern "C" int hmc_debug_tc_progress(int* mapped) { volatile int* p = mapped; return (int)cudaMemcpyToSymbol(g_tc_dbg, &p, sizeof(p)); }
#endif

#ifdef HMC_PROFILE_PHASES
extern "C" int hmc_debug_tc_cycles(unsigned long long* out16, int reset) {
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_tc_cycles, z, sizeof(z)); cudaMemcpyToSymbol(g_tc_apply, z, sizeof(unsigned long long) * 8); return 0; }
    cudaMemcpyFromSymbol(out16, g_tc_cycles, sizeof(unsigned long long) * 16);
    return 0;
}
extern "C" int hmc_debug_tc_apply(unsigned long long* out8) { cudaMemcpyFromSymbol(out8, g_tc_apply, sizeof(unsigned long long) * 8); return 0; }
#endif

constexpr size_t tc_smem_bytes(int npart) {
    return (size_t)npart * TC_BPART + sizeof(float) * (2 * TC_KP + 2 * TC_M * TC_SROW) + sizeof(TcShared) + 64;
}

static_assert(tc_smem_bytes(3) <= 232448, "shared memory of the tensor-core kernel exceeds 227 KB");

bool hmc_random_tc_supported(const hmc_random_args& a, const char** why) {
    if (a.dtype != HMC_F32) { *why = "float32 only"; return false; }
    if (a.target.Mit || a.target.Pt || a.target.Lct) { *why = "identity momentum metric only"; return false; }
    if (a.target.D > TC_ND || a.target.D < 4 || (a.target.D % 4) != 0) { *why = "D <= 100 and a multiple of 4 (smaller targets run zero padded in the 100-wide tile)"; return false; }
    if (a.iter_end <= a.iter_begin) { *why = "needs at least one iteration"; return false; }
    if (!a.state_g) { *why = "state_g scratch required"; return false; }
    return true;
}

template <bool UDT, int PREC, bool DFULL>
static int tc_launch_d(const hmc_random_args& a, int grid, unsigned int* queue, int* progress, int nsb, int SB, cudaStream_t stream) {
    const size_t smem = tc_smem_bytes(TcPrec<PREC>::NPART);
    HMC_CUDA_CHECK(cudaFuncSetAttribute(hmc_random_tc_kernel<UDT, PREC, DFULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hmc_random_tc_kernel<UDT, PREC, DFULL><<<grid, TC_NT, smem, stream>>>(a, queue, progress, nsb, SB);
    HMC_CUDA_CHECK(cudaGetLastError());
    return HMC_OK;
}

template <bool UDT, int PREC>
static int tc_launch(const hmc_random_args& a, int grid, unsigned int* queue, int* progress, int nsb, int SB, cudaStream_t stream) {
    return a.target.D == TC_ND ? tc_launch_d<UDT, PREC, true>(a, grid, queue, progress, nsb, SB, stream)
                               : tc_launch_d<UDT, PREC, false>(a, grid, queue, progress, nsb, SB, stream);
}

int hmc_random_run_tc(const hmc_random_args& a, cudaStream_t stream) {
    const bool udt = (a.flags & HMC_FLAG_UNIFORM_DT) != 0;      // the step size is the same for all dimensions
    // fp16x2 split (three part products instead of six) when the caller vouches for the range of q - mu (flags bit 1);
    // HMC_B200_TC_PREC=bf16x3|fp16x2 overrides (tests, measurements)
    bool fp16 = (a.flags & HMC_FLAG_TC_FP16X2) != 0;
    if (const char* e = getenv("HMC_B200_TC_PREC")) fp16 = (e[0] == 'f');
    int dev = 0, sms = 0;
    HMC_CUDA_CHECK(cudaGetDevice(&dev));
    HMC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int grid = (a.Nchain + TC_M - 1) / TC_M;
    if (grid > sms) grid = sms;                 // persistent: one CTA per SM, chain slots pull chains from the queue
    // Sub-blocks: the work queue hands out (chain, sub-block of the launch's iteration block) units, which makes the last wave
    // of a launch finer: 65,536 chains on 18,944 slots are 3.46 waves, and at the end of a launch the slots run dry one by one
    // over the length of a unit -- with whole chains as units that tail costs ~half a chain's run time out of 3.5.  A unit costs
    // ~8 passes of set-up and hand-off (through state_q / state_eprev and a progress word), so units of about 100 iterations
    // (~1,300 passes) are used when a launch is long enough; short launches (< 200 iterations) keep whole chains (measured at 50
    // iterations: no gain).  HMC_B200_TC_SUBBLOCKS=n forces n units per chain (tests/test_random_gpu.py exercises the hand-off).
    const int niter = a.iter_end - a.iter_begin;
    int nsb = niter >= 200 ? (niter + 50) / 100 : 1;
    if (const char* e = getenv("HMC_B200_TC_SUBBLOCKS")) { const int n = atoi(e); if (n >= 1 && n <= niter) nsb = n; }
    const int SB = (niter + nsb - 1) / nsb;
    nsb = (niter + SB - 1) / SB;
    unsigned int* queue = (unsigned int*)a.state_g;             // [0] queue head, [16 .. 16 + Nchain) per-chain progress
    int* progress = (int*)a.state_g + 16;
    // (one more word behind the progress counters: "a start point left the fp16 range", kept across the launches of a run)
    HMC_CUDA_CHECK(cudaMemsetAsync(queue, 0, sizeof(int) * (16 + (size_t)a.Nchain + (a.iter_begin == 0 ? 1 : 0)), stream));
    if (fp16) return udt ? tc_launch<true, PREC_FP16X2>(a, grid, queue, progress, nsb, SB, stream)
                         : tc_launch<false, PREC_FP16X2>(a, grid, queue, progress, nsb, SB, stream);
    return udt ? tc_launch<true, PREC_BF16X3>(a, grid, queue, progress, nsb, SB, stream)
               : tc_launch<false, PREC_BF16X3>(a, grid, queue, progress, nsb, SB, stream);
}
