// NUTS kernel: ONE WARP PER CHAIN runs the reference's iterative tree doubling
// (HMC_sampler.gen_sample_NUTS, /root/reference/samplers.py:495-808) as a per-chain state machine; float or
// double; identity or dense momentum metric (cov_p != I: the force is M^-1 P (q - mu) as the reference writes it, Q9,
// V and K take their own mat-vecs with P and M^-1, momenta are Lc z; samplers.py:352-356, 811-817, 825-829, 835-837).  Lane l owns dimensions l, l+32, ...; control flow is warp-uniform (every
// per-chain scalar is computed redundantly by all lanes, in float64 -- SURVEY H5: the progressive-sampling
// weights exp(E_max - E) overflow float32).
//
// The reference's slot table (find_next / retrieve_save_index, utils.py:222-237) and its index rules
// (check_points / release_fast, utils.py:246-304, 367-385) are replaced by closed forms (pinned by
// tests/test_oracle_golden.py against the reference's own functions for m <= 1024):
//   points checked at even m :  l = m - 2^j + 1,  j = tz(m) .. 1
//   slot of saved point l    :  popcount((l-1) >> 1)   (point 1 -> slot 0); never collides among live points
// so the check-point stack needs no search and no explicit release.
//
// Per-chain scratch (HBM, L2-resident): rows of D_pad elements
//   [0, d_max]            q of the saved points        [d_max+1, 2 d_max+1]  p of the saved points
//   R+0 live_new q   R+1 live_old q   R+2 left q   R+3 left p   R+4 right q   R+5 right p      (R = 2 (d_max+1))
//   R+6 cursors of the injected draw streams (so that a run can be split over several launches)
#include "hmc_common.cuh"

namespace {

enum { HMC_STREAM_NUTS_INNER = 3 };

template <typename T, int NJ, bool SM>
__global__ void __launch_bounds__(128) hmc_nuts_generic_kernel(hmc_nuts_args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = a.target.D, Dp = a.target.D_pad;
    const T* Ft = (const T*)a.target.Ft;
    const T* Pt = (const T*)a.target.Pt;        // NULL for the identity metric (then F == P)
    const T* Mit = (const T*)a.target.Mit;
    const T* Lct = (const T*)a.target.Lct;
    T* xs = (T*)smem_raw + (size_t)(threadIdx.x >> 5) * Dp;
    {
        T* s = (T*)smem_raw + (size_t)(blockDim.x >> 5) * Dp;
        if constexpr (SM) {      // Dp x Dp tile: the rows D..Dp-1 are zero so that the mat-vec needs no row checks
            for (int t = threadIdx.x; t < Dp * Dp; t += blockDim.x) s[t] = (t < D * Dp) ? Ft[t] : T(0);
            Ft = s;
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const long m = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= a.Nchain) return;
    const uint64_t gid = (uint64_t)(a.chain_id0 + m);
    const T* mu_g = (const T*)a.target.mu;
    const T* dt_g = (const T*)a.target.dt;
    T* q_chain = (T*)a.q_chain;
    const long Lc = 1 + (a.Niter - a.warm_up_num) / a.thin_rate;
    const int R = 2 * (a.d_max + 1);
    T* scr = (T*)a.scratch + (size_t)m * (R + 7) * Dp;
    long long* cursors = reinterpret_cast<long long*>(scr + (size_t)(R + 6) * Dp);
    const double vconst = a.target.v_const;

    T q[NJ], p[NJ], f[NJ], d[NJ], mu[NJ], dt[NJ], t1[NJ], t2[NJ];
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
        const int j = hmc_dim<NJ>(lane, i);
        mu[i] = (j < D) ? mu_g[j] : T(0);
        dt[i] = (j < D) ? dt_g[j] : T(0);
        q[i] = T(0); p[i] = T(0); f[i] = T(0);
    }
    auto row_store = [&](int row, const T (&v)[NJ]) {
#pragma unroll
        for (int i = 0; i < NJ; ++i) { const int j = hmc_dim<NJ>(lane, i); if (j < D) scr[(size_t)row * Dp + j] = v[i]; }
    };
    auto row_load = [&](int row, T (&v)[NJ]) {
#pragma unroll
        for (int i = 0; i < NJ; ++i) { const int j = hmc_dim<NJ>(lane, i); v[i] = (j < D) ? scr[(size_t)row * Dp + j] : T(0); }
    };
    auto draw_p = [&](int iter) {
        if (a.p_tape) {
            const double* src = a.p_tape + ((size_t)m * (a.Niter + 1) + iter) * D;
#pragma unroll
            for (int i = 0; i < NJ; ++i) { const int j = hmc_dim<NJ>(lane, i); p[i] = (j < D) ? (T)src[j] : T(0); }
        } else {
#pragma unroll
            for (int i = 0; i < NJ; ++i) {
                const int j = hmc_dim<NJ>(lane, i);
                if (j < D) {
                    float4 z = hmc_normal4(a.seed, gid, (uint32_t)iter, (uint32_t)(j >> 2));
                    const int r = j & 3;
                    p[i] = (T)(r == 0 ? z.x : r == 1 ? z.y : r == 2 ? z.z : z.w);
                } else p[i] = T(0);
            }
            if (Lct) {                        // p = Lc z ~ N(0, cov_p)   (samplers.py:829)
#pragma unroll
                for (int i = 0; i < NJ; ++i) t1[i] = p[i];
                matvec_t<T, NJ, false>(Lct, D, Dp, t1, p, lane, xs);
            }
        }
    };
    auto force = [&]() {                      // f = M^-1 P (q - mu)
#pragma unroll
        for (int i = 0; i < NJ; ++i) d[i] = q[i] - mu[i];
        matvec_t<T, NJ, SM>(Ft, D, Dp, d, f, lane, xs);
    };
    auto energy = [&]() -> double {           // E(q, p) with d, f current (samplers.py:819-823): one warp reduction
        T s = T(0);
        if (Pt) {                             // dense metric: V = 0.5 d.P d, K = 0.5 p.M^-1 p (samplers.py:811-817)
            matvec_t<T, NJ, false>(Pt, D, Dp, d, t1, lane, xs);
            matvec_t<T, NJ, false>(Mit, D, Dp, p, t2, lane, xs);
#pragma unroll
            for (int i = 0; i < NJ; ++i) s = fma(d[i], t1[i], fma(p[i], t2[i], s));
        } else {
#pragma unroll
            for (int i = 0; i < NJ; ++i) s = fma(d[i], f[i], fma(p[i], p[i], s));
        }
        return 0.5 * (double)warp_sum<T>(s) + vconst;
    };
    auto leap = [&]() {                       // one leapfrog step, f = force at q on entry and on exit (samplers.py:831-839)
#pragma unroll
        for (int i = 0; i < NJ; ++i) { p[i] = p[i] - dt[i] * f[i] / T(2); q[i] = q[i] + dt[i] * p[i]; }
        force();
#pragma unroll
        for (int i = 0; i < NJ; ++i) p[i] = p[i] - dt[i] * f[i] / T(2);
    };
    // injected streams are consumed in order per chain; Philox draws are keyed by position (iteration, depth, step)
    long n_dir = 0, n_u = 0;
    auto draw_dir = [&](int iter, int depth, double* u_biased) -> int {
        if (a.dir_tape) return a.dir_tape[(size_t)m * a.tape_dir_stride + n_dir++];
        const Philox4 r = philox4x32_10((uint32_t)gid, (uint32_t)iter, (uint32_t)depth,
                                        HMC_STREAM_NUTS | ((uint32_t)(gid >> 32) << 8), (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
        *u_biased = ((double)(r.y >> 8) + 0.5) * 5.9604644775390625e-08;
        return (int)(r.x >> 31);
    };
    Philox4 u_cache; u_cache.x = u_cache.y = u_cache.z = u_cache.w = 0u;
    uint32_t u_cache_slot = 0xffffffffu;
    auto draw_u_inner = [&](int iter, int depth, int k) -> double {
        if (a.u_tape) return a.u_tape[(size_t)m * a.tape_u_stride + n_u++];
        // step k of the sub-trajectory of depth d uses word (k & 3) of Philox counter ((1 << d) + k) >> 2
        const uint32_t slot = ((uint32_t)iter << 22) ^ (((1u << depth) + (uint32_t)k) >> 2);
        if (slot != u_cache_slot) {
            u_cache = philox4x32_10((uint32_t)gid, (uint32_t)iter, ((1u << depth) + (uint32_t)k) >> 2,
                                    HMC_STREAM_NUTS_INNER | ((uint32_t)(gid >> 32) << 8), (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
            u_cache_slot = slot;
        }
        const uint32_t w = (k & 3) == 0 ? u_cache.x : (k & 3) == 1 ? u_cache.y : (k & 3) == 2 ? u_cache.z : u_cache.w;
        return ((double)(w >> 8) + 0.5) * 5.9604644775390625e-08;
    };

    double E_previous;
    unsigned long long c_leap = 0, c_doubling = 0, c_instab = 0, c_dmax = 0;
    if (a.iter_begin == 0) {                                            // samplers.py:548-555
        const T* qs = (const T*)a.q_start + (size_t)m * D;
#pragma unroll
        for (int i = 0; i < NJ; ++i) { const int j = hmc_dim<NJ>(lane, i); if (j < D) { q[i] = qs[j]; q_chain[(size_t)m * Lc * D + j] = q[i]; } }
        draw_p(0);
        force();
        const double E0 = energy();
        if (lane == 0) { a.E_chain[(size_t)m * Lc] = E0; a.dE_chain[(size_t)m * Lc] = 0.0; }
        E_previous = E0;
    } else {
        const T* qs = (const T*)a.state_q + (size_t)m * D;
#pragma unroll
        for (int i = 0; i < NJ; ++i) { const int j = hmc_dim<NJ>(lane, i); if (j < D) q[i] = qs[j]; }
        E_previous = a.state_eprev[m];
        if (a.dir_tape) { n_dir = (long)cursors[0]; n_u = (long)cursors[1]; }   // resume the injected streams
    }

    for (int it = a.iter_begin + 1; it <= a.iter_end; ++it) {           // samplers.py:563
        draw_p(it);                                                     // samplers.py:565
        force();
        const double E_initial = energy();                              // samplers.py:569
        const bool keep = it >= a.warm_up_num;
        const long idx = keep ? (it - a.warm_up_num) / a.thin_rate : 0;
        if (keep && lane == 0) {                                        // samplers.py:571-573
            a.E_chain[(size_t)m * Lc + idx] = E_initial;
            a.dE_chain[(size_t)m * Lc + idx] = E_initial - E_previous;
        }
        // samplers.py:577-594: live point, both boundary points, running maxima
        row_store(R + 1, q);                                            // live_old
        row_store(R + 2, q);                                            // left q
#pragma unroll
        for (int i = 0; i < NJ; ++i) t1[i] = -p[i];
        row_store(R + 3, t1);                                           // left p = -p
        row_store(R + 4, q);                                            // right q
        row_store(R + 5, p);                                            // right p
        double E_max_old = E_initial, pi_old = 1.0;
        bool left_term = false, right_term = false;
        int depth = 0;
        __syncwarp();
        while (!(left_term && right_term)) {                            // samplers.py:595 (Q6: both ends)
            if (depth > a.d_max - 1) {                                  // samplers.py:596-598 (Q7)
                c_dmax++;
                if (a.status && lane == 0) a.status[m] |= 1;
                break;
            }
            c_doubling++;
            const int L_new_sub = 1 << depth;                           // samplers.py:604
            double u_biased = 0.0;
            const int u_dir = draw_dir(it, depth, &u_biased);           // samplers.py:608
            if (u_dir == 0) { row_load(R + 4, q); row_load(R + 5, p); } // samplers.py:611-614
            else { row_load(R + 2, q); row_load(R + 3, p); }
            force();
            leap();
            c_leap++;
            row_store(R + 0, q);                                        // live_new = first point
            double E_max_new_now = energy();                            // samplers.py:618
            double pi_new = 1.0;
            row_store(0, q);                                            // point 1 -> slot 0 (samplers.py:623-626)
            row_store(a.d_max + 1, p);
            bool reject = false;
            for (int k = 1; k < L_new_sub; ++k) {                       // samplers.py:637
                leap();
                c_leap++;
                const double E_tmp = energy();                          // samplers.py:643
                if (fabs(E_tmp - E_initial) > 1000.0) {                 // samplers.py:647-651
                    reject = true;
                    c_instab++;
                    break;
                }
                const int pt = k + 1;
                if (pt & 1) {                                           // odd point: save (samplers.py:654-658)
                    const int slot = __popc((unsigned)(pt - 1) >> 1);
                    row_store(slot, q);
                    row_store(a.d_max + 1 + slot, p);
                } else {                                                // even point: sub-tree U-turn checks (:699-736)
                    __syncwarp();
                    const int tz = __ffs(pt) - 1;
                    for (int j = tz; j >= 1; --j) {
                        const int l = pt - (1 << j) + 1;
                        const int slot = __popc((unsigned)(l - 1) >> 1);
                        row_load(slot, t1);                             // q_check
                        row_load(a.d_max + 1 + slot, t2);               // p_check
                        T s_cur = T(0), s_chk = T(0);                   // (q - q_check).p  and  (q - q_check).p_check
#pragma unroll
                        for (int i = 0; i < NJ; ++i) { const T dq = q[i] - t1[i]; s_cur = fma(dq, p[i], s_cur); s_chk = fma(dq, t2[i], s_chk); }
                        const T a_cur = warp_sum<T>(s_cur), a_chk = warp_sum<T>(s_chk);
                        // forward : Dq = q - q_chk, right_p = p, left_p = -p_chk   -> (Dq.p < 0) and (Dq.p_chk < 0)
                        // backward: Dq = q_chk - q, right_p = -p_chk, left_p = p   -> (Dq.(-p_chk) < 0) and (-Dq.p < 0)
                        //           i.e. ((q - q_chk).p_chk < 0) and ((q - q_chk).p < 0): the same two signs.
                        if (a_cur < 0 && a_chk < 0) { reject = true; break; }
                    }
                    if (reject) break;
                }
                // uniform progressive sampling inside the new sub-trajectory (samplers.py:743-751)
                // numerator = exp(E_max - E_tmp) and the rescaling exp(E_max - E_max_previous): one of the two
                // exponents is exactly 0, so a single exp() gives both factors
                const double E_prev_max = E_max_new_now;
                E_max_new_now = fmax(E_prev_max, E_tmp);
                const double ex = exp(fabs(E_tmp - E_prev_max));
                const bool new_max = E_tmp > E_prev_max;
                const double numer = new_max ? 1.0 : ex;
                pi_new = numer + (new_max ? ex : 1.0) * pi_new;
                const double r = numer / pi_new;
                const double u = draw_u_inner(it, depth, k);
                if (u < r) row_store(R + 0, q);
            }
            if (reject) break;                                          // samplers.py:754-755 (sample stays live_old)
            if (u_dir == 0) { row_store(R + 4, q); row_store(R + 5, p); }   // samplers.py:758-761
            else { row_store(R + 2, q); row_store(R + 3, p); }
            // biased choice between the old trajectory and the new sub-trajectory (samplers.py:766-776, Q8)
            const double rb = exp(-(E_max_new_now - E_max_old)) * pi_old / pi_new;
            const double E_max_old_prev = E_max_old;
            E_max_old = fmax(E_max_old_prev, E_max_new_now);
            pi_old = exp(-(E_max_new_now - E_max_old)) * pi_new + exp(-(E_max_old_prev - E_max_old)) * pi_old;
            const double A = fmin(1.0, rb);
            const double ub = a.u_tape ? a.u_tape[(size_t)m * a.tape_u_stride + n_u++] : u_biased;
            __syncwarp();
            if (ub < A) { row_load(R + 0, t1); row_store(R + 1, t1); }
            // whole-trajectory U-turn test (samplers.py:779-781)
            __syncwarp();
            row_load(R + 2, t1);                                        // left q
            row_load(R + 4, t2);                                        // right q
            T s_r = T(0), s_l = T(0);
            {
                T lp[NJ], rp[NJ];
                row_load(R + 3, lp);
                row_load(R + 5, rp);
#pragma unroll
                for (int i = 0; i < NJ; ++i) { const T dq = t2[i] - t1[i]; s_r = fma(dq, rp[i], s_r); s_l = fma(-dq, lp[i], s_l); }
            }
            right_term = warp_sum<T>(s_r) < T(0);
            left_term = warp_sum<T>(s_l) < T(0);
            depth += 1;                                                 // samplers.py:784
        }
        __syncwarp();
        row_load(R + 1, q);                                             // q_tmp = live_point_q_old
        E_previous = E_initial;                                         // samplers.py:787
        if (keep) {                                                     // samplers.py:790-791
            T* dst = q_chain + ((size_t)m * Lc + idx) * D;
#pragma unroll
            for (int i = 0; i < NJ; ++i) { const int j = hmc_dim<NJ>(lane, i); if (j < D) dst[j] = q[i]; }
        }
    }
    {
        T* qs = (T*)a.state_q + (size_t)m * D;
#pragma unroll
        for (int i = 0; i < NJ; ++i) { const int j = hmc_dim<NJ>(lane, i); if (j < D) qs[j] = q[i]; }
        if (lane == 0) {
            a.state_eprev[m] = E_previous;
            cursors[0] = n_dir; cursors[1] = n_u;
            if (a.n_leapfrog) a.n_leapfrog[m] += (int64_t)c_leap;
            atomicAdd(a.counters + 0, c_leap);
            atomicAdd(a.counters + 1, c_doubling);
            atomicAdd(a.counters + 2, c_instab);
            atomicAdd(a.counters + 3, c_dmax);
        }
    }
}

template <typename T, int NJ>
int launch_nuts(const hmc_nuts_args& a, cudaStream_t stream) {
    const int warps = 4;
    const int blocks = (a.Nchain + warps - 1) / warps;
    const size_t mat = (size_t)a.target.D * a.target.D_pad * sizeof(T);
    size_t smem = (size_t)warps * a.target.D_pad * sizeof(T);
    const size_t matp = (size_t)a.target.D_pad * a.target.D_pad * sizeof(T);
    (void)mat;
    if (smem + matp <= 200 * 1024) {
        smem += matp;
        auto kern = hmc_nuts_generic_kernel<T, NJ, true>;
        HMC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, warps * 32, smem, stream>>>(a);
    } else {
        auto kern = hmc_nuts_generic_kernel<T, NJ, false>;
        HMC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, warps * 32, smem, stream>>>(a);
    }
    HMC_CUDA_CHECK(cudaGetLastError());
    return HMC_OK;
}

template <typename T>
int dispatch_nuts(const hmc_nuts_args& a, cudaStream_t stream) {
    const int D = a.target.D;
    if (D <= 32) return launch_nuts<T, 1>(a, stream);
    if (D <= 128) return launch_nuts<T, 4>(a, stream);
    if (D <= 256) return launch_nuts<T, 8>(a, stream);
    hmc_set_error("NUTS kernel supports D <= 256 in this build (got %d)", D);
    return HMC_E_UNSUPPORTED;
}

}  // namespace

int hmc_nuts_run_generic(const hmc_nuts_args& a, cudaStream_t stream) {
    if (a.dtype == HMC_F32) return dispatch_nuts<float>(a, stream);
    return dispatch_nuts<double>(a, stream);
}
