// Generic random-trajectory-length HMC kernel: ONE WARP PER CHAIN, float or double, D <= 1024,
// dense momentum metric and per-dimension dt supported.  This is the parity workhorse (its float64
// instantiation is compared with the float64 oracle to ~1e-10) and the path for everything the FFMA2
// fast kernel (random_fast.cu) does not cover.  It follows HMC_sampler.gen_sample_random
// (/root/reference/samplers.py:387-491) statement by statement; lane `l` owns dimensions l, l+32, ...
#include "hmc_common.cuh"

namespace {

template <typename T, int NJ, bool SM>
__global__ void __launch_bounds__(128) hmc_random_generic_kernel(hmc_random_args a, int smem_mask) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = a.target.D, Dp = a.target.D_pad;
    const T* Ft = (const T*)a.target.Ft;
    const T* Pt = (const T*)a.target.Pt;
    const T* Mit = (const T*)a.target.Mit;
    const T* Lct = (const T*)a.target.Lct;
    T* xs = (T*)smem_raw + (size_t)(threadIdx.x >> 5) * Dp;   // per-warp staging vector
    {   // stage the matrices that fit into shared memory (decided by the host)
        T* s = (T*)smem_raw + (size_t)(blockDim.x >> 5) * Dp;
        const int n = D * Dp;
        if constexpr (SM) {      // identity metric, F fits: Dp x Dp tile with zero rows D..Dp-1 (no row checks in the mat-vec)
            for (int t = threadIdx.x; t < Dp * Dp; t += blockDim.x) s[t] = (t < n) ? Ft[t] : T(0);
            Ft = s;
        } else if (smem_mask & 1) { for (int t = threadIdx.x; t < n; t += blockDim.x) s[t] = Ft[t]; Ft = s; s += n; }
        if ((smem_mask & 2) && Pt) { for (int t = threadIdx.x; t < n; t += blockDim.x) s[t] = Pt[t]; Pt = s; s += n; }
        if ((smem_mask & 4) && Mit) { for (int t = threadIdx.x; t < n; t += blockDim.x) s[t] = Mit[t]; Mit = s; s += n; }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const long m = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= a.Nchain) return;
    const uint64_t gid = (uint64_t)(a.chain_id0 + m);
    const T* mu_g = (const T*)a.target.mu;
    const T* dt_g = (const T*)a.target.dt;
    T* q_chain = (T*)a.q_chain;
    const long Lc = 1 + (a.Niter - a.warm_up_num) / a.thin_rate;   // samplers.py:31
    const long Lrow = a.store_ring > 0 ? a.store_ring : Lc;         // rows allocated per chain (ring of the last store_ring stored samples)

    T q[NJ], p[NJ], f[NJ], d[NJ], mu[NJ], dt[NJ], tmp[NJ];
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
        const int j = hmc_dim<NJ>(lane, i);
        mu[i] = (j < D) ? mu_g[j] : T(0);
        dt[i] = (j < D) ? dt_g[j] : T(0);
        q[i] = T(0); p[i] = T(0); f[i] = T(0);
    }

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
            if (Lct) {   // p = Lc z ~ N(0, cov_p)   (samplers.py:829)
#pragma unroll
                for (int i = 0; i < NJ; ++i) tmp[i] = p[i];
                matvec_t<T, NJ, false>(Lct, D, Dp, tmp, p, lane, xs);
            }
        }
    };
    auto force = [&]() {          // f = F (q - mu)
#pragma unroll
        for (int i = 0; i < NJ; ++i) d[i] = q[i] - mu[i];
        matvec_t<T, NJ, SM>(Ft, D, Dp, d, f, lane, xs);
    };
    auto potential = [&]() -> double {   // V(q); requires d, f current for q
        if (Pt) { matvec_t<T, NJ, false>(Pt, D, Dp, d, tmp, lane, xs); return 0.5 * dot_warp<T, NJ>(d, tmp) + a.target.v_const; }
        return 0.5 * dot_warp<T, NJ>(d, f) + a.target.v_const;
    };
    auto kinetic = [&]() -> double {
        if (Mit) { matvec_t<T, NJ, false>(Mit, D, Dp, p, tmp, lane, xs); return 0.5 * dot_warp<T, NJ>(p, tmp); }
        return 0.5 * dot_warp<T, NJ>(p, p);
    };

    double E_previous;
    unsigned long long acc_warm = 0, acc_post = 0, sumL = 0, sumL2 = 0;
    if (a.iter_begin == 0) {                                            // samplers.py:413-420
        const T* qs = (const T*)a.q_start + (size_t)m * D;
#pragma unroll
        for (int i = 0; i < NJ; ++i) { const int j = hmc_dim<NJ>(lane, i); if (j < D) { q[i] = qs[j]; q_chain[(size_t)m * Lrow * D + j] = q[i]; } }
        draw_p(0);
        force();
        const double E0 = potential() + kinetic();
        if (lane == 0) { a.E_chain[(size_t)m * Lrow] = E0; a.dE_chain[(size_t)m * Lrow] = 0.0; }
        E_previous = E0;
        if (a.decision_chain && gid == 0 && lane == 0) a.decision_chain[a.N_save_chain0] = 0;
    } else {
        const T* qs = (const T*)a.state_q + (size_t)m * D;
#pragma unroll
        for (int i = 0; i < NJ; ++i) { const int j = hmc_dim<NJ>(lane, i); if (j < D) q[i] = qs[j]; }
        E_previous = a.state_eprev[m];
    }

    T q_init[NJ];
    for (int it = a.iter_begin + 1; it <= a.iter_end; ++it) {           // samplers.py:428
#pragma unroll
        for (int i = 0; i < NJ; ++i) q_init[i] = q[i];
        draw_p(it);                                                     // samplers.py:431
        force();
        const double E_initial = potential() + kinetic();               // samplers.py:434
        const bool keep = it >= a.warm_up_num;
        const long idx = keep ? ((it - a.warm_up_num) / a.thin_rate) % Lrow : 0;
        if (keep && lane == 0) {                                        // samplers.py:436-438
            a.E_chain[(size_t)m * Lrow + idx] = E_initial;
            a.dE_chain[(size_t)m * Lrow + idx] = E_initial - E_previous;
        }
        int L; double u;
        if (a.L_tape) { L = a.L_tape[(size_t)m * a.Niter + it - 1]; u = a.u_tape[(size_t)m * a.Niter + it - 1]; }
        else hmc_scalar_draws(a.seed, gid, (uint32_t)it, a.L_low, a.L_high, &L, &u);
        sumL += L; sumL2 += (unsigned long long)L * L;
        const bool trace = a.phi_q && gid == 0 && it <= a.N_save_chain0;
        double* phi = trace ? a.phi_q + (size_t)(it - 1) * a.L_high * 2 : nullptr;
        auto trace_row = [&](int row) {       // first two coordinates of the current position (samplers.py:445, 452)
            if constexpr (NJ % 4 == 0) { if (lane == 0) { phi[2 * row] = (double)q[0]; if (D > 1) phi[2 * row + 1] = (double)q[1]; } }
            else { if (lane < 2 && lane < D) phi[2 * row + lane] = (double)q[0]; }
        };
        if (trace) { trace_row(0); if (lane == 0) a.phi_len[it - 1] = L + 1; }
        for (int l = 1; l <= L; ++l) {                                  // samplers.py:448-452, leap_frog :831-839
#pragma unroll
            for (int i = 0; i < NJ; ++i) {
                p[i] = p[i] - dt[i] * f[i] / T(2);
                q[i] = q[i] + dt[i] * p[i];
            }
            force();
#pragma unroll
            for (int i = 0; i < NJ; ++i) p[i] = p[i] - dt[i] * f[i] / T(2);
            if (trace) trace_row(l);
        }
        const double E_final = potential() + kinetic();                 // samplers.py:455
        const double dE = E_final - E_initial;
        E_previous = E_initial;                                         // samplers.py:460
        const double lnu = log(u);
        const bool accepted = (dE < 0) || (lnu < -dE);                  // samplers.py:462
        if (accepted) {
            if (trace && lane == 0) a.decision_chain[it - 1] = 1;
            if (keep) acc_post++; else acc_warm++;
        } else {
            if (trace && lane == 0) a.decision_chain[it - 1] = 0;
#pragma unroll
            for (int i = 0; i < NJ; ++i) q[i] = q_init[i];
        }
        if (keep) {                                                     // samplers.py:465-471 (Q4: negative-index writes skipped)
            T* dst = q_chain + ((size_t)m * Lrow + idx) * D;
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
            if (a.counters) {
                atomicAdd(a.counters + 0, acc_warm);
                atomicAdd(a.counters + 1, acc_post);
                atomicAdd(a.counters + 2, sumL);
                atomicAdd(a.counters + 3, sumL2);
            }
        }
    }
}

template <typename T, int NJ>
int launch_generic(const hmc_random_args& a, cudaStream_t stream) {
    const int warps = 4;
    const int blocks = (a.Nchain + warps - 1) / warps;
    const size_t mat = (size_t)a.target.D * a.target.D_pad * sizeof(T);
    int mask = 0;
    size_t smem = (size_t)warps * a.target.D_pad * sizeof(T);
    const size_t budget = 200 * 1024;
    if (smem + mat <= budget) { mask |= 1; smem += mat; }
    if (a.target.Pt && smem + mat <= budget) { mask |= 2; smem += mat; }
    if (a.target.Mit && smem + mat <= budget) { mask |= 4; smem += mat; }
    const size_t matp = (size_t)a.target.D_pad * a.target.D_pad * sizeof(T);
    const size_t smem_sm = (size_t)warps * a.target.D_pad * sizeof(T) + matp;
    if (!a.target.Pt && !a.target.Mit && !a.target.Lct && smem_sm <= budget) {
        auto kern = hmc_random_generic_kernel<T, NJ, true>;
        HMC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sm));
        kern<<<blocks, warps * 32, smem_sm, stream>>>(a, 0);
    } else {
        auto kern = hmc_random_generic_kernel<T, NJ, false>;
        HMC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, warps * 32, smem, stream>>>(a, mask);
    }
    HMC_CUDA_CHECK(cudaGetLastError());
    return HMC_OK;
}

template <typename T>
int dispatch_generic(const hmc_random_args& a, cudaStream_t stream) {
    const int D = a.target.D;
    if (D <= 32) return launch_generic<T, 1>(a, stream);
    if (D <= 128) return launch_generic<T, 4>(a, stream);
    if (D <= 256) return launch_generic<T, 8>(a, stream);
    if (D <= 1024) return launch_generic<T, 32>(a, stream);
    hmc_set_error("generic kernel supports D <= 1024 (got %d)", D);
    return HMC_E_UNSUPPORTED;
}

// HMC_sampler.leap_frog (/root/reference/samplers.py:831-839) for a batch of (p, q) pairs: one warp per pair, `nsteps` steps as
// the reference writes one (force F (q - mu) = M^-1 P (q - mu) at both ends of every step, no reuse between steps: the public
// primitive, not the fused loop).  Matrices are read from global memory / L2 (small batches: a parity and API entry point).
template <typename T, int NJ>
__global__ void __launch_bounds__(128) hmc_leap_frog_kernel(hmc_target t, long B, const T* __restrict__ p_old, const T* __restrict__ q_old,
                                                            T* __restrict__ p_new, T* __restrict__ q_new, int nsteps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = t.D, Dp = t.D_pad, lane = threadIdx.x & 31;
    T* xs = (T*)smem_raw + (size_t)(threadIdx.x >> 5) * Dp;
    const long m = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= B) return;
    const T* Ft = (const T*)t.Ft;
    T q[NJ], p[NJ], f[NJ], d[NJ], mu[NJ], dt[NJ];
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
        const int j = hmc_dim<NJ>(lane, i);
        mu[i] = (j < D) ? ((const T*)t.mu)[j] : T(0);
        dt[i] = (j < D) ? ((const T*)t.dt)[j] : T(0);
        q[i] = (j < D) ? q_old[(size_t)m * D + j] : T(0);
        p[i] = (j < D) ? p_old[(size_t)m * D + j] : T(0);
        f[i] = T(0);
    }
    auto force = [&]() {
#pragma unroll
        for (int i = 0; i < NJ; ++i) d[i] = q[i] - mu[i];
        matvec_t<T, NJ, false>(Ft, D, Dp, d, f, lane, xs);
    };
    for (int s = 0; s < nsteps; ++s) {
        force();
#pragma unroll
        for (int i = 0; i < NJ; ++i) {
            p[i] = p[i] - dt[i] * f[i] / T(2);                          // samplers.py:835
            q[i] = q[i] + dt[i] * p[i];                                 // samplers.py:836
        }
        force();
#pragma unroll
        for (int i = 0; i < NJ; ++i) p[i] = p[i] - dt[i] * f[i] / T(2); // samplers.py:837
    }
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
        const int j = hmc_dim<NJ>(lane, i);
        if (j < D) { p_new[(size_t)m * D + j] = p[i]; q_new[(size_t)m * D + j] = q[i]; }
    }
}

template <typename T, int NJ>
int launch_leap_frog(const hmc_target& t, long B, const void* p_old, const void* q_old, void* p_new, void* q_new, int nsteps, cudaStream_t stream) {
    const int warps = 4;
    const size_t smem = (size_t)warps * t.D_pad * sizeof(T);
    hmc_leap_frog_kernel<T, NJ><<<(unsigned)((B + warps - 1) / warps), warps * 32, smem, stream>>>(t, B, (const T*)p_old, (const T*)q_old, (T*)p_new, (T*)q_new, nsteps);
    HMC_CUDA_CHECK(cudaGetLastError());
    return HMC_OK;
}

template <typename T>
int dispatch_leap_frog(const hmc_target& t, long B, const void* p_old, const void* q_old, void* p_new, void* q_new, int nsteps, cudaStream_t stream) {
    if (t.D <= 32) return launch_leap_frog<T, 1>(t, B, p_old, q_old, p_new, q_new, nsteps, stream);
    if (t.D <= 128) return launch_leap_frog<T, 4>(t, B, p_old, q_old, p_new, q_new, nsteps, stream);
    if (t.D <= 256) return launch_leap_frog<T, 8>(t, B, p_old, q_old, p_new, q_new, nsteps, stream);
    if (t.D <= 1024) return launch_leap_frog<T, 32>(t, B, p_old, q_old, p_new, q_new, nsteps, stream);
    hmc_set_error("hmc_leap_frog supports D <= 1024 (got %d)", t.D);
    return HMC_E_UNSUPPORTED;
}

}  // namespace

extern "C" int hmc_leap_frog(int32_t dtype, const hmc_target* target, int64_t B, const void* p_old, const void* q_old, void* p_new, void* q_new,
                             int32_t nsteps, void* cuda_stream) {
    if (!target || !p_old || !q_old || !p_new || !q_new || !target->Ft || !target->mu || !target->dt) { hmc_set_error("hmc_leap_frog: NULL buffer"); return HMC_E_BADARG; }
    if ((dtype != HMC_F32 && dtype != HMC_F64) || B < 1 || nsteps < 1 || target->D < 1) { hmc_set_error("hmc_leap_frog: need dtype HMC_F32 / HMC_F64, B >= 1, nsteps >= 1"); return HMC_E_BADARG; }
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    if (dtype == HMC_F32) return dispatch_leap_frog<float>(*target, (long)B, p_old, q_old, p_new, q_new, nsteps, stream);
    return dispatch_leap_frog<double>(*target, (long)B, p_old, q_old, p_new, q_new, nsteps, stream);
}

int hmc_random_run_generic(const hmc_random_args& a, cudaStream_t stream) {
    if (a.dtype == HMC_F32) return dispatch_generic<float>(a, stream);
    return dispatch_generic<double>(a, stream);
}
