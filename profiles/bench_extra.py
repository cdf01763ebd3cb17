"""Secondary measurements (not the bench.py contract line): diagnostics kernels against the HBM roofline, NUTS
kernel throughput (BASELINE config 4), Case 2c burn-in (config 3, one GPU's share), Case 3d ESS/sec.
Prints one JSON object per measurement."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "understanding-hmc_b200"))
import numpy as np, torch
import hmc_b200_lib as L, samplers as S, utils as U

lib = L.load()
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
HBM = float(peaks["hbm_gbs"])
which = sys.argv[1:] or ["diag", "nuts", "case2c", "case3d"]


def ev_time(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

if "diag" in which:
    # a C2-sized stream: 65,536 chains x 1000 stored samples x 100 dims, float32 = 26.2 GB (AR(1) data made on the device)
    Nc, N, D = 65536, 1000, 100
    x = torch.empty((Nc, N, D), dtype=torch.float32, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    x[:, 0] = torch.randn((Nc, D), device="cuda", generator=g)
    for t in range(1, N):
        x[:, t] = 0.9 * x[:, t - 1] + 0.4359 * torch.randn((Nc, D), device="cuda", generator=g)
    n = N // 2
    mom = torch.empty((4, D), dtype=torch.float64, device="cuda")
    buf = torch.empty((32, D), dtype=torch.float64, device="cuda")
    st = L.current_stream_ptr()
    ms_m = ev_time(lambda: L.check(lib.hmc_diag_moments(L.HMC_F32, L.ptr(x), Nc, n, D, x.stride(0), L.ptr(mom), st)))
    ms_v = ev_time(lambda: L.check(lib.hmc_diag_variogram(L.HMC_F32, L.ptr(x), Nc, n, D, x.stride(0), 1, 32, L.ptr(buf), st)))
    ms_v2 = ev_time(lambda: L.check(lib.hmc_diag_variogram(L.HMC_F32, L.ptr(x), Nc, n, D, x.stride(0), 129, 32, L.ptr(buf), st)))
    gb = Nc * N * D * 4 / 1e9
    t0 = time.perf_counter(); R, ne = U.convergence_stats(x, thin_rate=1, warm_up_num=0); torch.cuda.synchronize(); wall = time.perf_counter() - t0
    print(json.dumps({"what": "diagnostics on 65536x1000x100 f32 (26.2 GB)", "moments_ms": ms_m, "moments_GBs": gb / ms_m * 1e3,
                      "moments_frac_of_measured_hbm": gb / ms_m * 1e3 / HBM, "variogram32_lag1_ms": ms_v,
                      "variogram32_lag1_GBs_algorithmic": gb / ms_v * 1e3, "variogram32_lag129_ms": ms_v2,
                      "variogram_flop_per_s": 3.0 * 32 * Nc * N * D / (ms_v * 1e-3), "convergence_stats_wall_s": wall,
                      "rhat_median": float(np.median(R)), "n_eff_median": float(np.median(ne)), "hbm_peak_GBs": HBM}))
    del x

if "nuts" in which:
    D, Nc, Niter = 100, 65536, 10
    spec = S.MVNSpec.from_cov(np.zeros(D), S.equicorrelated_cov(D, 0.95))
    q0 = (np.random.RandomState(0).standard_normal((Nc, D)) * 1.4).astype(np.float32)
    for dt in (0.2, 0.1):
        H = S.HMC_sampler(D, None, None, Nchain=Nc, Niter=Niter, sampler_type="NUTS", dt=dt, d_max=10, dtype="float32",
                          seed=1, target=spec, on_dmax="stop")
        H.gen_sample(q0, verbose=False)
        print(json.dumps({"what": "NUTS D=100 rho=0.95 65536 chains, dt=%g, d_max=10, on_dmax=stop" % dt, "kernel_ms": H.kernel_ms,
                          "leapfrogs": H.n_leapfrog_total, "leapfrogs_per_iteration": H.n_leapfrog_total / float(Nc * Niter),
                          "grad_evals_per_s": H.n_leapfrog_total / (H.kernel_ms * 1e-3), "dmax_hits": H.n_dmax,
                          "tflops": H.n_leapfrog_total * 2e4 / (H.kernel_ms * 1e-3) / 1e12}))

if "case2c" in which:
    # Case 2c (case2-script.py:136-181): unit MVN D=100, start ~ N(0, 100 I), chain 0 at (1000, -750, 0, ...)
    D, Nc, Niter, warm = 100, 65536, 400, 200
    spec = S.MVNSpec.from_cov(np.zeros(D), np.eye(D))
    q0 = (np.random.RandomState(0).standard_normal((Nc, D)) * 10).astype(np.float32)
    q0[0] = 0; q0[0, 0] = 1000; q0[0, 1] = -750
    H = S.HMC_sampler(D, None, None, Nchain=Nc, Niter=Niter, warm_up_num=warm, sampler_type="Random", dt=0.1, L_low=5,
                      L_high=20, dtype="float32", seed=2, target=spec)
    t0 = time.perf_counter(); H.gen_sample(q0, verbose=False, quiet=True); H.compute_convergence_stats(); wall = time.perf_counter() - t0
    x = H.q_chain_device[:, 1:, :]
    print(json.dumps({"what": "Case 2c share of one GPU: 65536 chains, 400 iterations, warm-up 200", "kernel_ms": H.kernel_ms,
                      "grad_evals_per_s": H.sum_L / (H.kernel_ms * 1e-3), "accept_R": H.accept_R, "accept_R_warm_up": H.accept_R_warm_up,
                      "rhat_median": float(np.median(H.R_q)), "rhat_max": float(np.max(H.R_q)), "n_eff_median": float(np.median(H.n_eff_q)),
                      "sample_mean_abs_max": float(x.mean(dim=(0, 1)).abs().max()), "sample_std_min": float(x.std(dim=(0, 1)).min()),
                      "sample_std_max": float(x.std(dim=(0, 1)).max()), "wall_s": wall}))

if "case3d" in which:
    # Case 3d (case3-script-2.py): D=100 rho=0.95, L in [50,200): the configuration with a meaningful ESS/sec
    D, Nc, Niter, warm = 100, 65536, 300, 100
    spec = S.MVNSpec.from_cov(np.zeros(D), S.equicorrelated_cov(D, 0.95))
    q0 = (np.random.RandomState(0).standard_normal((Nc, D)) * 1.4).astype(np.float32)
    H = S.HMC_sampler(D, None, None, Nchain=Nc, Niter=Niter, warm_up_num=warm, sampler_type="Random", dt=0.1, L_low=50,
                      L_high=200, dtype="float32", seed=3, target=spec)
    t0 = time.perf_counter(); H.gen_sample(q0, verbose=False, quiet=True); t1 = time.perf_counter(); H.compute_convergence_stats(); wall = time.perf_counter() - t0
    x = H.q_chain_device[:, 1:, :]
    print(json.dumps({"what": "Case 3d: D=100 rho=0.95 L in [50,200), 65536 chains, 300 iterations, warm-up 100", "kernel_ms": H.kernel_ms,
                      "grad_evals_per_s": H.sum_L / (H.kernel_ms * 1e-3), "accept_R": H.accept_R, "rhat_median": float(np.median(H.R_q)),
                      "n_eff_median": float(np.median(H.n_eff_q)), "stored": Nc * (H.L_chain - 1),
                      "ess_per_stored": float(np.median(H.n_eff_q)) / (Nc * (H.L_chain - 1)),
                      "ess_per_sec_gen_sample": float(np.median(H.n_eff_q)) / (t1 - t0), "ess_per_sec_incl_stats": float(np.median(H.n_eff_q)) / wall,
                      "sample_std_median": float(x.std(dim=(0, 1)).median()), "wall_s": wall}))
