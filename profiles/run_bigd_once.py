"""BASELINE config 5 on one GPU: dense-covariance MVN D = 1024, 131,072 chains, L in [100, 500), dt = 0.1 (large-D GEMM path)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "understanding-hmc_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import samplers as S, utils as U

D = int(os.environ.get("BIGD_D", "1024")); Nc = int(os.environ.get("BIGD_CHAINS", "131072")); Niter = int(os.environ.get("BIGD_NITER", "2"))
Llo, Lhi = int(os.environ.get("BIGD_LLO", "100")), int(os.environ.get("BIGD_LHI", "500"))
rng = np.random.RandomState(0)
lam = np.exp(rng.uniform(np.log(0.05), np.log(100.0), D))
Q, _ = np.linalg.qr(rng.standard_normal((D, D)))
cov = (Q * lam) @ Q.T; cov = 0.5 * (cov + cov.T)
P = (Q / lam) @ Q.T; P = 0.5 * (P + P.T)
spec = S.MVNSpec(np.zeros(D), P, 0.5 * (D * np.log(2 * np.pi) + np.log(lam).sum()))
q0 = U.start_pts(np.zeros(D), np.diag(lam.mean() * np.ones(D)), Nc, device="cuda", seed=1)
for prec in (os.environ.get("BIGD_PREC", "fp16x2,bf16x3").split(",")):
    H = S.HMC_sampler(D, None, None, Nchain=Nc, Niter=Niter, warm_up_num=Niter // 2, sampler_type="Random", dt=0.1, L_low=Llo, L_high=Lhi,
                      dtype="float32", kernel="bigd", tc_precision=prec, seed=3, target=spec)
    t0 = time.time(); H.gen_sample(q0, verbose=False, quiet=True); torch.cuda.synchronize(); wall = time.time() - t0
    ge = H.sum_L + Nc * Niter
    nprod = 3 if prec == "fp16x2" else 6
    print(json.dumps({"config": "D=%d chains=%d Niter=%d L=[%d,%d) %s" % (D, Nc, Niter, Llo, Lhi, prec), "kernel_ms": H.kernel_ms, "wall_s": wall,
                      "accept_R": H.accept_R, "leapfrog_grad_evals_per_sec": H.sum_L / (H.kernel_ms * 1e-3),
                      "algorithmic_tflops": ge * 2.0 * D * D / (H.kernel_ms * 1e-3) / 1e12,
                      "tensor_tflops_executed": ge * 2.0 * D * D * nprod / (H.kernel_ms * 1e-3) / 1e12}), flush=True)
    del H; torch.cuda.empty_cache()
