"""Where the end-to-end step goes (bench.py's e2e leg): host-side timing of the public-API calls with a synchronize
after each stage.  Usage: python profiles/e2e_breakdown.py [chains] [iterations]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "understanding-hmc_b200"))
import numpy as np, torch
import samplers as S, utils as U
D, Nc, IB = 100, int(sys.argv[1]) if len(sys.argv) > 1 else 65536, int(sys.argv[2]) if len(sys.argv) > 2 else 50
spec = S.MVNSpec.from_cov(np.zeros(D), S.equicorrelated_cov(D, 0.95))
q0 = torch.from_numpy((np.random.RandomState(0).standard_normal((Nc, D)) * 1.4).astype(np.float32)).pin_memory()
def sync(): torch.cuda.synchronize()
for rep in range(4):
    H = S.HMC_sampler(D, None, None, Nchain=Nc, Niter=IB, thin_rate=1, warm_up_num=0, sampler_type="Random", dt=0.1,
                      L_low=5, L_high=20, dtype="float32", kernel="auto", seed=5 + rep, target=spec)
    sync(); t0 = time.perf_counter()
    run = H.prepare_random(q0); sync(); t1 = time.perf_counter()
    del run
    H.gen_sample(q0, verbose=False, quiet=True); sync(); t2 = time.perf_counter()
    H.compute_convergence_stats(); sync(); t3 = time.perf_counter()
    print("rep %d: prepare_random alone %.2f ms | gen_sample (prepare + kernel %.2f ms + counters) %.2f ms | convergence stats %.2f ms"
          % (rep, (t1 - t0) * 1e3, H.kernel_ms, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
    del H
