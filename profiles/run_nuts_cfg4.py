"""BASELINE config 4 on one GPU: NUTS, D = 100, rho = 0.95, 65,536 chains, dt = 0.1, d_max = 10 (on_dmax = "stop"), 4 iterations."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "understanding-hmc_b200"))
import numpy as np
import samplers as S
D, Nc, Niter = 100, int(sys.argv[1]) if len(sys.argv) > 1 else 65536, int(sys.argv[2]) if len(sys.argv) > 2 else 4
spec = S.MVNSpec.from_cov(np.zeros(D), S.equicorrelated_cov(D, 0.95))
q0 = (np.random.RandomState(0).standard_normal((Nc, D)) * 1.4).astype(np.float32)
for rep in range(2):
    H = S.HMC_sampler(D, None, None, Nchain=Nc, Niter=Niter, sampler_type="NUTS", dt=0.1, d_max=10, dtype="float32", seed=1 + rep,
                      target=spec, on_dmax="stop")
    H.gen_sample(q0, verbose=False)
    print("kernel %s: %.1f ms, %d leapfrogs (%.1f per iteration), %.3e leapfrogs/s" % (H.nuts_kernel, H.kernel_ms, H.n_leapfrog_total,
          H.n_leapfrog_total / float(Nc * Niter), H.n_leapfrog_total / (H.kernel_ms * 1e-3)))
