"""Small driver for ncu: one NUTS launch (D=100, rho=0.95, dt=0.2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "understanding-hmc_b200"))
import numpy as np
import samplers as S
D, Nc, Niter = 100, int(sys.argv[1]) if len(sys.argv) > 1 else 9472, int(sys.argv[2]) if len(sys.argv) > 2 else 3
spec = S.MVNSpec.from_cov(np.zeros(D), S.equicorrelated_cov(D, 0.95))
q0 = (np.random.RandomState(0).standard_normal((Nc, D)) * 1.4).astype(np.float32)
H = S.HMC_sampler(D, None, None, Nchain=Nc, Niter=Niter, sampler_type="NUTS", dt=0.2, d_max=10, dtype="float32", seed=1,
                  target=spec, on_dmax="stop")
H.gen_sample(q0, verbose=False)
print("kernel ms", H.kernel_ms, "leapfrogs", H.n_leapfrog_total, "per s", H.n_leapfrog_total / (H.kernel_ms * 1e-3))
