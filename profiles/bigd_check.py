"""Large-D path from an out-of-equilibrium start, L in [100, 500): acceptance and energies against the float64 generic kernel on
the same Philox draws (both splits)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "understanding-hmc_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import samplers as S, utils as U
D, Nc, Niter = 1024, 256, 2
rng = np.random.RandomState(0)
lam = np.exp(rng.uniform(np.log(0.05), np.log(100.0), D))
Q, _ = np.linalg.qr(rng.standard_normal((D, D)))
P = (Q / lam) @ Q.T; P = 0.5 * (P + P.T)
spec = S.MVNSpec(np.zeros(D), P, 0.5 * (D * np.log(2 * np.pi) + np.log(lam).sum()))
q0 = U.start_pts(np.zeros(D), np.diag(lam.mean() * np.ones(D)), Nc, device="cuda", seed=1)
kw = dict(Nchain=Nc, Niter=Niter, warm_up_num=1, sampler_type="Random", dt=0.1, L_low=100, L_high=500, seed=3, target=spec)
G = S.HMC_sampler(D, None, None, dtype="float64", kernel="generic", **kw); G.gen_sample(q0.double(), verbose=False, quiet=True)
print("generic f64: accept", G.accept_R, "E[0,:]", G.E_chain[0, :, 0], "dE", G.dE_chain[0, :, 0])
for prec in ("fp16x2", "bf16x3"):
    H = S.HMC_sampler(D, None, None, dtype="float32", kernel="bigd", tc_precision=prec, **kw); H.gen_sample(q0, verbose=False, quiet=True)
    rel = np.linalg.norm(H.q_chain[:, -1] - G.q_chain[:, -1], axis=1) / np.linalg.norm(G.q_chain[:, -1], axis=1)
    print(prec, ": accept", H.accept_R, "sumL", H.sum_L, G.sum_L, "E[0,:]", H.E_chain[0, :, 0], "| last-sample rel err median %.2e max %.2e" % (np.median(rel), rel.max()),
          "| E rel err max %.2e" % np.max(np.abs(H.E_chain - G.E_chain) / np.abs(G.E_chain)))
