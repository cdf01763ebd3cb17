"""Small invocations of every sampling kernel for compute-sanitizer (racecheck / memcheck / synccheck).

  compute-sanitizer --tool racecheck python profiles/sanitize_case.py tc
Cases: tc | tc_short (L in {1,2}) | tc_vdt (vector dt, thinning, warm-up) | fast | generic | nuts | diag
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "understanding-hmc_b200")):
    sys.path.insert(0, p)
import numpy as np   # noqa: E402
import samplers as S  # noqa: E402


def main():
    case = sys.argv[1]
    D = 100
    spec = S.MVNSpec.from_cov(np.zeros(D), S.equicorrelated_cov(D, 0.95))
    rng = np.random.RandomState(3)
    if case in ("tc", "tc_short", "tc_vdt", "fast", "generic"):
        Nchain = 300 if case != "generic" else 40
        kw = dict(Nchain=Nchain, Niter=6, thin_rate=1, warm_up_num=0, sampler_type="Random", dt=0.1, L_low=5, L_high=20,
                  dtype="float32", seed=11, target=spec, kernel=case.split("_")[0])
        if case == "tc_short":
            kw.update(L_low=1, L_high=3, Niter=8)
        if case == "tc_vdt":
            kw.update(dt=0.06 + 0.08 * np.arange(D) / D, thin_rate=3, warm_up_num=2, Niter=9, iter_block=4)
        H = S.HMC_sampler(D, None, None, **kw)
        H.gen_sample(rng.standard_normal((Nchain, D)).astype(np.float32) * 1.4, N_save_chain0=2, verbose=False, quiet=True)
        H.compute_convergence_stats()
        print(case, "accept", H.accept_R, "sumL", H.sum_L, "Rhat med", float(np.median(H.R_q)))
    elif case == "nuts":
        Nchain = 24
        H = S.HMC_sampler(D, None, None, Nchain=Nchain, Niter=3, sampler_type="NUTS", dt=0.2, d_max=8, dtype="float32",
                          seed=5, target=spec, on_dmax="stop")
        H.gen_sample(rng.standard_normal((Nchain, D)) * 1.4, verbose=False)
        print(case, "leapfrogs", H.n_leapfrog_total)
    elif case == "diag":
        import torch
        import utils as U
        x = torch.randn((64, 1 + 2 * 150, D), device="cuda")
        R, ne = U.convergence_stats(x, thin_rate=1, warm_up_num=1)
        print(case, float(np.median(R)), float(np.median(ne)))
    else:
        raise SystemExit("unknown case " + case)


if __name__ == "__main__":
    main()
