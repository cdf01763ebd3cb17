"""Error statistics of the float32 kernels against the batched oracle (teacher-forced single iterations)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "understanding-hmc_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
from oracle import hmc_oracle as O
import samplers as S
from test_tc_scale_gpu import _case, D

B = 24000
for case in ("case3c", "case2c"):
    rng = np.random.RandomState(2026)
    tgt, q_init = _case(case, B, rng)
    p = rng.standard_normal((B, D)).astype(np.float32).astype(float)
    L = rng.randint(5, 20, size=B).astype(np.int32)
    u = np.full(B, 1e-300)
    want = O.one_iteration_batch(tgt, q_init, p, L, u, 0.1)
    p_tape = np.zeros((B, 2, D)); p_tape[:, 1] = p
    for kernel, prec, env in (("tc", "fp16x2", None), ("tc", "bf16x3", None), ("fast", "bf16x3", None), ("generic", "bf16x3", None)):
        os.environ.pop("HMC_B200_TC_PREC", None)
        if env:
            os.environ["HMC_B200_TC_PREC"] = env
        H = S.HMC_sampler(D, None, None, Nchain=B, Niter=1, sampler_type="Random", dt=0.1, L_low=5, L_high=20, dtype="float32",
                          kernel=kernel, tc_precision=prec, target=S.MVNSpec.from_cov(tgt.q0, tgt.cov0),
                          draws=dict(p_tape=p_tape, L_tape=L.reshape(B, 1), u_tape=u.reshape(B, 1)))
        H.gen_sample(q_init, verbose=False, quiet=True)
        got = H.q_chain[:, 1, :]
        amp = np.maximum(np.linalg.norm(want["q_prop"], axis=1), np.linalg.norm(q_init, axis=1))
        rel = np.linalg.norm(got - want["q_prop"], axis=1) / amp
        eE = H.E_chain[:, 1, 0] - want["E_init"]
        relE = np.abs(eE) / np.maximum(1.0, np.abs(want["E_init"]))
        half = B // 2
        print("%s %-8s %-7s q rel-L2 med %.2e p99 %.2e max %.2e | E_init abs err: start-dist med %.2e max %.2e, typical med %.2e max %.2e, mean signed %.2e | relE max %.2e"
              % (case, kernel, prec + (":" + env if env else ""), np.median(rel), np.quantile(rel, .99), rel.max(), np.median(np.abs(eE[:half])), np.abs(eE[:half]).max(),
                 np.median(np.abs(eE[half:])), np.abs(eE[half:]).max(), eE[half:].mean(), relE.max()), flush=True)
