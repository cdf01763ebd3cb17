"""Small driver for ncu: launches the fused kernel a few times on a one-wave problem (148 SMs x 192 chains)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "understanding-hmc_b200"))
import numpy as np, torch
import hmc_b200_lib as L, samplers as S
D, Nc, IB, NL = 100, int(sys.argv[1]) if len(sys.argv) > 1 else 28416, int(sys.argv[2]) if len(sys.argv) > 2 else 10, 3
spec = S.MVNSpec.from_cov(np.zeros(D), S.equicorrelated_cov(D, 0.95))
q0 = (np.random.RandomState(0).standard_normal((Nc, D)) * 1.4).astype(np.float32)
H = S.HMC_sampler(D, None, None, Nchain=Nc, Niter=IB * NL, sampler_type="Random", dt=0.1, L_low=5, L_high=20,
                  dtype="float32", kernel=os.environ.get("HMC_B200_KERNEL", "fast"), seed=1, target=spec)
run = H.prepare_random(q0)
lib = L.load()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(NL + 1)]
ev[0].record()
for i in range(NL):
    run["args"].iter_begin, run["args"].iter_end = i * IB, (i + 1) * IB
    L.check(lib.hmc_random_run(run["args"], L.current_stream_ptr()))
    ev[i + 1].record()
torch.cuda.synchronize()
c = run["counters"].cpu().numpy()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(NL)]
print("launch ms", ms, "sumL", int(c[2]), "grad-evals/s (last)", c[2] / NL / (ms[-1] * 1e-3))
