"""Cycles per phase of the tensor-core NUTS kernel's worker loop (build: nuts_tc.cu with -DHMC_PROFILE_PHASES -> libhmc_b200_prof.so)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "understanding-hmc_b200"))
import hmc_b200_lib as L
L.LIB_PATH = os.path.join(ROOT, "understanding-hmc_b200", "libhmc_b200_prof.so")
import numpy as np, torch, samplers as S, utils as U
import io, contextlib
D, Nc, Niter = 100, int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 4
spec = S.MVNSpec.from_cov(np.zeros(D), S.equicorrelated_cov(D, 0.95))
lib = L.load()
lib.hmc_debug_nuts_tc_cycles.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
q0 = U.start_pts(np.zeros(D), 2.0 * np.eye(D), Nc, device="cuda", seed=93)
for rep in range(2):
    lib.hmc_debug_nuts_tc_cycles(None, 1)
    H = S.HMC_sampler(D, None, None, Nchain=Nc, Niter=Niter, sampler_type="NUTS", dt=0.1, d_max=10, dtype="float32", seed=7, target=spec,
                      on_dmax="stop", kernel="tc")
    with contextlib.redirect_stdout(io.StringIO()):
        H.gen_sample(q0, verbose=False)
    out = (C.c_ulonglong * 8)()
    lib.hmc_debug_nuts_tc_cycles(out, 0)
    v = np.array(list(out), dtype=float)
    n = v[7]
    print("chains %d: %.1f ms, %.3g leapfrogs/s, %.0f passes per warp; cycles per pass: wait for the gradient %.0f | B consume + saves + check dots %.0f | barrier 1 %.0f | "
          "C state machine %.0f | barrier 2 %.0f | D second half + row moves + advance %.0f | E operand rows %.0f | total %.0f"
          % (Nc, H.kernel_ms, H.n_leapfrog_total / (H.kernel_ms * 1e-3), n / (16 * min(148, (Nc + 127) // 128)), v[0] / n, v[1] / n, v[2] / n, v[3] / n, v[4] / n, v[5] / n, v[6] / n, v[:7].sum() / n))
