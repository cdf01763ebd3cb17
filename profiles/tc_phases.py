"""Per-phase cycle accounting of the tensor-core trajectory kernel (make -C understanding-hmc_b200/csrc prof: clock() around the
phases, -DHMC_PROFILE_PHASES).  Usage: python profiles/tc_phases.py [chains] [iterations]"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "understanding-hmc_b200"))
import numpy as np, torch
import hmc_b200_lib as L
L.LIB_PATH = os.path.join(ROOT, "understanding-hmc_b200", "csrc", "_prof", "libhmc_b200_prof.so")
import samplers as S
D, Nc, IB = 100, int(sys.argv[1]) if len(sys.argv) > 1 else 65536, int(sys.argv[2]) if len(sys.argv) > 2 else 100
spec = S.MVNSpec.from_cov(np.zeros(D), S.equicorrelated_cov(D, 0.95))
q0 = (np.random.RandomState(0).standard_normal((Nc, D)) * 1.4).astype(np.float32)
lib = L.load()
lib.hmc_debug_tc_cycles.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
H = S.HMC_sampler(D, None, None, Nchain=Nc, Niter=IB * 2, sampler_type="Random", dt=0.1, L_low=5, L_high=20,
                  dtype="float32", kernel="tc", seed=1, target=spec)
run = H.prepare_random(q0)
run["args"].iter_begin, run["args"].iter_end = 0, IB
L.check(lib.hmc_random_run(run["args"], L.current_stream_ptr())); torch.cuda.synchronize()
lib.hmc_debug_tc_cycles(None, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
run["args"].iter_begin, run["args"].iter_end = IB, 2 * IB
L.check(lib.hmc_random_run(run["args"], L.current_stream_ptr()))
e1.record(); torch.cuda.synchronize()
out = (C.c_ulonglong * 16)()
lib.hmc_debug_tc_cycles(out, 0)
v = np.array(list(out), dtype=float)
print("tensor-core kernel, %d chains x %d iterations: %.2f ms" % (Nc, IB, e0.elapsed_time(e1)))
for name, o, n in (("bookkeeping warps (slice 0)", 0, 4), ("other worker warps", 8, 12)):
    steps = v[o + 5]
    print("  %s: passes/warp %.0f; cycles per pass: apply commands %.0f | wait for MMA %.0f | TMEM read + update + re-split + TMEM write %.0f | "
          "S1 arrive + group barrier A %.0f | %s %.0f | group barrier B %.0f" % (name, steps / (148 * n), v[o + 7] / steps, v[o + 0] / steps, v[o + 1] / steps, v[o + 2] / steps,
                                                         "bookkeeping (P2)" if o == 0 else "momentum draws", v[o + 3] / steps, v[o + 6] / steps))
try:
    lib.hmc_debug_tc_apply.argtypes = [C.POINTER(C.c_ulonglong)]
    ap = (C.c_ulonglong * 8)(); lib.hmc_debug_tc_apply(ap)
    nw = v[5] + v[13]
    print("  inside apply (all worker warps, cycles per pass): command words + new / parked chains %.0f | start-point rows (accepted) %.0f | restore / take loads %.0f | warp sync %.0f"
          % (ap[0] / nw, ap[1] / nw, ap[2] / nw, ap[3] / nw))
except Exception as exc:
    print("  (no apply breakdown: %r)" % exc)
print("  issuing warp: %.0f cycles per pass inside the MMA issue" % (v[4] / (v[5] / 4)))
