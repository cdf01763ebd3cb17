"""Small driver for ncu: the diagnostics kernels on an AR(1) stream (Nchain x N x 100 float32), long series (windowed
variogram + moments) and short series (single-pass kernel).  Usage: run_diag_once.py [Nchain] [N]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "understanding-hmc_b200"))
import numpy as np, torch
import hmc_b200_lib as L
lib = L.load()
Nc, N, D = int(sys.argv[1]) if len(sys.argv) > 1 else 16384, int(sys.argv[2]) if len(sys.argv) > 2 else 1000, 100
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.empty((Nc, N, D), dtype=torch.float32, device="cuda")
x[:, 0] = torch.randn((Nc, D), device="cuda", generator=g)
for t in range(1, N):
    x[:, t] = 0.9 * x[:, t - 1] + 0.4359 * torch.randn((Nc, D), device="cuda", generator=g)
mom = torch.empty((4, D), dtype=torch.float64, device="cuda")
buf = torch.empty((32, D), dtype=torch.float64, device="cuda")
st = L.current_stream_ptr()
def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
n = N // 2
gb = Nc * N * D * 4 / 1e9
t_m = timed(lambda: L.check(lib.hmc_diag_moments(L.HMC_F32, L.ptr(x), Nc, n, D, x.stride(0), L.ptr(mom), st)))
t_v = timed(lambda: L.check(lib.hmc_diag_variogram(L.HMC_F32, L.ptr(x), Nc, n, D, x.stride(0), 1, 32, L.ptr(buf), st)))
print("long series n=%d: %.2f GB  moments %.3f ms = %.0f GB/s | variogram (32 lags) %.3f ms = %.0f GB/s algorithmic" % (n, gb, t_m, gb / t_m * 1e3, t_v, gb / t_v * 1e3))
if n <= 512:
    abuf = torch.empty((n - 1, D), dtype=torch.float64, device="cuda")
    ws = torch.empty((int(lib.hmc_diag_variogram_all_workspace_bytes(n, D)) // 8,), dtype=torch.float64, device="cuda")
    t_f = timed(lambda: L.check(lib.hmc_diag_variogram_all(L.HMC_F32, L.ptr(x), Nc, n, D, x.stride(0), n - 1, L.ptr(abuf), L.ptr(ws), ws.numel() * 8, st)))
    print("all lags (FFT) n=%d: %.3f ms = %.0f GB/s algorithmic, %.1f ns per (chain, dimension) transform" % (n, t_f, gb / t_f * 1e3, t_f * 1e6 / (Nc * D)))
ns = 25                                  # the bench's short series: 50 stored samples per chain
xs = x[:, :2 * ns]
gbs = Nc * 2 * ns * D * 4 / 1e9
t_s = timed(lambda: L.check(lib.hmc_diag_short_series(L.HMC_F32, L.ptr(xs), Nc, ns, D, xs.stride(0), ns - 1, L.ptr(mom), L.ptr(buf), st)))
print("short series n=%d: %.3f GB  single pass (moments + 24 lags) %.3f ms = %.0f GB/s" % (ns, gbs, t_s, gbs / t_s * 1e3))
