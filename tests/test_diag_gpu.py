"""Diagnostics kernels (csrc/diag.cu) against the oracle restatement of utils.convergence_stats."""
import numpy as np
import pytest

from oracle import hmc_oracle as O
from _util import load

pytestmark = pytest.mark.gpu


def _ar1(Nchain, N, D, phi, seed, offset=0.0):
    rng = np.random.RandomState(seed)
    x = np.zeros((Nchain, N, D))
    x[:, 0] = rng.standard_normal((Nchain, D))
    for t in range(1, N):
        x[:, t] = phi * x[:, t - 1] + np.sqrt(1 - phi * phi) * rng.standard_normal((Nchain, D))
    return x + offset


@pytest.mark.parametrize("shape", [(4, 101, 3, 0.5), (10, 400, 7, 0.9), (3, 64, 100, 0.0), (2, 7, 1, 0.3),
                                   (16, 1000, 2, 0.99),
                                   # float32 with D % 4 == 0: the all-lags FFT pass (csrc/diag_fft.cu), straight away for
                                   # n >= 192, after one windowed chunk below that
                                   (10, 400, 8, 0.9), (6, 1000, 4, 0.99), (12, 200, 100, 0.8), (5, 1024, 12, 0.97), (40, 801, 100, 0.95),
                                   (4, 1100, 4, 0.9)])     # n = 550 > 512: beyond the transform length, back on the windowed kernels
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_convergence_stats_matches_oracle(shape, dtype):
    import utils as U
    Nchain, N, D, phi = shape
    x = _ar1(Nchain, N, D, phi, seed=N + D, offset=0.25).astype(dtype)
    for thin, warm in [(1, 0), (2, 3), (5, 0)]:
        if (N - warm) // thin < 6:
            continue
        R, ne = U.convergence_stats(x, thin_rate=thin, warm_up_num=warm)
        R0, ne0 = O.convergence_stats(x.astype(float), thin_rate=thin, warm_up_num=warm)
        np.testing.assert_allclose(R, R0, rtol=1e-9 if dtype == "float64" else 1e-6)
        # ESS: identical truncation lag unless a pair sum sits at zero within float32 noise
        np.testing.assert_allclose(ne, ne0, rtol=1e-8 if dtype == "float64" else 2e-4)


@pytest.mark.parametrize("name", ["random_case1a", "random_case3c_small", "nuts_d10"])
def test_convergence_stats_on_reference_chains(name):
    import utils as U
    fx = load(name)
    R, ne = U.convergence_stats(fx["q_chain"][:, 1:, :], thin_rate=1, warm_up_num=0)
    np.testing.assert_allclose(R, fx["R_q"], rtol=1e-9)
    np.testing.assert_allclose(ne, fx["n_eff_q"], rtol=1e-8)


@pytest.mark.parametrize("shape", [(37, 400, 100, 0.95, 0.0), (130, 125, 8, 0.6, 3.0), (9, 512, 4, 0.99, -40.0), (3, 33, 12, 0.5, 0.0)])
def test_all_lags_fft_equals_windowed_numerators(shape):
    """hmc_diag_variogram_all (one FFT per chain and dimension) against hmc_diag_variogram on the float64 copy of the same
    float32 stream, every lag 1..n-1, and against the plain numpy sum.  The float32 transform leaves an error relative to
    the series' energy (the edge sums of squares, of which a strongly correlated series' small-lag numerators are a small
    difference): bound 2e-6 of the largest numerator, measured 2e-7 to 3e-7."""
    import torch
    import hmc_b200_lib as L
    Nchain, n, D, phi, offset = shape
    lib = L.load()
    x32 = _ar1(Nchain, 2 * n, D, phi, seed=n + D, offset=offset).astype(np.float32)
    x = torch.from_numpy(x32).cuda()
    x64 = x.double()
    out = torch.zeros((n - 1, D), dtype=torch.float64, device="cuda")
    ws = torch.zeros((int(lib.hmc_diag_variogram_all_workspace_bytes(n, D)) // 8,), dtype=torch.float64, device="cuda")
    st = L.current_stream_ptr()
    L.check(lib.hmc_diag_variogram_all(L.HMC_F32, L.ptr(x), Nchain, n, D, 2 * n * D, n - 1, L.ptr(out), L.ptr(ws), ws.numel() * 8, st))
    ref = torch.zeros((n - 1, D), dtype=torch.float64, device="cuda")
    for lag0 in range(1, n, 32):
        nl = min(32, n - lag0)
        L.check(lib.hmc_diag_variogram(L.HMC_F64, L.ptr(x64), Nchain, n, D, 2 * n * D, lag0, nl, L.ptr(ref[lag0 - 1:]), st))
    torch.cuda.synchronize()
    got, want = out.cpu().numpy(), ref.cpu().numpy()
    xs = x32.astype(np.float64).reshape(Nchain * 2, n, D)
    for t in (1, 2, n // 2, n - 1):
        np.testing.assert_allclose(want[t - 1], np.sum((xs[:, t:] - xs[:, :-t]) ** 2, axis=(0, 1)), rtol=1e-12)
    scale = want.max(axis=0)
    assert np.max(np.abs(got - want) / scale) < 2e-6
    # a second call re-zeroes its workspace
    L.check(lib.hmc_diag_variogram_all(L.HMC_F32, L.ptr(x), Nchain, n, D, 2 * n * D, n - 1, L.ptr(out), L.ptr(ws), ws.numel() * 8, st))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(np.isfinite(out.cpu().numpy()), True)
    np.testing.assert_allclose(out.cpu().numpy(), got, rtol=1e-6, atol=1e-6 * scale.max())


def test_all_lags_fft_rejects_what_it_does_not_cover():
    import torch
    import hmc_b200_lib as L
    lib = L.load()
    x = torch.zeros((4, 80, 6), dtype=torch.float32, device="cuda")
    out = torch.zeros((39, 6), dtype=torch.float64, device="cuda")
    ws = torch.zeros((1 << 16,), dtype=torch.float64, device="cuda")
    rc = lib.hmc_diag_variogram_all(L.HMC_F32, L.ptr(x), 4, 40, 6, 480, 39, L.ptr(out), L.ptr(ws), ws.numel() * 8, L.current_stream_ptr())
    assert rc == L.HMC_E_UNSUPPORTED                                  # D % 4 != 0
    rc = lib.hmc_diag_variogram_all(L.HMC_F64, L.ptr(x), 4, 40, 8, 640, 39, L.ptr(out), L.ptr(ws), ws.numel() * 8, L.current_stream_ptr())
    assert rc == L.HMC_E_UNSUPPORTED                                  # float64 streams keep the windowed kernels


def test_variogram_single_lag():
    import utils as U
    x = _ar1(5, 50, 3, 0.7, seed=1)
    chains = [x[m] for m in range(5)]
    for var_num, lag in [(0, 1), (2, 7), (1, 49)]:
        assert U.variogram(chains, var_num, lag) == pytest.approx(O.variogram(chains, var_num, lag), rel=1e-12)


def test_large_stream_properties():
    """Size-independent checks at a BASELINE-scale shape slice: constant chains -> zero variogram and std;
    a chain-index offset changes B but not W."""
    import torch
    import utils as U
    Nchain, N, D = 2048, 200, 100
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((Nchain, N, D), device="cuda", dtype=torch.float32, generator=g)
    R, ne = U.convergence_stats(x, thin_rate=1, warm_up_num=0)
    assert np.all(np.abs(R - 1) < 0.01)
    assert np.all(ne == Nchain * N)            # white noise: rho_1 < 0.01 -> sum_rho = 0 (Q2 branch)
    xs = x + torch.arange(Nchain, device="cuda", dtype=torch.float32)[:, None, None] * 0.01
    R2, _ = U.convergence_stats(xs, thin_rate=1, warm_up_num=0)
    assert np.all(R2 > R)
