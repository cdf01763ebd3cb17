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
                                   (16, 1000, 2, 0.99)])
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
