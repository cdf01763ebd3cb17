"""Device start points (utils.start_pts, /root/reference/utils.py:204-209) and the GPU reductions behind
sampler.plot_samples (samplers.py:84-113, 160-186, 209-250), checked against numpy on the host copies of a small run."""
import numpy as np
import pytest

from oracle import hmc_oracle as O

pytestmark = pytest.mark.gpu


def _small_run(dtype="float32", D=10, Nchain=200, Niter=61):
    import samplers as S
    cov = O.equicorrelated_cov(D, 0.6)
    spec = S.MVNSpec.from_cov(np.linspace(-1, 2, D), cov)
    rng = np.random.RandomState(1)
    q_start = rng.standard_normal((Nchain, D)) * 1.3 + spec.mu
    H = S.HMC_sampler(D, None, None, Nchain=Nchain, Niter=Niter, thin_rate=1, warm_up_num=10, sampler_type="Random",
                      dt=0.15, L_low=3, L_high=9, dtype=dtype, kernel="generic", seed=3, target=spec)
    H.gen_sample(q_start, verbose=False, quiet=True)
    return H, spec, cov


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_sample_summary_matches_numpy(dtype):
    H, spec, cov = _small_run(dtype)
    s = H.sample_summary()
    q, E, dE = H.q_chain, H.E_chain, H.dE_chain
    # samplers.py:209-216, 244-250
    np.testing.assert_allclose(s["q_mean"], q[:, 1:, :].mean(axis=(0, 1)), rtol=0, atol=1e-12 if dtype == "float64" else 1e-6)
    np.testing.assert_allclose(s["q_var"], q[:, 1:, :].std(axis=(0, 1)) ** 2, rtol=1e-10 if dtype == "float64" else 1e-5)
    for name, col in (("q1", 0), ("q2", 1)):                            # samplers.py:95-113, 160-175
        x = q[:, :, col].flatten()
        hi, lo = np.percentile(x, 97.5), np.percentile(x, 2.5)
        r, c = (hi - lo) * 2.5, (hi + lo) / 2.
        np.testing.assert_allclose(s[name + "_range"], (c - r / 2., c + r / 2.), rtol=1e-12, atol=1e-12)
        edges = np.arange(c - r / 2., c + r / 2., r / 100.)
        np.testing.assert_allclose(s[name + "_edges"], edges, rtol=1e-12, atol=1e-12)
        np.testing.assert_array_equal(s[name + "_hist"], np.histogram(x, bins=s[name + "_edges"])[0])
    Ec = E[:, 1:, :].flatten()
    Ec = Ec - np.mean(Ec)                                               # samplers.py:86-87
    assert s["E_mean"] == pytest.approx(np.mean(E[:, 1:, :]), rel=1e-12)
    lo, hi = np.percentile(Ec, 2.5), np.percentile(Ec, 97.5)            # samplers.py:177-183
    r, c = (hi - lo) * 2.5, (lo + hi) / 2.
    np.testing.assert_allclose(s["E_range"], (c - r / 2., c + r / 2.), rtol=1e-9, atol=1e-9)
    # bin counts on the summary's own edges: values within 1e-9 of an edge may fall either side of it
    for key, x in (("E_hist", E[:, 1:, :].flatten() - s["E_mean"]), ("dE_hist", dE[:, 1:, :].flatten())):
        want = np.histogram(x, bins=s["E_edges"])[0]
        assert np.abs(s[key] - want).sum() <= 2, (key, np.abs(s[key] - want).sum())
        assert s[key].sum() > 0


def test_device_percentile_and_histogram_edge_cases():
    import torch
    import utils as U
    rng = np.random.RandomState(0)
    for dt in (np.float32, np.float64):
        x = rng.standard_normal((37, 53)).astype(dt)
        x[3, 5] = -0.0
        x[4, 4] = 0.0
        x[:, 7] = 1.25                                                  # ties
        xt = torch.from_numpy(x).cuda()
        for view, ref in ((xt, x), (xt[:, 1::3], x[:, 1::3])):          # contiguous and strided series
            got = U.device_percentile(view, [0.0, 2.5, 50.0, 97.5, 100.0])
            np.testing.assert_allclose(got, np.percentile(ref.astype(np.float64).flatten(), [0.0, 2.5, 50.0, 97.5, 100.0]), rtol=1e-14, atol=0)
            edges = np.linspace(-2.0, 2.0, 41)
            cnt, below, above = U.device_histogram(view, edges)
            np.testing.assert_array_equal(cnt, np.histogram(ref.astype(np.float64).flatten(), bins=edges)[0])
            assert below == int((ref < -2.0).sum()) and above == int((ref > 2.0).sum())


def test_device_start_pts_distribution_and_sharding():
    import utils as U
    D, N = 12, 200000
    rng = np.random.RandomState(2)
    A = rng.standard_normal((D, D))
    cov = A @ A.T + D * np.eye(D)
    q0 = np.linspace(-3, 3, D)
    x = U.start_pts(q0, cov, N, device="cuda", seed=11).double().cpu().numpy()
    assert x.shape == (N, D)
    se = np.sqrt(np.diag(cov) / N)
    assert np.all(np.abs(x.mean(axis=0) - q0) < 5 * se)
    emp = np.cov(x.T)
    assert np.abs(emp - cov).max() < 6 * np.abs(cov).max() * np.sqrt(2.0 / N)
    # keyed by global chain id: a shard is the corresponding slice of the whole
    part = U.start_pts(q0, cov, 1000, device="cuda", seed=11, chain_id0=5000).double().cpu().numpy()
    np.testing.assert_array_equal(part, x[5000:6000])
    # diagonal covariance (case scripts: cov_start = 2 I, case3-script.py:57) and float64 output
    y = U.start_pts(np.zeros(D), 2.0 * np.eye(D), N, device="cuda", seed=3, dtype="float64").cpu().numpy()
    assert abs(y.std() - np.sqrt(2.0)) < 0.01 and abs(y.mean()) < 0.01
    # default path: the reference's own host call on the global numpy stream
    np.random.seed(4)
    a = U.start_pts(q0, cov, 5)
    np.random.seed(4)
    np.testing.assert_array_equal(a, np.random.multivariate_normal(q0, cov, size=5))


def test_diag_large_dimension_and_far_offset():
    """Diagnostics for D > 256 (dimension tiles) and chains far from zero (shifted between-chain sums)."""
    import torch
    import utils as U
    rng = np.random.RandomState(5)
    for (Nchain, N, D, off, dt) in ((6, 41, 600, 0.0, np.float32), (5, 80, 1024, 0.0, np.float64), (12, 90, 6, 3.0e4, np.float32),
                                    (8, 64, 7, 1.0e6, np.float64)):
        x = np.zeros((Nchain, N, D))
        for t in range(1, N):
            x[:, t] = 0.7 * x[:, t - 1] + rng.standard_normal((Nchain, D))
        x += rng.standard_normal((Nchain, 1, D)) * 0.5 + off
        x = x.astype(dt)
        R, ne = U.convergence_stats(torch.from_numpy(x).cuda(), thin_rate=1, warm_up_num=0)
        R0, ne0 = O.convergence_stats_fast(x.astype(np.float64) - off, 1, 0)
        tol = 1e-8 if dt == np.float64 else (2e-3 if off else 2e-5)
        np.testing.assert_allclose(R, R0, rtol=tol)
        np.testing.assert_allclose(ne, ne0, rtol=100 * tol if dt == np.float32 else 1e-6)
