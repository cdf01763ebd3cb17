"""Parity of the CUDA NUTS kernel (through the C-ABI) with the reference / oracle.

Tolerances: float64 kernel with the reference's own draws injected must reproduce the whole sample stream to
1e-9 (relative to the largest |q|) -- every direction coin, every progressive-sampling uniform and every
U-turn decision is then identical, which the leapfrog counts (N_total_steps) confirm exactly.  The float32
kernel is checked statistically (tree depths and sample moments) because one flipped U-turn decision changes
the rest of that chain."""
import numpy as np
import pytest

from oracle import hmc_oracle as O
from _util import NUTS_FIXTURES, load, sampler_from_fixture

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", NUTS_FIXTURES)
def test_f64_nuts_matches_reference(name):
    fx = load(name)
    H = sampler_from_fixture(fx, dtype="float64")
    H.gen_sample(fx["q_start"], verbose=False)
    qs = max(1.0, np.abs(fx["q_chain"]).max())
    es = max(1.0, np.abs(fx["E_chain"]).max())
    np.testing.assert_allclose(H.q_chain, fx["q_chain"], rtol=0, atol=1e-9 * qs)
    np.testing.assert_allclose(H.E_chain, fx["E_chain"], rtol=0, atol=1e-9 * es)
    np.testing.assert_allclose(H.dE_chain, fx["dE_chain"], rtol=0, atol=1e-9 * es)
    assert H.N_total_steps == int(fx["N_total_steps"])
    assert H.accept_R == 1.0
    H.compute_convergence_stats()
    np.testing.assert_allclose(H.R_q, fx["R_q"], rtol=1e-7)
    np.testing.assert_allclose(H.n_eff_q, fx["n_eff_q"], rtol=1e-5)


def test_f64_nuts_iteration_blocks_with_tapes():
    fx = load("nuts_d10")
    H = sampler_from_fixture(fx, dtype="float64", iter_block=7)
    H.gen_sample(fx["q_start"], verbose=False)
    np.testing.assert_allclose(H.q_chain, fx["q_chain"], rtol=0, atol=1e-9)
    assert H.N_total_steps == int(fx["N_total_steps"])


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_nuts_philox_statistics(dtype):
    """Own Philox draws, unit Gaussian D=10, dt=0.3: sample mean ~ 0, std ~ 1 within Monte-Carlo error."""
    import samplers as S
    D, Nchain, Niter, warm = 10, 512, 120, 20
    tgt = O.MVNTarget(np.zeros(D), np.eye(D))
    q_start = np.random.RandomState(0).standard_normal((Nchain, D)) * 1.5
    H = S.HMC_sampler(D, tgt.V, tgt.dVdq, Nchain=Nchain, Niter=Niter, warm_up_num=warm, sampler_type="NUTS", dt=0.3,
                      d_max=10, dtype=dtype, seed=3)
    H.gen_sample(q_start, verbose=False)
    x = H.q_chain[:, 1:, :]
    n = x.shape[0] * x.shape[1]
    assert np.all(np.abs(x.mean(axis=(0, 1))) < 5.0 / np.sqrt(n / 4))          # ESS >= n/4 for NUTS here
    assert np.all(np.abs(x.std(axis=(0, 1)) - 1.0) < 0.03)
    assert H.n_dmax == 0 and H.n_instability == 0
    lf = H.n_leapfrog_total / float(Nchain * Niter)
    assert 3 < lf < 40
    H.compute_convergence_stats()
    assert np.all(H.R_q < 1.01)


def test_nuts_dmax_assert_and_stop():
    """samplers.py:596-598: depth overflow aborts the run (AssertionError); on_dmax="stop" keeps the live point."""
    import samplers as S
    D, Nchain = 2, 8
    tgt = O.MVNTarget(np.zeros(D), np.eye(D))
    q_start = np.random.RandomState(1).standard_normal((Nchain, D))
    kw = dict(Nchain=Nchain, Niter=5, sampler_type="NUTS", dt=1e-3, d_max=3, dtype="float64", seed=1)
    H = S.HMC_sampler(D, tgt.V, tgt.dVdq, **kw)
    with pytest.raises(AssertionError):
        H.gen_sample(q_start, verbose=False)
    H2 = S.HMC_sampler(D, tgt.V, tgt.dVdq, on_dmax="stop", **kw)
    H2.gen_sample(q_start, verbose=False)
    assert H2.n_dmax == Nchain * 5 and np.all(H2.status == 1)
    assert H2.n_leapfrog_total == Nchain * 5 * 7          # 1 + 2 + 4 leapfrogs per iteration


def test_nuts_covariance_and_seed_dependence():
    """float64 NUTS on its own position-keyed Philox draws: sample covariance matches the target, two seeds give
    different streams."""
    import samplers as S
    D, Nchain, Niter = 5, 256, 60
    tgt = O.MVNTarget(np.zeros(D), O.equicorrelated_cov(D, 0.5))
    q_start = np.random.RandomState(2).standard_normal((Nchain, D))
    runs = []
    for seed in (1, 2):
        H = S.HMC_sampler(D, tgt.V, tgt.dVdq, Nchain=Nchain, Niter=Niter, warm_up_num=10, sampler_type="NUTS", dt=0.2,
                          d_max=10, dtype="float64", seed=seed)
        H.gen_sample(q_start, verbose=False)
        runs.append(H.q_chain[:, 1:, :])
    for x in runs:
        c = np.cov(x.reshape(-1, D).T)
        np.testing.assert_allclose(c, tgt.cov0, atol=0.08)
    assert np.abs(runs[0] - runs[1]).max() > 1e-3


# ---- chain-tiled tensor-core NUTS kernel (csrc/nuts_tc.cu): D = 100, float32 --------------------------------------------------
def _case3c(D=100):
    import samplers as S
    return S.MVNSpec.from_cov(np.zeros(D), O.equicorrelated_cov(D, 0.95))


def test_tc_nuts_teacher_forced_on_reference_tapes():
    """The reference's own direction coins and uniforms injected (fixture nuts_case3c_small: D = 100, rho = 0.95, dt = 0.2, run
    by the real reference): the float32 tensor-core kernel must take the same tree (leapfrog count per chain, N_total_steps)
    and land on the reference's samples to float32 accuracy."""
    fx = load("nuts_case3c_small")
    H = sampler_from_fixture(fx, dtype="float32", kernel="tc")
    H.gen_sample(fx["q_start"], verbose=False)
    assert H.nuts_kernel == "tc"
    assert H.N_total_steps == int(fx["N_total_steps"])
    qs = np.abs(fx["q_chain"]).max()
    np.testing.assert_allclose(H.q_chain, fx["q_chain"], rtol=0, atol=2e-4 * qs)
    np.testing.assert_allclose(H.E_chain, fx["E_chain"], rtol=2e-5, atol=2e-3)
    G = sampler_from_fixture(fx, dtype="float32", kernel="generic")
    G.gen_sample(fx["q_start"], verbose=False)
    np.testing.assert_array_equal(H.n_leapfrog, G.n_leapfrog)


def test_tc_nuts_matches_generic_on_philox_draws():
    """Same position-keyed Philox draws as the warp-per-chain kernel: per chain the same number of leapfrog steps (a float32
    near-tie in a U-turn test or a sampling uniform may change a tree: bounded), the same samples where the trees agree;
    iteration blocks (resume through state_q) reproduce the single-launch run bit for bit (zero mean: exact hand-off)."""
    import samplers as S
    D, Nchain, Niter = 100, 1500, 6
    spec = _case3c()
    q_start = (np.random.RandomState(3).standard_normal((Nchain, D)) * 1.4).astype(np.float32)
    kw = dict(Nchain=Nchain, Niter=Niter, thin_rate=1, warm_up_num=0, sampler_type="NUTS", dt=0.2, d_max=10, dtype="float32",
              seed=11, target=spec, on_dmax="stop", chain_id0=4242)
    T = S.HMC_sampler(D, None, None, kernel="tc", **kw)
    T.gen_sample(q_start, verbose=False)
    G = S.HMC_sampler(D, None, None, kernel="generic", **kw)
    G.gen_sample(q_start, verbose=False)
    same = T.n_leapfrog == G.n_leapfrog
    assert same.mean() > 0.97, same.mean()
    assert abs(T.n_leapfrog_total / float(G.n_leapfrog_total) - 1.0) < 0.01
    amp = np.linalg.norm(G.q_chain[:, -1], axis=1)
    rel = np.linalg.norm(T.q_chain[:, -1] - G.q_chain[:, -1], axis=1) / amp
    assert np.quantile(rel[same], 0.9) < 1e-3
    np.testing.assert_array_equal(T.q_chain[:, 0], G.q_chain[:, 0])
    np.testing.assert_allclose(T.E_chain[same][:, 1, 0], G.E_chain[same][:, 1, 0], rtol=2e-5, atol=2e-3)
    assert T.n_doublings > 0 and abs(T.n_doublings / float(G.n_doublings) - 1.0) < 0.01
    B = S.HMC_sampler(D, None, None, kernel="tc", iter_block=4, **kw)
    B.gen_sample(q_start, verbose=False)
    np.testing.assert_array_equal(B.q_chain, T.q_chain)
    np.testing.assert_array_equal(B.E_chain, T.E_chain)
    np.testing.assert_array_equal(B.n_leapfrog, T.n_leapfrog)
    T2 = S.HMC_sampler(D, None, None, kernel="tc", **kw)                       # determinism (hand-rolled synchronisation)
    T2.gen_sample(q_start, verbose=False)
    np.testing.assert_array_equal(T2.q_chain, T.q_chain)


def test_tc_nuts_statistics_and_slot_refill():
    """More chains than resident slots (148 x 128): finished slots take new chains from the queue.  Case-3c target, dt = 0.2:
    sample moments within Monte-Carlo error, Rhat ~ 1, d_max semantics (assert raises, stop counts)."""
    import samplers as S
    D, Nchain, Niter, warm = 100, 24000, 24, 8
    spec = _case3c()
    cov = O.equicorrelated_cov(D, 0.95)
    q_start = (np.random.RandomState(5).standard_normal((Nchain, D)) @ np.linalg.cholesky(cov).T).astype(np.float32)
    H = S.HMC_sampler(D, None, None, Nchain=Nchain, Niter=Niter, warm_up_num=warm, sampler_type="NUTS", dt=0.2, d_max=10,
                      dtype="float32", seed=9, target=spec, on_dmax="stop")
    H.gen_sample(q_start, verbose=False)
    assert H.nuts_kernel == "tc"
    s = H.sample_summary()
    n_ind = Nchain                                                            # independent chains bound the Monte-Carlo error
    assert np.abs(s["q_mean"]).max() < 5.0 / np.sqrt(n_ind)
    assert np.abs(np.sqrt(s["q_var"]) - 1.0).max() < 0.03
    assert H.n_instability == 0
    lf = H.n_leapfrog_total / float(Nchain * Niter)
    assert 20 < lf < 200
    H.compute_convergence_stats()
    assert np.all(np.isfinite(H.R_q)) and np.all(H.n_eff_q > 0)       # (16 strongly correlated samples per chain: Rhat ~ 1.3 by construction)
    A = S.HMC_sampler(D, None, None, Nchain=256, Niter=3, sampler_type="NUTS", dt=1e-3, d_max=3, dtype="float32", seed=1,
                      target=spec, kernel="tc")
    with pytest.raises(AssertionError):
        A.gen_sample(q_start[:256], verbose=False)
    A2 = S.HMC_sampler(D, None, None, Nchain=256, Niter=3, sampler_type="NUTS", dt=1e-3, d_max=3, dtype="float32", seed=1,
                       target=spec, kernel="tc", on_dmax="stop")
    A2.gen_sample(q_start[:256], verbose=False)
    assert A2.n_dmax == 256 * 3 and np.all(A2.status == 1) and A2.n_leapfrog_total == 256 * 3 * 7
