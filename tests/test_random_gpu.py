"""Parity of the CUDA random-trajectory sampler (through the C-ABI) with the reference / oracle.

Tolerances (BASELINE.json north_star: "leapfrog trajectories rel 1e-5 over the case-script L; identical
accept/reject decisions on the reference's draws"):
  * float64 kernel, reference draws injected, free running: whole sample stream within 1e-9 (relative to the
    largest |q|), energies within 1e-9 relative, identical acceptance counts and chain-0 decisions;
  * float32 kernels, teacher forced (every iteration restarted from the oracle's q_initial): trajectory end
    points within rel-L2 1e-5 for L < 20 (SURVEY H4 measured 2e-7..6e-6), decisions identical except
    near ties |ln u + dE| < 1e-5 * max(1, |E|) (float32 energy resolution), which are counted and bounded.
"""
import numpy as np
import pytest

from oracle import hmc_oracle as O
from _util import RANDOM_FIXTURES, load, flat_tape, sampler_from_fixture

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", RANDOM_FIXTURES)
def test_f64_free_running_matches_reference(name):
    fx = load(name)
    nsave = int(fx["N_save_chain0"])
    H = sampler_from_fixture(fx, dtype="float64", kernel="generic")
    H.gen_sample(fx["q_start"], N_save_chain0=nsave, verbose=False)
    qs = max(1.0, np.abs(fx["q_chain"]).max())
    es = max(1.0, np.abs(fx["E_chain"]).max())
    np.testing.assert_allclose(H.q_chain, fx["q_chain"], rtol=0, atol=1e-9 * qs)
    np.testing.assert_allclose(H.E_chain, fx["E_chain"], rtol=0, atol=1e-9 * es)
    np.testing.assert_allclose(H.dE_chain, fx["dE_chain"], rtol=0, atol=1e-9 * es)
    assert H.accept_R == pytest.approx(float(fx["accept_R"]), abs=1e-15)
    if not np.isnan(fx["accept_R_warm_up"]):
        assert H.accept_R_warm_up == pytest.approx(float(fx["accept_R_warm_up"]), abs=1e-15)
    assert H.N_total_steps == int(fx["N_total_steps"])
    if nsave > 0:
        np.testing.assert_array_equal(H.decision_chain, fx["decision_chain"])
        np.testing.assert_array_equal(np.array([p.shape[0] for p in H.phi_q]), fx["phi_len"])
        np.testing.assert_allclose(np.concatenate(H.phi_q, axis=0), fx["phi_q"], rtol=0, atol=1e-9 * qs)
    H.compute_convergence_stats()
    np.testing.assert_allclose(H.R_q, fx["R_q"], rtol=1e-7)
    np.testing.assert_allclose(H.n_eff_q, fx["n_eff_q"], rtol=1e-5)


def _teacher_forced(fx, dtype, kernel, always_accept):
    """One single-iteration chain per (chain, iteration) of the oracle run."""
    import samplers as S
    tgt = O.MVNTarget(fx["q0"], fx["cov0"])
    D, Nchain, Niter = int(fx["D"]), int(fx["Nchain"]), int(fx["Niter"])
    dt = fx["dt"] if fx["dt"].ndim else float(fx["dt"])
    R = O.gen_sample_random(D, tgt.V, tgt.dVdq, fx["q_start"], O.TapeDraws(flat_tape(fx)), Nchain, Niter,
                            int(fx["thin_rate"]), int(fx["warm_up_num"]), dt, int(fx["L_low"]), int(fx["L_high"]),
                            cov_p=fx["cov_p"], record=True)
    B = Nchain * Niter
    q_init = R.q_init.reshape(B, D)
    p_tape = np.zeros((B, 2, D))
    p_tape[:, 1] = R.p_tape[:, 1:].reshape(B, D)
    L_tape = R.L_tape.reshape(B, 1)
    u = R.u_tape.reshape(B, 1).copy()
    if always_accept:
        u[:] = 1e-300
    cov_p = None if np.array_equal(fx["cov_p"], np.eye(D)) else fx["cov_p"]
    H = S.HMC_sampler(D, None, None, Nchain=B, Niter=1, thin_rate=1, warm_up_num=0, cov_p=cov_p,
                      sampler_type="Random", dt=dt, L_low=int(fx["L_low"]), L_high=int(fx["L_high"]), dtype=dtype,
                      kernel=kernel, target=S.MVNSpec.from_cov(fx["q0"], fx["cov0"]),
                      draws=dict(p_tape=p_tape, L_tape=L_tape, u_tape=u))
    H.gen_sample(q_init, verbose=False)
    return R, H, B


FAST_OK = ("random_case3c_small", "random_case2c_small")     # the FFMA2 kernel covers 40 < D <= 100


@pytest.mark.parametrize("kernel", ["generic", "fast", "tc"])
@pytest.mark.parametrize("name", ["random_case1a", "random_d10_thin3", "random_case3c_small", "random_case2c_small"])
def test_f32_teacher_forced_trajectories(name, kernel):
    if kernel in ("fast", "tc") and name not in FAST_OK:
        pytest.skip("fused kernels: fast covers 20 < D <= 128, tc covers D = 100")
    fx = load(name)
    R, H, B = _teacher_forced(fx, "float32", kernel, always_accept=True)
    D = int(fx["D"])
    got = H.q_chain[:, 1, :]
    want = R.q_prop.reshape(B, D)
    # relative to the amplitude of the trajectory (an end point may pass arbitrarily close to the origin)
    amp = np.maximum(np.linalg.norm(want, axis=1), np.linalg.norm(R.q_init.reshape(B, D), axis=1))
    rel = np.linalg.norm(got - want, axis=1) / amp
    assert rel.max() < 1e-5, "worst rel-L2 trajectory error %.3g" % rel.max()
    # E_initial is stored at index 1 (warm_up_num = 0): computed from the float32-rounded start point, so its
    # error is |grad V| * 6e-8 |q| (cond(P) ~ 1900 at rho = 0.95): a few 1e-6 relative.
    E_want = R.E_init_iter.reshape(B)
    np.testing.assert_allclose(H.E_chain[:, 1, 0], E_want, rtol=3e-5, atol=3e-5)


@pytest.mark.parametrize("kernel", ["generic", "fast", "tc"])
@pytest.mark.parametrize("name", ["random_case1a", "random_d10_thin3", "random_case3c_small", "random_case2c_small"])
def test_f32_teacher_forced_decisions(name, kernel):
    if kernel in ("fast", "tc") and name not in FAST_OK:
        pytest.skip("fused kernels: fast covers 20 < D <= 128, tc covers D = 100")
    fx = load(name)
    R, H, B = _teacher_forced(fx, "float32", kernel, always_accept=False)
    D = int(fx["D"])
    want_dec = R.decision.reshape(B)
    want_q = np.where(want_dec[:, None] == 1, R.q_prop.reshape(B, D), R.q_init.reshape(B, D))
    got_q = H.q_chain[:, 1, :]
    qi = R.q_init.reshape(B, D)
    got_dec = np.linalg.norm(got_q - qi, axis=1) > 1e-5 * np.linalg.norm(qi, axis=1)       # moved <=> accepted
    dE = R.dE_iter.reshape(B)
    lnu = np.log(R.u_tape.reshape(B))
    margin = np.abs(lnu + dE)
    near_tie = (dE >= 0) & (margin < 1e-5 * np.maximum(1.0, np.abs(R.E_init_iter.reshape(B))))
    differ = got_dec != (want_dec == 1)
    assert not np.any(differ & ~near_tie), "decision flips away from ties: %d" % int(np.sum(differ & ~near_tie))
    assert near_tie.mean() < 0.02
    ok = ~differ
    amp = np.maximum(np.linalg.norm(want_q, axis=1), np.linalg.norm(qi, axis=1))
    rel = np.linalg.norm(got_q[ok] - want_q[ok], axis=1) / amp[ok]
    assert rel.max() < 1e-5


def test_f64_long_trajectories_case3d():
    """L in [50,200): float32 cannot hold rel 1e-5 here, the float64 instantiation must."""
    fx = load("random_case3d_small")
    R, H, B = _teacher_forced(fx, "float64", "generic", always_accept=True)
    want = R.q_prop.reshape(B, int(fx["D"]))
    rel = np.linalg.norm(H.q_chain[:, 1, :] - want, axis=1) / np.linalg.norm(want, axis=1)
    assert rel.max() < 1e-10


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_philox_path_matches_oracle_on_device_draws(dtype):
    """Free-running with the kernel's own Philox draws: the oracle is fed exactly those draws
    (hmc_philox_draws) and must produce the same stream (float64) / the same statistics (float32)."""
    import torch
    import hmc_b200_lib as L
    import samplers as S
    D, rho, Nchain, Niter, warm, thin = 10, 0.9, 64, 50, 10, 2
    tgt = O.MVNTarget(np.zeros(D), O.equicorrelated_cov(D, rho))
    rng = np.random.RandomState(3)
    q_start = rng.standard_normal((Nchain, D)) * 1.5
    seed, id0 = 1234, 1000
    H = S.HMC_sampler(D, tgt.V, tgt.dVdq, Nchain=Nchain, Niter=Niter, thin_rate=thin, warm_up_num=warm,
                      sampler_type="Random", dt=0.1, L_low=5, L_high=20, dtype=dtype, kernel="generic", seed=seed,
                      chain_id0=id0)
    H.gen_sample(q_start, verbose=False)
    lib = L.load()
    p = torch.zeros((Nchain, Niter + 1, D), dtype=torch.float64, device="cuda")
    Lt = torch.zeros((Nchain, Niter), dtype=torch.int32, device="cuda")
    u = torch.zeros((Nchain, Niter), dtype=torch.float64, device="cuda")
    L.check(lib.hmc_philox_draws(seed, id0, Nchain, Niter, D, 5, 20, L.ptr(p), L.ptr(Lt), L.ptr(u), L.current_stream_ptr()))
    torch.cuda.synchronize()
    fx = dict(Nchain=Nchain, Niter=Niter, p_tape=p.cpu().numpy(), L_tape=Lt.cpu().numpy(), u_tape=u.cpu().numpy())
    assert fx["L_tape"].min() >= 5 and fx["L_tape"].max() <= 19
    assert abs(fx["p_tape"].mean()) < 0.02 and abs(fx["p_tape"].std() - 1) < 0.02
    assert 0 < fx["u_tape"].min() and fx["u_tape"].max() < 1
    R = O.gen_sample_random(D, tgt.V, tgt.dVdq, q_start, O.TapeDraws(flat_tape(fx)), Nchain, Niter, thin, warm, 0.1, 5, 20)
    if dtype == "float64":
        np.testing.assert_allclose(H.q_chain, R.q_chain, rtol=0, atol=1e-9)
        np.testing.assert_allclose(H.E_chain, R.E_chain, rtol=0, atol=1e-8)
        assert H.accept_R == R.accept_R
    else:
        assert abs(H.accept_R - R.accept_R) < 0.01
        same = np.linalg.norm(H.q_chain[:, 1] - R.q_chain[:, 1], axis=1) / np.linalg.norm(R.q_chain[:, 1], axis=1)
        assert same.max() < 1e-4      # first stored sample: still on the same trajectory


def test_sharding_invariance_and_iteration_blocks():
    """Chains split over two 'devices' (chain_id0 offsets) and iterations split over several launches give
    bit-identical streams: Philox is keyed by global chain id and iteration."""
    import samplers as S
    D, Nchain, Niter = 6, 40, 30
    tgt = O.MVNTarget(np.zeros(D), O.equicorrelated_cov(D, 0.5))
    q_start = np.random.RandomState(0).standard_normal((Nchain, D))
    kw = dict(Niter=Niter, thin_rate=1, warm_up_num=5, sampler_type="Random", dt=0.2, L_low=3, L_high=9,
              dtype="float32", kernel="generic", seed=99, target=S.MVNSpec.from_cov(np.zeros(D), tgt.cov0))
    full = S.HMC_sampler(D, None, None, Nchain=Nchain, **kw)
    full.gen_sample(q_start, verbose=False)
    a = S.HMC_sampler(D, None, None, Nchain=16, chain_id0=0, iter_block=7, **kw)
    a.gen_sample(q_start[:16], verbose=False)
    b = S.HMC_sampler(D, None, None, Nchain=24, chain_id0=16, iter_block=11, **kw)
    b.gen_sample(q_start[16:], verbose=False)
    np.testing.assert_array_equal(np.concatenate([a.q_chain, b.q_chain]), full.q_chain)
    np.testing.assert_array_equal(np.concatenate([a.E_chain, b.E_chain]), full.E_chain)
    np.testing.assert_array_equal(np.concatenate([a.dE_chain, b.dE_chain]), full.dE_chain)


def test_fixed_sampler_matches_reference_constant_length_run():
    """sampler_type="Fixed" (a silent no-op upstream, samplers.py:378-383) runs the live random-length loop with a constant
    L.  Pinned to the REAL reference: the fixture random_constL is the reference's Random sampler with L_low = L, L_high =
    L + 1 (np.random.randint(7, 8) is always 7; tests/golden/make_golden.py) -- same draws, float64 kernel, free running:
    the whole sample stream, energies, acceptance, N_total_steps and the chain-0 record must be the reference's."""
    import samplers as S
    fx = load("random_constL")
    D, Lfix = int(fx["D"]), int(fx["L_low"])
    assert np.all(fx["L_tape"] == Lfix)
    nsave = int(fx["N_save_chain0"])
    H = S.HMC_sampler(D, None, None, Nchain=int(fx["Nchain"]), Niter=int(fx["Niter"]), thin_rate=int(fx["thin_rate"]),
                      warm_up_num=int(fx["warm_up_num"]), sampler_type="Fixed", L=Lfix, dt=float(fx["dt"]), dtype="float64",
                      kernel="generic", target=S.MVNSpec.from_cov(fx["q0"], fx["cov0"]),
                      draws=dict(p_tape=fx["p_tape"], L_tape=fx["L_tape"], u_tape=fx["u_tape"]))
    H.gen_sample(fx["q_start"], N_save_chain0=nsave, verbose=False)
    np.testing.assert_allclose(H.q_chain, fx["q_chain"], rtol=0, atol=1e-9 * max(1.0, np.abs(fx["q_chain"]).max()))
    np.testing.assert_allclose(H.E_chain, fx["E_chain"], rtol=0, atol=1e-9 * max(1.0, np.abs(fx["E_chain"]).max()))
    assert H.accept_R == pytest.approx(float(fx["accept_R"]), abs=1e-15)
    assert H.N_total_steps == int(fx["N_total_steps"])
    np.testing.assert_array_equal(H.decision_chain, fx["decision_chain"])
    # and with the kernels' own Philox draws: every trajectory has exactly L steps
    P = S.HMC_sampler(4, None, None, Nchain=8, Niter=20, sampler_type="Fixed", L=7, dt=0.1, dtype="float64", kernel="generic",
                      seed=5, target=S.MVNSpec.from_cov(np.zeros(4), np.eye(4)))
    P.gen_sample(np.random.RandomState(1).standard_normal((8, 4)), verbose=False)
    assert P.sum_L == 7 * 8 * 20
    assert P.N_total_steps == 8 * (1 + 2 * 20) + 4 * 49 * 8 * 20


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-12), ("float32", 2e-6)])
def test_leap_frog_primitive_matches_oracle(dtype, tol):
    """HMC_sampler.leap_frog (samplers.py:831-839) through the C-ABI (hmc_leap_frog): single pairs and batches, scalar and
    per-dimension dt, identity and dense momentum metric (force times M^-1, q moved by p: Q9), against the oracle's statement-by-
    statement restatement."""
    import samplers as S
    rng = np.random.RandomState(3)
    for D, dense_metric, vec_dt in [(2, False, False), (10, True, True), (100, False, False), (100, True, True), (300, False, True)]:
        cov = O.equicorrelated_cov(D, 0.6) if D != 300 else np.diag(rng.uniform(0.5, 2.0, D))
        tgt = O.MVNTarget(rng.standard_normal(D) * 0.3, cov)
        A = rng.standard_normal((D, D)) * 0.2
        cov_p = (np.eye(D) + A @ A.T) if dense_metric else None
        dt = (0.05 + 0.1 * rng.uniform(size=D)) if vec_dt else 0.1
        H = S.HMC_sampler(D, tgt.V, tgt.dVdq, Nchain=2, Niter=1, sampler_type="Random", dt=dt, global_dt=not vec_dt, L_low=1, L_high=2,
                          cov_p=cov_p, dtype=dtype)
        inv_cov_p = np.eye(D) if cov_p is None else np.linalg.inv(cov_p)
        p, q = rng.standard_normal((7, D)), rng.standard_normal((7, D)) * 1.5 + tgt.q0
        pn, qn = H.leap_frog(p, q)
        for b in range(7):
            pw, qw = O.leap_frog(p[b], q[b], dt, inv_cov_p, tgt.dVdq)
            scale = max(1.0, np.abs(pw).max(), np.abs(qw).max())
            assert np.abs(pn[b] - pw).max() < tol * scale and np.abs(qn[b] - qw).max() < tol * scale
        p1, q1 = H.leap_frog(p[0], q[0])                                   # a single pair keeps its shape
        assert p1.shape == q1.shape == (D,)
        np.testing.assert_array_equal(p1, pn[0]); np.testing.assert_array_equal(q1, qn[0])
        p3, q3 = H.leap_frog(p, q, nsteps=3)                               # repeated steps = three calls
        pr, qr = p, q
        for _ in range(3):
            pr, qr = H.leap_frog(pr, qr)
        np.testing.assert_allclose(p3, pr, rtol=0, atol=10 * tol); np.testing.assert_allclose(q3, qr, rtol=0, atol=10 * tol)


@pytest.mark.parametrize("kernel,D", [("tc", 100), ("fast", 100), ("generic", 12), ("bigd", 256)])
def test_store_ring_keeps_the_last_stored_samples(kernel, D):
    """hmc_random_args.store_ring = R (C-ABI only: streaming runs keep the last R stored samples per chain): stored index j lives in
    row j % R of q_chain / E_chain / dE_chain.  Same seed => the ring run's rows equal the full run's last R samples bit for bit."""
    import torch
    import hmc_b200_lib as L
    import samplers as S
    Nchain, Niter, warm, thin, R = 300, 31, 4, 2, 5
    spec = S.MVNSpec.from_cov(np.zeros(D), O.equicorrelated_cov(D, 0.6))
    q_start = (np.random.RandomState(D).standard_normal((Nchain, D)) * 1.2).astype(np.float32)
    kw = dict(Nchain=Nchain, Niter=Niter, thin_rate=thin, warm_up_num=warm, sampler_type="Random", dt=0.1, L_low=3, L_high=9,
              dtype="float32", seed=41, target=spec, kernel=kernel)
    F = S.HMC_sampler(D, None, None, **kw)
    F.gen_sample(q_start, verbose=False, quiet=True)
    Lc = F.L_chain
    assert Lc == 1 + (Niter - warm) // thin and Lc > R + 2
    H = S.HMC_sampler(D, None, None, **kw)
    run = H.prepare_random(q_start)
    a = run["args"]
    qr = torch.zeros((Nchain, R, D), dtype=torch.float32, device="cuda")
    Er = torch.zeros((Nchain, R), dtype=torch.float64, device="cuda")
    dEr = torch.zeros((Nchain, R), dtype=torch.float64, device="cuda")
    a.q_chain, a.E_chain, a.dE_chain, a.store_ring = qr.data_ptr(), Er.data_ptr(), dEr.data_ptr(), R
    a.iter_begin, a.iter_end = 0, Niter
    L.check(L.load().hmc_random_run(a, L.current_stream_ptr()))
    torch.cuda.synchronize()
    qr, Er, dEr = qr.cpu().numpy(), Er.cpu().numpy(), dEr.cpu().numpy()
    for j in range(Lc - R, Lc):
        np.testing.assert_array_equal(qr[:, j % R], F.q_chain[:, j].astype(np.float32))
        np.testing.assert_array_equal(Er[:, j % R], F.E_chain[:, j, 0])
        np.testing.assert_array_equal(dEr[:, j % R], F.dE_chain[:, j, 0])


def test_target_extraction_and_errors():
    import samplers as S
    import hmc_b200_lib as L
    D = 5
    rng = np.random.RandomState(2)
    A = rng.standard_normal((D, D))
    cov = A @ A.T + D * np.eye(D)
    q0 = rng.standard_normal(D)
    tgt = O.MVNTarget(q0, cov)
    spec = S.extract_mvn_target(D, tgt.V, tgt.dVdq)
    np.testing.assert_allclose(spec.P, tgt.inv_cov0, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(spec.mu, q0, rtol=1e-9, atol=1e-12)
    assert spec.const == pytest.approx(tgt.const, rel=1e-12)
    with pytest.raises(NotImplementedError):
        S.extract_mvn_target(D, lambda q: np.sum(q ** 4), lambda q: 4 * q ** 3)
    with pytest.raises(AssertionError):                                    # samplers.py:332
        S.HMC_sampler(D, tgt.V, tgt.dVdq, sampler_type="Random", L_low=2, L_high=5)
    H = S.HMC_sampler(D, tgt.V, tgt.dVdq, Nchain=3, Niter=5, sampler_type="Random", L_low=2, L_high=5, dt=0.1)
    with pytest.raises(AssertionError):                                    # samplers.py:396
        H.gen_sample(np.zeros((4, D)), verbose=False)
    bad = S.HMC_sampler(D, tgt.V, tgt.dVdq, Nchain=3, Niter=5, sampler_type="Random", L_low=5, L_high=5, dt=0.1)
    with pytest.raises(L.HMCError) as ei:
        bad.gen_sample(np.zeros((3, D)), verbose=False)
    assert ei.value.code == L.HMC_E_BADARG


@pytest.mark.parametrize("kernel", ["fast", "tc"])
def test_fast_kernel_slot_refill_matches_generic(kernel):
    """More chains than resident slots (148 SMs x 192): finished slots pull new chains from the queue, over
    several iteration-block launches.  Same Philox draws as the generic kernel => same streams up to float32
    summation order (first trajectory rel 1e-5, acceptance within 0.2 %)."""
    import samplers as S
    D, Nchain, Niter = 100, 40000, 4
    spec = S.MVNSpec.from_cov(np.zeros(D), O.equicorrelated_cov(D, 0.95))
    q_start = np.random.RandomState(5).standard_normal((Nchain, D)).astype(np.float32) * 1.4
    kw = dict(Nchain=Nchain, Niter=Niter, thin_rate=1, warm_up_num=0, sampler_type="Random", dt=0.1, L_low=5,
              L_high=20, dtype="float32", seed=11, target=spec)
    F = S.HMC_sampler(D, None, None, kernel=kernel, iter_block=3, **kw)
    F.gen_sample(q_start, verbose=False, quiet=True)
    G = S.HMC_sampler(D, None, None, kernel="generic", **kw)
    G.gen_sample(q_start, verbose=False, quiet=True)
    assert F.sum_L == G.sum_L
    qf, qg = F.q_chain, G.q_chain
    np.testing.assert_array_equal(qf[:, 0], qg[:, 0])
    amp = np.linalg.norm(q_start.astype(float), axis=1)
    rel = np.linalg.norm(qf[:, 1] - qg[:, 1], axis=1) / amp
    assert np.quantile(rel, 0.999) < 1e-5
    assert abs(F.accept_R - G.accept_R) < 2e-3
    assert np.all(np.isfinite(qf)) and np.abs(qf[:, -1]).max() < 50
    np.testing.assert_allclose(F.E_chain[:, :2, 0], G.E_chain[:, :2, 0], rtol=1e-5)


@pytest.mark.parametrize("nsb", [2, 5])
def test_tc_kernel_sub_blocks_match_generic(nsb, monkeypatch):
    """The tensor-core kernel's work queue hands out (chain, sub-block of the iteration block) units; a chain's units
    may run in different CTAs and pass position / E_prev through state_q / state_eprev.  Forced here with few chains
    (every later unit has to wait for its predecessor): same Philox draws as the generic kernel => same trajectory
    lengths, first trajectory to rel 1e-5, same chain-0 trajectory record, matching energies and acceptance."""
    import samplers as S
    monkeypatch.setenv("HMC_B200_TC_SUBBLOCKS", str(nsb))
    D, Nchain, Niter = 100, 700, 10
    spec = S.MVNSpec.from_cov(np.linspace(-1, 1, D), O.equicorrelated_cov(D, 0.95))
    q_start = np.random.RandomState(8).standard_normal((Nchain, D)).astype(np.float32) * 1.4
    kw = dict(Nchain=Nchain, Niter=Niter, thin_rate=1, warm_up_num=0, sampler_type="Random", dt=0.1, L_low=5,
              L_high=20, dtype="float32", seed=3, target=spec)
    F = S.HMC_sampler(D, None, None, kernel="tc", **kw)
    F.gen_sample(q_start, N_save_chain0=4, verbose=False, quiet=True)
    monkeypatch.delenv("HMC_B200_TC_SUBBLOCKS")
    G = S.HMC_sampler(D, None, None, kernel="generic", **kw)
    G.gen_sample(q_start, N_save_chain0=4, verbose=False, quiet=True)
    assert F.sum_L == G.sum_L
    np.testing.assert_array_equal(F.q_chain[:, 0], G.q_chain[:, 0])
    amp = np.linalg.norm(q_start.astype(float), axis=1)
    rel = np.linalg.norm(F.q_chain[:, 1] - G.q_chain[:, 1], axis=1) / amp
    assert np.quantile(rel, 0.99) < 1e-5
    assert np.all(np.isfinite(F.q_chain)) and np.abs(F.q_chain[:, -1]).max() < 50
    assert abs(F.accept_R - G.accept_R) < 2e-2
    np.testing.assert_allclose(F.E_chain[:, :2, 0], G.E_chain[:, :2, 0], rtol=1e-5)
    # every stored energy difference is E_init(it) - E_init(it - 1) of the same chain, also across unit boundaries
    np.testing.assert_allclose(F.dE_chain[:, 1:, 0], np.diff(F.E_chain[:, :, 0], axis=1), rtol=0, atol=2e-4)
    assert [len(x) for x in F.phi_q] == [len(x) for x in G.phi_q]
    np.testing.assert_allclose(F.phi_q[0], G.phi_q[0], rtol=0, atol=1e-4)


def test_tc_kernel_vector_dt_thinning_warmup_match_generic():
    """Tensor-core kernel off the benchmark path: per-dimension dt (the non-uniform-dt instantiation), thinning and
    warm-up (several iterations map to one stored index, the last write wins), a chain-id offset (multi-GPU shard) and
    iteration blocks.  Same Philox draws as the generic kernel => same trajectory lengths; with identical accept
    decisions the stored samples agree chain by chain, so the fraction of chains whose LAST stored sample agrees to
    1e-3 must be large (float32 streams separate only after a decision flips at a near-tie)."""
    import samplers as S
    D, Nchain, Niter = 100, 1500, 13
    mu = np.linspace(-2, 2, D)
    spec = S.MVNSpec.from_cov(mu, O.equicorrelated_cov(D, 0.9))
    q_start = (np.random.RandomState(21).standard_normal((Nchain, D)) * 1.3 + mu).astype(np.float32)
    dt = 0.06 + 0.08 * np.arange(D) / D
    kw = dict(Nchain=Nchain, Niter=Niter, thin_rate=3, warm_up_num=4, sampler_type="Random", dt=dt, L_low=3,
              L_high=12, dtype="float32", seed=9, target=spec, chain_id0=123456)
    F = S.HMC_sampler(D, None, None, kernel="tc", iter_block=5, **kw)
    F.gen_sample(q_start, verbose=False, quiet=True)
    G = S.HMC_sampler(D, None, None, kernel="generic", **kw)
    G.gen_sample(q_start, verbose=False, quiet=True)
    assert F.sum_L == G.sum_L
    assert F.q_chain.shape == G.q_chain.shape == (Nchain, 1 + (Niter - 4) // 3, D)
    assert abs(F.accept_R - G.accept_R) < 2e-2 and abs(F.accept_R_warm_up - G.accept_R_warm_up) < 2e-2
    amp = np.linalg.norm(G.q_chain[:, -1] - mu, axis=1)
    rel_last = np.linalg.norm(F.q_chain[:, -1] - G.q_chain[:, -1], axis=1) / amp
    assert np.mean(rel_last < 1e-3) > 0.9, "only %.3f of the chains agree at the last stored sample" % np.mean(rel_last < 1e-3)
    rel_first = np.linalg.norm(F.q_chain[:, 0] - G.q_chain[:, 0], axis=1) / amp
    assert np.quantile(rel_first, 0.95) < 1e-4
    same = rel_last < 1e-3
    np.testing.assert_allclose(F.E_chain[same, :, 0], G.E_chain[same, :, 0], rtol=2e-4, atol=2e-3)
    np.testing.assert_allclose(F.dE_chain[same, :, 0], G.dE_chain[same, :, 0], rtol=0, atol=5e-3)


def test_tc_kernel_very_short_trajectories_match_generic():
    """L in {1, 2}: the decision of a trajectory comes one or two passes after it started, before the momentum drawn
    ahead for the next iteration can be staged -- the chain has to wait for it (the `ready` / `need_take` path of the
    bookkeeping), also across chain refills (more chains than one CTA's 128 slots, few iterations)."""
    import samplers as S
    D, Nchain, Niter = 100, 700, 15
    spec = S.MVNSpec.from_cov(np.zeros(D), O.equicorrelated_cov(D, 0.95))
    q_start = (np.random.RandomState(31).standard_normal((Nchain, D)) * 1.4).astype(np.float32)
    kw = dict(Nchain=Nchain, Niter=Niter, thin_rate=1, warm_up_num=0, sampler_type="Random", dt=0.1, L_low=1,
              L_high=3, dtype="float32", seed=17, target=spec)
    F = S.HMC_sampler(D, None, None, kernel="tc", **kw)
    F.gen_sample(q_start, N_save_chain0=5, verbose=False, quiet=True)
    G = S.HMC_sampler(D, None, None, kernel="generic", **kw)
    G.gen_sample(q_start, N_save_chain0=5, verbose=False, quiet=True)
    assert F.sum_L == G.sum_L
    assert [len(x) for x in F.phi_q] == [len(x) for x in G.phi_q]
    amp = np.linalg.norm(q_start.astype(float), axis=1)
    rel = np.linalg.norm(F.q_chain[:, 1] - G.q_chain[:, 1], axis=1) / amp
    assert np.quantile(rel, 0.99) < 1e-5
    rel_last = np.linalg.norm(F.q_chain[:, -1] - G.q_chain[:, -1], axis=1) / amp
    assert np.mean(rel_last < 1e-3) > 0.9
    assert abs(F.accept_R - G.accept_R) < 2e-2
    np.testing.assert_allclose(F.E_chain[:, :2, 0], G.E_chain[:, :2, 0], rtol=1e-5)
    np.testing.assert_allclose(F.dE_chain[:, 1:, 0], np.diff(F.E_chain[:, :, 0], axis=1), rtol=0, atol=2e-4)


@pytest.mark.parametrize("D,rho,uniform_dt", [(96, 0.95, True), (64, 0.9, True), (52, 0.5, False), (8, 0.3, True)])
def test_tc_kernel_smaller_dimensions_match_generic(D, rho, uniform_dt):
    """Targets with D < 100 (a multiple of 4) run zero padded in the tensor-core kernel's 100-wide tile: same Philox draws as
    the generic kernel => same trajectory lengths, the same first trajectory (rel 1e-5), matching acceptance and energies,
    with warm-up, thinning, iteration blocks, a chain-id offset, slot refill (more chains than one CTA's slots) and the
    chain-0 trace.  "auto" picks the tensor-core kernel from D = 52 up."""
    import samplers as S
    Nchain, Niter = 2500, 9
    mu = np.linspace(-1, 1, D)
    spec = S.MVNSpec.from_cov(mu, O.equicorrelated_cov(D, rho))
    q_start = (np.random.RandomState(D).standard_normal((Nchain, D)) * 1.2 + mu).astype(np.float32)
    dt = 0.1 if uniform_dt else 0.05 + 0.1 * np.arange(D) / D
    kw = dict(Nchain=Nchain, Niter=Niter, thin_rate=2, warm_up_num=3, sampler_type="Random", dt=dt, L_low=4,
              L_high=11, dtype="float32", seed=5, target=spec)
    F = S.HMC_sampler(D, None, None, kernel="tc" if D < 52 else "auto", iter_block=4, **kw)
    F.gen_sample(q_start, N_save_chain0=3, verbose=False, quiet=True)
    G = S.HMC_sampler(D, None, None, kernel="generic", **kw)
    G.gen_sample(q_start, N_save_chain0=3, verbose=False, quiet=True)
    assert F.sum_L == G.sum_L
    assert F.q_chain.shape == G.q_chain.shape == (Nchain, 1 + (Niter - 3) // 2, D)
    amp = np.linalg.norm(q_start.astype(float) - mu, axis=1) + 1.0
    rel = np.linalg.norm(F.q_chain[:, 0] - G.q_chain[:, 0], axis=1) / amp          # index 0: overwritten at i = warm (3 iterations in)
    assert np.quantile(rel, 0.9) < 1e-4
    rel_last = np.linalg.norm(F.q_chain[:, -1] - G.q_chain[:, -1], axis=1) / amp
    assert np.mean(rel_last < 1e-3) > 0.9
    assert abs(F.accept_R - G.accept_R) < 1e-2
    same = rel_last < 1e-3
    np.testing.assert_allclose(F.E_chain[same, :, 0], G.E_chain[same, :, 0], rtol=2e-4, atol=2e-3)
    assert [len(x) for x in F.phi_q] == [len(x) for x in G.phi_q]
    np.testing.assert_allclose(F.phi_q[0], G.phi_q[0], rtol=0, atol=1e-4)
    assert np.all(np.isfinite(F.q_chain))
    # one iteration from the same start, no warm-up: the first trajectory itself
    kw1 = dict(kw, Niter=1, thin_rate=1, warm_up_num=0)
    F1 = S.HMC_sampler(D, None, None, kernel="tc", **kw1); F1.gen_sample(q_start, verbose=False, quiet=True)
    G1 = S.HMC_sampler(D, None, None, kernel="generic", **kw1); G1.gen_sample(q_start, verbose=False, quiet=True)
    rel1 = np.linalg.norm(F1.q_chain[:, 1] - G1.q_chain[:, 1], axis=1) / amp
    assert np.quantile(rel1, 0.995) < 1e-5
    np.testing.assert_allclose(F1.E_chain[:, 0, 0], G1.E_chain[:, 0, 0], rtol=2e-5)


@pytest.mark.parametrize("D,rho", [(128, 0.9), (101, 0.5), (64, 0.95), (33, 0.3), (24, 0.0)])
def test_fast_kernel_other_dimensions_match_generic(D, rho):
    """The fused kernel's other tile shapes (20 < D <= 128; padded dimensions, per-dimension dt when D is odd):
    same Philox draws as the generic kernel => same first trajectory (rel 1e-5), same trajectory lengths, matching
    acceptance rate."""
    import samplers as S
    Nchain, Niter = 3000, 6
    spec = S.MVNSpec.from_cov(np.linspace(-1, 1, D), O.equicorrelated_cov(D, rho))
    q_start = np.random.RandomState(D).standard_normal((Nchain, D)).astype(np.float32) * 1.2 + np.linspace(-1, 1, D).astype(np.float32)
    dt = 0.1 if D % 2 == 0 else 0.05 + 0.1 * np.arange(D) / D
    kw = dict(Nchain=Nchain, Niter=Niter, thin_rate=2, warm_up_num=2, sampler_type="Random", dt=dt, L_low=4,
              L_high=11, dtype="float32", seed=5, target=spec, chain_id0=77)
    F = S.HMC_sampler(D, None, None, kernel="fast", iter_block=4, **kw)
    F.gen_sample(q_start, verbose=False, quiet=True)
    G = S.HMC_sampler(D, None, None, kernel="generic", **kw)
    G.gen_sample(q_start, verbose=False, quiet=True)
    assert F.sum_L == G.sum_L
    amp = np.linalg.norm(q_start.astype(float) - spec.mu, axis=1) + 1.0
    rel = np.linalg.norm(F.q_chain[:, 1] - G.q_chain[:, 1], axis=1) / amp
    assert np.quantile(rel, 0.995) < 1e-5
    assert abs(F.accept_R - G.accept_R) < 5e-3
    np.testing.assert_allclose(F.E_chain[:, 0, 0], G.E_chain[:, 0, 0], rtol=2e-5)
    assert np.all(np.isfinite(F.q_chain))


def test_generic_kernel_large_dimension_matches_oracle():
    """D = 1024 (BASELINE config 5 shape): the generic kernel's widest instantiation, precision matrix read from
    HBM/L2.  Teacher-forced single iterations against the float64 oracle, float64 and float32."""
    import samplers as S
    D, B = 1024, 6
    rng = np.random.RandomState(7)
    lam = np.exp(rng.uniform(np.log(0.05), np.log(100.0), D))        # log-uniform spectrum
    Qm, _ = np.linalg.qr(rng.standard_normal((D, D)))
    cov = (Qm * lam) @ Qm.T
    cov = 0.5 * (cov + cov.T)
    tgt = O.MVNTarget(np.zeros(D), cov)
    q_init = rng.standard_normal((B, D)) * np.sqrt(lam).mean()
    p_tape = np.zeros((B, 2, D)); p_tape[:, 1] = rng.standard_normal((B, D))
    L_tape = rng.randint(20, 60, size=(B, 1)).astype(np.int32)
    u_tape = np.full((B, 1), 1e-300)
    want = np.zeros((B, D))
    for b in range(B):
        q, p = q_init[b].copy(), p_tape[b, 1].copy()
        for _ in range(int(L_tape[b, 0])):
            p, q = O.leap_frog(p, q, 0.1, np.eye(D), tgt.dVdq)
        want[b] = q
    spec = S.MVNSpec(np.zeros(D), tgt.inv_cov0, tgt.const)
    for dtype, tol in (("float64", 1e-9), ("float32", 2e-3)):
        H = S.HMC_sampler(D, None, None, Nchain=B, Niter=1, sampler_type="Random", dt=0.1, L_low=20, L_high=60,
                          dtype=dtype, kernel="generic", target=spec, draws=dict(p_tape=p_tape, L_tape=L_tape, u_tape=u_tape))
        H.gen_sample(q_init, verbose=False, quiet=True)
        rel = np.linalg.norm(H.q_chain[:, 1] - want, axis=1) / np.linalg.norm(q_init, axis=1)
        assert rel.max() < tol, (dtype, rel.max())
