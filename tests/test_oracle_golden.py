"""The oracle restatement against the committed reference-generated fixtures (runs anywhere, no GPU)."""
import numpy as np
import pytest

from oracle import hmc_oracle as O
from _util import RANDOM_FIXTURES, NUTS_FIXTURES, load, flat_tape, ChainTapeDraws


def _target(fx):
    return O.MVNTarget(fx["q0"], fx["cov0"])


@pytest.mark.parametrize("name", RANDOM_FIXTURES)
def test_random_matches_reference(name):
    fx = load(name)
    tgt = _target(fx)
    D, Nchain, Niter = int(fx["D"]), int(fx["Nchain"]), int(fx["Niter"])
    draws = O.TapeDraws(flat_tape(fx))
    dt = fx["dt"] if fx["dt"].ndim else float(fx["dt"])
    R = O.gen_sample_random(D, tgt.V, tgt.dVdq, fx["q_start"], draws, Nchain, Niter, int(fx["thin_rate"]),
                            int(fx["warm_up_num"]), dt, int(fx["L_low"]), int(fx["L_high"]), cov_p=fx["cov_p"],
                            N_save_chain0=int(fx["N_save_chain0"]), record=True)
    assert draws.exhausted()
    scale = max(1.0, np.abs(fx["E_chain"]).max())
    np.testing.assert_allclose(R.q_chain, fx["q_chain"], rtol=0, atol=1e-12 * max(1.0, np.abs(fx["q_chain"]).max()))
    np.testing.assert_allclose(R.E_chain, fx["E_chain"], rtol=0, atol=1e-12 * scale)
    np.testing.assert_allclose(R.dE_chain, fx["dE_chain"], rtol=0, atol=1e-11 * scale)
    assert R.accept_R == pytest.approx(float(fx["accept_R"]), abs=0)
    if not np.isnan(fx["accept_R_warm_up"]):
        assert R.accept_R_warm_up == pytest.approx(float(fx["accept_R_warm_up"]), abs=0)
    assert R.N_total_steps == int(fx["N_total_steps"])
    assert O.n_total_steps_random(Nchain, Niter, D, fx["L_tape"]) == int(fx["N_total_steps"])
    if int(fx["N_save_chain0"]) > 0:
        np.testing.assert_array_equal(R.decision_chain, fx["decision_chain"])
        np.testing.assert_allclose(np.concatenate(R.phi_q, axis=0), fx["phi_q"], rtol=0, atol=1e-12)
    Rq, neff = O.convergence_stats(R.q_chain[:, 1:, :], 1, 0)
    np.testing.assert_allclose(Rq, fx["R_q"], rtol=1e-10)
    np.testing.assert_allclose(neff, fx["n_eff_q"], rtol=1e-9)
    Rf, nf = O.convergence_stats_fast(fx["q_chain"][:, 1:, :], 1, 0)
    np.testing.assert_allclose(Rf, fx["R_q"], rtol=1e-12)
    np.testing.assert_allclose(nf, fx["n_eff_q"], rtol=1e-11)


@pytest.mark.parametrize("name", NUTS_FIXTURES)
def test_nuts_matches_reference(name):
    fx = load(name)
    tgt = _target(fx)
    D, Nchain, Niter = int(fx["D"]), int(fx["Nchain"]), int(fx["Niter"])
    draws = ChainTapeDraws(fx["p_tape"], fx["dir_tape"], fx["u_tape"])
    R = O.gen_sample_NUTS(D, tgt.V, tgt.dVdq, fx["q_start"], draws, Nchain, Niter, int(fx["thin_rate"]),
                          int(fx["warm_up_num"]), float(fx["dt"]), int(fx["d_max"]), cov_p=fx["cov_p"], record=True)
    scale = max(1.0, np.abs(fx["E_chain"]).max())
    np.testing.assert_allclose(R.q_chain, fx["q_chain"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(R.E_chain, fx["E_chain"], rtol=0, atol=1e-12 * scale)
    np.testing.assert_allclose(R.dE_chain, fx["dE_chain"], rtol=0, atol=1e-11 * scale)
    assert R.N_total_steps == int(fx["N_total_steps"])
    assert O.n_total_steps_nuts(Nchain, Niter, D, R.n_leapfrog.sum()) == int(fx["N_total_steps"])
    for m in range(Nchain):            # every recorded draw consumed, none missing
        assert len(R.dir_tape[m]) == int(fx["n_dir"][m])
        assert len(R.u_tape[m]) == int(fx["n_u"][m])
    Rq, neff = O.convergence_stats(R.q_chain[:, 1:, :], 1, 0)
    np.testing.assert_allclose(Rq, fx["R_q"], rtol=1e-10)
    np.testing.assert_allclose(neff, fx["n_eff_q"], rtol=1e-9)


def test_index_helpers_known_answers():
    """README:332-358 tables (regenerated from the reference's own functions) and the closed forms."""
    kat = load("nuts_index_kat")
    off = 0
    for m, n in zip(kat["cp_m"], kat["cp_len"]):
        want = kat["cp_flat"][off:off + n].tolist()
        off += n
        assert O.check_points(int(m)).tolist() == want
        assert O.check_points_closed(int(m)) == want
    for m, l, r, rf in kat["release"]:
        assert bool(r) == bool(rf)
        assert O.release(int(m), int(l)) == bool(r)
        assert O.release_closed(int(m), int(l)) == bool(r)
    # README:332-347 spot values
    assert O.check_points(8).tolist() == [1, 5, 7]
    assert O.check_points(16).tolist() == [1, 9, 13, 15]
    assert O.check_points(24).tolist() == [17, 21, 23]
    for (m, l) in [(4, 3), (8, 5), (8, 7), (12, 11), (20, 19), (24, 21), (24, 23), (28, 27)]:   # README:349-358
        assert O.release(m, l)
    for (m, l) in [(10, 9), (14, 13)]:
        assert not O.release(m, l)


def test_slot_closed_form_is_collision_free():
    """Replays utils.test_NUTS_binary_tree_flatten's bookkeeping (utils.py:387-423) for d=10 and checks that
    popcount((l-1)>>1) never maps two simultaneously-live points to one slot and stays below d."""
    d = 10
    table = np.ones(d + 1, dtype=int) * -1
    table[0] = 1
    for m in range(2, 2 ** d + 1):
        if m % 2 == 1:
            table[O.find_next(table)] = m
            live = [int(x) for x in table if x > 0]
            slots = [O.slot_closed(l) for l in live]
            assert len(set(slots)) == len(slots)
            assert max(slots) < d
        else:
            for l in O.check_points(m):
                idx = O.retrieve_save_index(table, l)
                assert idx is not None
                if l > 1 and O.release(m, l):
                    table[idx] = -1


@pytest.mark.parametrize("name", ["random_case3c_small", "random_case2c_small", "random_vecdt_covp", "random_case1a"])
def test_batched_iteration_matches_reference(name):
    """``one_iteration_batch`` (all chains of an iteration at once; what the at-scale GPU parity tests use) reproduces the
    reference's per-iteration records: proposals, energy differences and accept decisions of the fixture's own draws."""
    fx = load(name)
    tgt = _target(fx)
    D, Nchain, Niter = int(fx["D"]), int(fx["Nchain"]), int(fx["Niter"])
    dt = fx["dt"] if fx["dt"].ndim else float(fx["dt"])
    R = O.gen_sample_random(D, tgt.V, tgt.dVdq, fx["q_start"], O.TapeDraws(flat_tape(fx)), Nchain, Niter,
                            int(fx["thin_rate"]), int(fx["warm_up_num"]), dt, int(fx["L_low"]), int(fx["L_high"]),
                            cov_p=fx["cov_p"], record=True)
    B = Nchain * Niter
    out = O.one_iteration_batch(tgt, R.q_init.reshape(B, D), R.p_tape[:, 1:].reshape(B, D), R.L_tape.reshape(B),
                                R.u_tape.reshape(B), dt, cov_p=fx["cov_p"])
    qs = max(1.0, np.abs(R.q_prop).max())
    np.testing.assert_allclose(out["q_prop"], R.q_prop.reshape(B, D), rtol=0, atol=1e-11 * qs)
    es = max(1.0, np.abs(R.E_init_iter).max())
    np.testing.assert_allclose(out["E_init"], R.E_init_iter.reshape(B), rtol=0, atol=1e-11 * es)
    np.testing.assert_allclose(out["dE"], R.dE_iter.reshape(B), rtol=0, atol=1e-9 * es)
    clear = np.abs(np.log(R.u_tape.reshape(B)) + R.dE_iter.reshape(B)) > 1e-8 * es
    np.testing.assert_array_equal(out["decision"][clear], R.decision.reshape(B)[clear])
