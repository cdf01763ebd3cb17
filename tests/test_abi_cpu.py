"""No-GPU checks: the C-ABI library loads and exports every symbol include/hmc_b200.h declares; ctypes mirrors
have the C struct sizes; host-side finishing logic of the diagnostics equals the oracle's."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import hmc_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import hmc_b200_lib as L
    if not os.path.isfile(L.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    hdr = open(os.path.join(ROOT, "include", "hmc_b200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|const char\*)\s+(hmc_\w+)\s*\(", hdr, flags=re.M))
    assert declared == set(L.EXPORTS)
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert L.load().hmc_version() == 100


def test_ctypes_structs_match_c_layout(tmp_path):
    import hmc_b200_lib as L
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "hmc_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(hmc_target),sizeof(hmc_random_args),sizeof(hmc_nuts_args),'
                   'offsetof(hmc_random_args,store_ring),offsetof(hmc_nuts_args,n_leapfrog));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert [int(v) for v in out] == [ctypes.sizeof(L.Target), ctypes.sizeof(L.RandomArgs), ctypes.sizeof(L.NutsArgs),
                                     L.RandomArgs.store_ring.offset, L.NutsArgs.n_leapfrog.offset]


def test_missing_library_fails_loudly(monkeypatch):
    import hmc_b200_lib as L
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", "/nonexistent/libhmc_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L.load()


def test_host_finishing_rule_equals_oracle():
    """utils._finish_n_eff (fed lag chunks, as the GPU path does) == the oracle's sequential rule."""
    import utils as U
    rng = np.random.RandomState(0)
    for trial in range(20):
        Nchain, N, D = 4, 60 + trial, 5
        phi = rng.uniform(-0.5, 0.98)
        x = np.zeros((Nchain, N, D))
        for t in range(1, N):
            x[:, t] = phi * x[:, t - 1] + rng.standard_normal((Nchain, D))
        n, m, std_j, mean_j, halves = O.rhat_moments(x, 1, 0)
        W = std_j.mean(0)
        B = np.sum((mean_j - mean_j.mean(0)) ** 2, 0) * n / float(m - 1)
        var = W * (n - 1) / float(n) + B / float(n)
        V = np.zeros((n, D))
        for t in range(1, n):
            V[t] = np.sum((halves[:, t:] - halves[:, :-t]) ** 2, axis=(0, 1)) / float(m * (n - t))
        state = U._NeffState(D)
        lag0 = 1
        chunk = 1 + trial % 7
        while lag0 <= n - 1:
            nl = min(chunk, n - lag0)
            if U._finish_n_eff(var, [V[lag0 + k] for k in range(nl)], m, n, state):
                break
            lag0 += nl
        got = m * n / (1 + 2 * state.sum_rho)
        _, want = O.convergence_stats(x, 1, 0)
        np.testing.assert_allclose(got, want, rtol=1e-12)


def test_index_helper_mirrors():
    import utils as U
    for m in range(2, 513, 2):
        assert U.check_points(m).tolist() == O.check_points(m).tolist()
        for l in O.check_points(m):
            if l != 1:
                assert U.release(m, int(l)) == O.release(m, int(l))
                assert U.release_fast(m, int(l)) == O.release(m, int(l))
