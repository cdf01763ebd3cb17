"""World-size-2 gloo test (CPU) of the only exchange step on the path: the all-reduce of the split-chain moment
and variogram partial sums for Rhat / ESS.  Each rank holds a contiguous slab of chains; the partials
are produced here by numpy (on the GPU box they come from csrc/diag.cu), the reduce + finishing logic is the
product code (utils._stats_from_partials)."""
import os
import socket
import sys

import numpy as np
import pytest

from oracle import hmc_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, x, out_q, all_lags=False):
    sys.path.insert(0, os.path.join(ROOT, "understanding-hmc_b200"))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import utils as U
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    Nchain = x.shape[0]
    lo, hi = rank * Nchain // world, (rank + 1) * Nchain // world
    xl = x[lo:hi]
    n = xl.shape[1] // 2
    halves = xl[:, :2 * n].reshape((hi - lo) * 2, n, xl.shape[2])

    def moments_fn(buf):               # the record csrc/diag.cu produces: sums relative to the rank's own shift c (row 3)
        sd = np.std(halves, ddof=1, axis=1)
        c = halves[0, 0]
        mu = np.mean(halves, axis=1) - c
        buf[:4] = torch.from_numpy(np.stack([sd.sum(0), mu.sum(0), (mu ** 2).sum(0), c]))
        return 0

    def variogram_fn(lag0, nl, buf):
        rows = [np.sum((halves[:, t:] - halves[:, :-t]) ** 2, axis=(0, 1)) for t in range(lag0, lag0 + nl)]
        buf[:nl] = torch.from_numpy(np.stack(rows))

    def all_lags_fn(buf):              # what hmc_diag_variogram_all returns: the numerators of every lag 1..n-1
        rows = [np.sum((halves[:, t:] - halves[:, :-t]) ** 2, axis=(0, 1)) for t in range(1, n)]
        buf[:n - 1] = torch.from_numpy(np.stack(rows))

    R, ne = U._stats_from_partials(moments_fn, variogram_fn, n, x.shape[2], 2 * (hi - lo), group=None, lag_chunk=8,
                                   device=torch.device("cpu"), all_lags_fn=all_lags_fn if all_lags else None)
    out_q.put((rank, R, ne))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,offset,N,all_lags", [(2, 0.0, 80, False), (2, 1.0e7, 80, False),
                                                     (2, 0.0, 80, True),      # short chains: one windowed chunk, then the all-lags buffer
                                                     (2, 0.5, 400, True)])    # long chains (n >= FFT_MIN_N): the all-lags buffer straight away
def test_sharded_rhat_ess_equals_single_rank(world, offset, N, all_lags):
    import torch.multiprocessing as mp
    rng = np.random.RandomState(0)
    Nchain, D = 6, 4
    x = np.zeros((Nchain, N, D))
    for t in range(1, N):
        x[:, t] = 0.8 * x[:, t - 1] + rng.standard_normal((Nchain, D))
    x += rng.standard_normal((Nchain, 1, D)) * 0.3          # between-chain spread so that B matters
    x += offset                                             # chains far from zero: the between-chain sum must not cancel
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, x, q, all_lags)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    R0, ne0 = O.convergence_stats(x - offset, 1, 0)         # (Rhat / n_eff are shift invariant; the oracle's own sums are not robust)
    for rank, R, ne in results:
        np.testing.assert_allclose(R, R0, rtol=1e-8 if offset else 1e-10)
        np.testing.assert_allclose(ne, ne0, rtol=1e-6 if offset else 1e-9)
