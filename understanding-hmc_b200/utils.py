"""Host-side mirror of the reference's ``utils.py`` for the HMC hot path (B200 build).

Same public names as /root/reference/utils.py so that ``from utils import *`` in the case scripts resolves:
``np`` (with the removed aliases ``np.float`` / ``np.int`` restored, case1-script.py:28), ``time``,
``convergence_stats``, ``variogram``, ``acceptance_rate``, ``start_pts``, ``normal_lnL``, the NUTS index helpers
and -- when the optional libraries are installed -- ``plt``, ``mpl``, ``Ellipse``, ``multivariate_normal``,
``norm``, ``chi2`` (utils.py:1-19).  matplotlib is optional (plotting is out of scope).

``convergence_stats`` runs on the GPU (csrc/diag.cu through the C-ABI) -- there is no CPU implementation here.
"""
import time  # noqa: F401  (re-exported, utils.py:3)

import numpy as np

# Python-2-era numpy aliases the drivers use (case1-script.py:28 `dtype=np.float`).
if not hasattr(np, "float"):
    np.float = float
if not hasattr(np, "int"):
    np.int = int

try:  # utils.py:2, 12, 19 -- optional here
    import matplotlib as mpl
    mpl.use("Agg")
    import matplotlib.pyplot as plt
    from matplotlib.patches import Ellipse
    HAVE_MPL = True
except Exception:  # pragma: no cover - image has no matplotlib
    mpl = None
    plt = None
    Ellipse = None
    HAVE_MPL = False

from scipy.stats import multivariate_normal, norm, chi2  # noqa: E402,F401  (utils.py:5-7)

import hmc_b200_lib as _L  # noqa: E402


# ----------------------------------------------------------------------------------------------------------
# Plot helpers (utils.py:21-71): geometry is kept, drawing needs matplotlib.
# ----------------------------------------------------------------------------------------------------------
def cov_ellipse(cov, q=None, nsig=None, **kwargs):
    """Width, height and rotation of a covariance ellipse (utils.py:21-53)."""
    if q is not None:
        q = np.asarray(q)
    elif nsig is not None:
        q = 2 * norm.cdf(nsig) - 1
    else:
        raise ValueError("One of `q` and `nsig` should be specified.")
    r2 = chi2.ppf(q, 2)
    val, vec = np.linalg.eigh(cov)
    width, height = 2 * np.sqrt(val[:, None] * r2)
    rotation = np.degrees(np.arctan2(*vec[::-1, 0]))
    return width, height, rotation


def plot_cov_ellipse(ax, mus, covs, var_num1, var_num2, MoG_color="Blue", lw=2):
    """1- and 2-sigma ellipses (utils.py:55-71); a no-op notice without matplotlib."""
    if not HAVE_MPL:
        print("plot_cov_ellipse: matplotlib not available, skipped")
        return
    N_ellip = len(mus)
    for i in range(N_ellip):
        cov = np.asarray(covs[i])
        cov = [[cov[var_num1, var_num1], cov[var_num1, var_num2]], [cov[var_num2, var_num1], cov[var_num2, var_num2]]]
        mu = np.asarray(mus[i])
        mu = [mu[var_num1], mu[var_num2]]
        for j in [1, 2]:
            width, height, theta = cov_ellipse(cov, q=None, nsig=j)
            e = Ellipse(xy=mu, width=width, height=height, angle=theta, lw=lw)
            ax.add_artist(e)
            e.set_alpha(1)
            e.set_facecolor("none")
            e.set_edgecolor(MoG_color)
    return


# ----------------------------------------------------------------------------------------------------------
# Convergence statistics on the GPU (utils.py:77-179)
# ----------------------------------------------------------------------------------------------------------
class _NeffState(object):
    """Per-dimension state of the sequential truncation rule (vectorised over dimensions)."""

    def __init__(self, D):
        self.rho = np.zeros((0, D))             # rho[k] = autocorrelation at lag k + 1, for the lags received so far
        self.t = 1                              # common loop index of the dimensions still running (utils.py:141)
        self.started = False
        self.done = np.zeros(D, dtype=bool)
        self.t_stop = np.zeros(D, dtype=int)    # number of leading rho terms summed (utils.py:154)
        self.sum_rho = np.zeros(D)

    def close(self, mask, t_stop):
        """Finish the dimensions in ``mask``: sum_rho = max(0, sum(rho[:t_stop])) (utils.py:154-156)."""
        if not np.any(mask):
            return
        t_stop = np.broadcast_to(t_stop, mask.shape)
        csum = np.vstack([np.zeros((1, self.rho.shape[1])), np.cumsum(self.rho, axis=0)])
        idx = np.nonzero(mask)[0]
        s = csum[np.minimum(t_stop[idx], self.rho.shape[0]), idx]
        self.sum_rho[idx] = np.where(s < 0, 0.0, s)
        self.t_stop[idx] = t_stop[idx]
        self.done[idx] = True


def _finish_n_eff(var, V_rows, m, n, state):
    """The sequential truncation rule of utils.py:130-157 applied to the lags received so far, all dimensions at once.

    ``state``: a _NeffState.  Returns True when every dimension is done."""
    st = state
    rho_new = 1. - np.asarray(V_rows, dtype=float).reshape(len(V_rows), -1) / (2 * var)      # utils.py:134-135, 144
    st.rho = np.vstack([st.rho, rho_new])
    have = st.rho.shape[0]
    if not st.started and have >= 2:
        st.started = True
        early = (~st.done) & ((st.rho[0] < 1e-2) | (st.rho[0] < 1e-2))       # Q2: rho_t1 tested twice (utils.py:136)
        st.sum_rho[early] = 0
        st.done |= early
    if not st.started:
        return False
    hi = min(n - 2, have - 1)                                               # the loop needs rho[t + 1] (utils.py:141-152)
    if hi > st.t:
        ts = np.arange(st.t, hi)
        cond = ((ts % 2) == 1)[:, None] & ((st.rho[ts] + st.rho[ts + 1]) < 0)
        hit = cond.any(axis=0) & ~st.done
        st.close(hit, ts[cond.argmax(axis=0)])
        st.t = hi
    if st.t >= n - 2:
        st.close(~st.done, np.full(st.done.shape, st.t))
    return bool(np.all(st.done))


def _stats_from_partials(moments_fn, variogram_fn, n, D, m_local, group=None, lag_chunk=32):
    """Rhat / n_eff from per-rank partial sums (utils.py:107-157).

    ``moments_fn()`` -> float64 tensor (3, D): sum_j std_j, sum_j mean_j, sum_j mean_j^2 over the LOCAL split
    chains; ``variogram_fn(lag0, nl)`` -> float64 tensor (nl, D) of local variogram numerators.  When
    torch.distributed is initialised (and ``group is not False``) the partials are all-reduced (sum) -- the only
    collective on the whole path; everything after that is O(lags * D) host arithmetic."""
    import torch
    import torch.distributed as dist
    distributed = dist.is_available() and dist.is_initialized() and (group is not False)
    grp = None if group in (None, False) else group
    mom = moments_fn()
    cnt = torch.tensor([float(m_local)], dtype=torch.float64, device=mom.device)
    if distributed:
        dist.all_reduce(mom, group=grp)
        dist.all_reduce(cnt, group=grp)
    mom_h = mom.cpu().numpy()
    m = int(round(float(cnt.item())))
    W = mom_h[0] / m                                                # utils.py:112 (mean of std, Q1)
    mean_all = mom_h[1] / m                                         # utils.py:119
    B = (mom_h[2] - m * mean_all ** 2) * n / float(m - 1)           # utils.py:120
    var = W * (n - 1) / float(n) + B / float(n)                     # utils.py:123
    R = np.sqrt(var / W)                                            # utils.py:126

    state = _NeffState(D)
    lag0 = 1
    max_lag = n - 1
    while lag0 <= max_lag:
        nl = min(lag_chunk, max_lag - lag0 + 1)
        buf = variogram_fn(lag0, nl)
        if distributed:
            dist.all_reduce(buf, group=grp)
        rows = buf[:nl].cpu().numpy()
        V_rows = [rows[k] / float(m * (n - (lag0 + k))) for k in range(nl)]      # utils.py:177
        if _finish_n_eff(var, V_rows, m, n, state):
            break
        lag0 += nl
    state.close(~state.done, np.full(D, state.t))        # chains too short to ever start (n < 3): use what exists
    n_eff = m * n / (1 + 2 * state.sum_rho)                                             # utils.py:157
    return R, n_eff


def _device_stats(x, n, group=None, lag_chunk=32):
    """x: CUDA tensor (Nchain_local, >=2n, D), float32/float64, last dim contiguous.  Returns (R, n_eff) numpy.
    The per-device partial sums come from csrc/diag.cu through the C-ABI."""
    import torch
    lib = _L.load()
    assert x.is_cuda and x.dim() == 3 and x.stride(2) == 1 and x.stride(1) == x.shape[2]
    Nchain, _, D = x.shape
    dtype = _L.HMC_F32 if x.dtype == torch.float32 else _L.HMC_F64
    stride_chain = x.stride(0)

    fused = {}

    def moments_fn():
        mom = torch.empty((3, D), dtype=torch.float64, device=x.device)
        if 2 <= n <= 32 and lag_chunk == 32 and D <= 128:    # short series: moments and every lag in one pass over the samples
            buf = torch.empty((lag_chunk, D), dtype=torch.float64, device=x.device)
            _L.check(lib.hmc_diag_short_series(dtype, _L.ptr(x), Nchain, n, D, stride_chain, max(1, n - 1), _L.ptr(mom),
                                               _L.ptr(buf), _L.current_stream_ptr()))
            fused["lags"] = buf
            return mom
        _L.check(lib.hmc_diag_moments(dtype, _L.ptr(x), Nchain, n, D, stride_chain, _L.ptr(mom), _L.current_stream_ptr()))
        return mom

    def variogram_fn(lag0, nl):
        if lag0 == 1 and "lags" in fused:
            return fused.pop("lags")
        buf = torch.empty((lag_chunk, D), dtype=torch.float64, device=x.device)
        _L.check(lib.hmc_diag_variogram(dtype, _L.ptr(x), Nchain, n, D, stride_chain, lag0, nl, _L.ptr(buf),
                                        _L.current_stream_ptr()))
        return buf

    return _stats_from_partials(moments_fn, variogram_fn, n, D, 2 * Nchain, group=group, lag_chunk=lag_chunk)


def convergence_stats(q_chain, thin_rate=5, warm_up_num=0, group=None):
    """Split-chain Rhat and variogram ESS per dimension (utils.py:77-159), computed on the GPU.

    ``q_chain``: (Nchain, Niter, D) numpy array or CUDA torch tensor.  Reference quirks kept: W is a mean of
    standard deviations (Q1), ``rho_t1`` is tested twice (Q2)."""
    import torch
    if isinstance(q_chain, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(q_chain)).cuda()
    else:
        x = q_chain
    Nchain, Niter, D = x.shape
    assert Nchain > 1                                               # utils.py:85
    x = x[:, warm_up_num:, :][:, ::thin_rate, :]                    # utils.py:91-94
    L_chain = x.shape[1]
    if (L_chain % 2) != 0:                                          # utils.py:96-99
        x = x[:, :L_chain - 1]
    n = L_chain // 2                                                # utils.py:102
    if not (x.stride(2) == 1 and x.stride(1) == D):
        x = x.contiguous()
    return _device_stats(x, n, group=group)


def variogram(chains, var_num, t_lag):
    """V_t of BDA3 (11.7) for one variable and one lag (utils.py:161-179), on the GPU."""
    import torch
    lib = _L.load()
    m = len(chains)
    n = chains[0].shape[0]
    x = torch.from_numpy(np.ascontiguousarray(np.stack([np.asarray(c)[:, var_num] for c in chains])[:, :, None])).cuda()
    # each "chain" of length n is treated as one split chain: use n_half = n with a dummy second half of stride 0
    x2 = torch.cat([x, x], dim=1).contiguous()                      # halves [0,n) and [n,2n) are the same series
    out = torch.empty((1, 1), dtype=torch.float64, device=x.device)
    _L.check(lib.hmc_diag_variogram(_L.HMC_F64, _L.ptr(x2), m, n, 1, 2 * n, t_lag, 1, _L.ptr(out),
                                    _L.current_stream_ptr()))
    return float(out.item()) / 2.0 / float(m * (n - t_lag))


def acceptance_rate(decision_chain, start=None, end=None):
    """Mean of a decision record (utils.py:183-200)."""
    _, Niter, _ = decision_chain.shape
    if start is None and end is None:
        return np.sum(decision_chain, axis=(1, 2)) / Niter
    if end > 0:
        Niter = end - start
    else:
        Niter = Niter - start
    return np.sum(decision_chain[:, start:end, :], axis=(1, 2)) / Niter


def start_pts(q0, cov0, size):
    """Starting points ~ N(q0, cov0) (utils.py:204-209)."""
    return np.random.multivariate_normal(q0, cov0, size=size)


def normal_lnL(q, q0, cov0):
    """Multivariate-normal log-density (utils.py:213-218)."""
    return multivariate_normal.logpdf(q, mean=q0, cov=cov0)


# ----------------------------------------------------------------------------------------------------------
# NUTS index helpers (utils.py:222-311, 367-385) in closed form; the CUDA kernel uses the same formulas.
# ----------------------------------------------------------------------------------------------------------
def find_next(table):
    for i, e in enumerate(table):
        if e == -1:
            return i


def retrieve_save_index(table, l):
    for i, m in enumerate(table):
        if m == l:
            return i


def power_of_two(r):
    assert type(r) == int
    return (r & (r - 1)) == 0


def power_of_two_fast(r):
    r = int(r)
    return (r & (r - 1)) == 0


def _tz(m):
    m = int(m)
    return (m & -m).bit_length() - 1


def check_points(m):
    """Points against which even point m is U-turn checked: m - 2^j + 1 for j = tz(m) .. 1 (utils.py:246-283)."""
    assert (m % 2) == 0
    return np.asarray([int(m) - (1 << j) + 1 for j in range(_tz(m), 0, -1)])


def release(m, l):
    """True if check point l is no longer needed after the check at m (utils.py:286-304)."""
    assert (l != 1) and (m % 2) == 0
    return release_fast(m, l)


def release_fast(m, l):
    m, l = int(m), int(l)
    return (l > 1) and (l != m - (1 << _tz(m)) + 1)


def rvs(dim=3):
    """Haar-distributed random rotation matrix (same role as utils.py:424-441; QR of a Gaussian matrix with the
    sign convention fixed, determinant forced to +1)."""
    A = np.random.normal(size=(dim, dim))
    Q, Rm = np.linalg.qr(A)
    Q = Q * np.sign(np.diag(Rm))
    if np.linalg.det(Q) < 0:
        Q[:, 0] = -Q[:, 0]
    return Q
