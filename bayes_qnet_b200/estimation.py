"""Estimation drivers with the reference's signatures (estimation.py:29-257): ``bayes``, ``sem``,
``sample_departures``.  Each outer iteration of the reference -- one sweep over the hidden times, then one
parameter update -- is one fused step of the GPU engine, and all iterations of a call run in a single
kernel launch.  ``n_chains > 1`` runs that many independent chains from the same initial state (new; the
reference is single-chain): the return value stays ``(mus, arrvl)`` for chain 0, and the whole ensemble is
available from ``last_run()``.
"""
import time

import numpy

from . import qnet as _qnet
from . import diagnostics
from .engine import GibbsEngine

MH, SLICE = 0, 1
_sampling_type = SLICE
_last = None


def set_sampler(name):
    """sampling_type: SLICE (default) or MH (estimation.py:20-25)."""
    global _sampling_type
    if name == "SLICE":
        _sampling_type = SLICE
    elif name == "MH":
        _sampling_type = MH
    else:
        raise Exception("Could not understand estimator %s" % name)


def _mode():
    return "SLICE" if _sampling_type == SLICE else "MH"


class Run(object):
    """Everything one estimation call produced, for all chains."""

    def __init__(self, engine, theta, mean_wait, logp, seconds):
        self.engine = engine
        self.theta = theta          # [n_iter, n_chains, n_params]
        self.mean_wait = mean_wait  # [n_iter, n_chains, n_queues]
        self.logp = logp            # [n_iter, n_chains]
        self.seconds = seconds
        self.stats = engine.stats()

    def rhat(self):
        return diagnostics.rhat(self.theta)

    def ess(self):
        return diagnostics.ess(self.theta)

    def posterior_mean(self):
        return self.theta.mean(axis=(0, 1))


def last_run():
    return _last


def _run(net, arrv, num_iter, update, n_chains, seed, device, report_fn, return_arrv, gibbs_iter=1, reverse=False,
         burn=0, burn_report_fn=None, momentum=0):
    """num_iter outer iterations of sweep(s) + parameter update.  Without per-iteration callbacks the whole run is
    one kernel launch; with them every iteration is launched separately and the callbacks see what the reference's
    see: in ``bayes`` the state after the sweep with the parameters the sweep USED (estimation.py:148-155: report_fn
    runs before sample_parameters), in ``sem`` the state with the NEW estimate (estimation.py:65-79)."""
    global _last
    reverse = bool(reverse) and _mode() == "SLICE"  # the reference only reverses slice sweeps (estimation.py:135)
    eng = GibbsEngine(net, arrv, n_chains=n_chains, seed=seed, device=device)
    t0 = time.time()
    all_sampled = []
    stepwise = bool(report_fn or burn_report_fn or return_arrv or gibbs_iter != 1 or momentum)
    if stepwise:
        thetas = []
        mu_old = None
        for i in range(num_iter):
            if gibbs_iter > 1:
                eng.sweep(gibbs_iter - 1, _mode(), "NONE", tmax_per_sweep=False, reverse=reverse)
            split = (update == "BAYES") and bool(report_fn or return_arrv)
            eng.sweep(1, _mode(), "NONE" if split else update, trace=True, reverse=reverse)
            if update == "STEM" and momentum > 0 and mu_old is not None:  # estimation.py:64-66
                eng.set_params(momentum * mu_old + (1.0 - momentum) * eng.params())
            if not split:
                net.parameters = list(eng.params(0, 1)[0])
            cb = report_fn if i >= burn else burn_report_fn
            if cb or (return_arrv and i >= burn):
                snap = eng.to_arrivals(0)
                if cb is report_fn and cb:
                    cb(net, snap, i, float(eng.traces()[2][-1, 0]))
                elif cb:
                    cb(net, snap, i)  # burn_report_fn takes no log-prob (estimation.py:79)
                if return_arrv and i >= burn:
                    all_sampled.append(snap)
            if split:  # posterior draw after the callback, as the reference orders them
                _qnet.capi_update_only(eng, "BAYES")
                net.parameters = list(eng.params(0, 1)[0])
            mu_old = eng.params()
            thetas.append(mu_old)
        _, mw, lp = eng.traces()
        theta = numpy.stack(thetas) if thetas else numpy.empty((0, n_chains, eng.P))
    else:
        eng.sweep(num_iter, _mode(), update, trace=True, reverse=reverse)
        theta, mw, lp = eng.traces()
    secs = time.time() - t0
    eng.write_back(arrv, 0)
    net.parameters = list(eng.params(0, 1)[0])
    st = eng.stats()
    _qnet._absorb(dict(nevents=0, nreject=0, num_diff=0, diff_len=0), st)
    _last = Run(eng, theta[burn:], mw[burn:], lp[burn:], secs)
    mus = [list(theta[i, 0, :]) for i in range(burn, theta.shape[0])]
    if not return_arrv:
        all_sampled = [arrv]
    return mus, all_sampled


def bayes(net, arrv, num_iter, report_fn=None, return_arrv=False, reverse=False, sweeten=0, tag="MU",
          n_chains=1, seed=None, device=0):
    """Gibbs sampling of hidden times and service parameters (estimation.py:104-169)."""
    _qnet.reset_stats()
    if sweeten:
        net.slice_resample(arrv, sweeten, reverse=reverse)
    net.estimate(arrv)  # initialise the parameters from the Gibbs initialisation (estimation.py:124)
    return _run(net, arrv, num_iter, "BAYES", n_chains, seed, device, report_fn, return_arrv, reverse=reverse)


def _logsumexp(v):
    v = numpy.asarray(v, dtype=numpy.float64)
    m = numpy.max(v)
    if not numpy.isfinite(m):
        return float(m)
    return float(m + numpy.log(numpy.sum(numpy.exp(v - m))))


def chib_evidence(net, obs, M_special=100, M=100, debug=False, seed=None, device=0, strict=False):
    """Chib's estimate of the log evidence log p(observed) with the Murray-Salakhutdinov correction
    (estimation.py:274-330): p(v) = p(v, h*, theta*) / p(h*, theta* | v), the denominator averaged over a forward and
    a backward (reverse-scan) chain started at the special state of the transition densities
    Qnet.gibbs_kernel + Qnet.parameter_kernel into it.  Slice sampler only.
    strict=False (default) is the reference's transition density: events whose target lies outside their support
    when their turn comes are retried after the others (qnet.pyx:756-805).  That is not the density of the sweep
    itself and biases the estimate by 0.15-0.25 nats on the reference's own closed-form check (tests/test_gpu_chib.py);
    strict=True makes such a transition impossible (density 0), with which Chib's identity holds."""
    if _mode() != "SLICE":
        raise NotImplementedError("chib_evidence with the MH sampler")
    net.estimate(obs)
    rs = numpy.random.RandomState(seed)
    with GibbsEngine(net, obs, n_chains=1, seed=seed, device=device) as eng:
        # 1. the special state: the most probable one of a first run (Reporter, estimation.py:343-366)
        best = (-numpy.inf, None, None)
        for i in range(M_special):
            eng.sweep(1, "SLICE", "NONE")
            lp = float(eng.log_prob()[0])
            if i > M_special // 2 and lp > best[0]:
                best = (lp, eng.state(0, 1)[0].copy(), eng.params(0, 1)[0].copy())
            _qnet.capi_update_only(eng, "BAYES")
        if best[1] is None:
            best = (float(eng.log_prob()[0]), eng.state(0, 1)[0].copy(), eng.params(0, 1)[0].copy())
        _, d_star, theta_star = best
        arrv_star = obs.duplicate()
        eng.set_state(d_star)
        eng.set_params(theta_star)
        eng.write_back(arrv_star, 0)
        p_v_hstar = float(eng.log_prob()[0])  # log p(v, h* | theta*), improper prior on theta as in the reference
        # 2./3. forward chain of M - u steps, backward chain of u steps, both from the special state
        u = int(rs.randint(M))
        cnd = []
        for n_steps, reverse in ((M - u, False), (u, True)):
            eng.set_state(d_star)
            eng.set_params(theta_star)
            for _ in range(n_steps):
                eng.sweep(1, "SLICE", "NONE", reverse=reverse)
                theta = list(eng.params(0, 1)[0])
                snap = eng.to_arrivals(0)
                val = float(eng.gibbs_kernel(d_star, strict=strict)[0]) + net.parameter_kernel(theta, list(theta_star), snap)
                cnd.append(val)
                _qnet.capi_update_only(eng, "BAYES")
        net.parameters = list(theta_star)
    p_hstar_given_v = _logsumexp(cnd) - numpy.log(M)
    if debug:
        print("CHIB log(p (v, h*)) = %.5f\nCHIB log(p (h*|v)) = %.5f" % (p_v_hstar, p_hstar_given_v))
    return p_v_hstar - p_hstar_given_v


def sem(net, arrv, burn, num_iter, gibbs_iter=1, momentum=0, return_arrv=False, report_fn=None,
        burn_report_fn=None, debug=False, n_chains=1, seed=None, device=0):
    """Stochastic EM (estimation.py:29-100): burn + num_iter iterations of gibbs_iter sweeps followed by the ML
    update, optionally damped -- theta <- momentum * theta_old + (1 - momentum) * theta_ML (estimation.py:64-66).
    report_fn(net, arrv, i, lp) runs after every post-burn iteration, burn_report_fn(net, arrv, i) before that."""
    _qnet.reset_stats()
    net.estimate(arrv)
    return _run(net, arrv, burn + num_iter, "STEM", n_chains, seed, device, report_fn, return_arrv, gibbs_iter,
                burn=burn, burn_report_fn=burn_report_fn, momentum=float(momentum))


def sample_departures(net, arrv, num_iter, report_fn=None, debug=False, n_chains=1, seed=None, device=0):
    """Sweeps with the parameters held fixed (estimation.py:171-224)."""
    _qnet.reset_stats()
    return _run(net, arrv, num_iter, "NONE", n_chains, seed, device, report_fn, True)
