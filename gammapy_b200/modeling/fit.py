"""``Fit`` front end driving the batched GPU likelihood (modeling/fit.py:147-519 in the reference).

iminuit is not a dependency.  The optimiser is a variable-metric (BFGS) minimiser in the reference's
*factor* space (modeling/iminuit.py:23-51: ``MinuitLikelihood.fcn(*factors)``, errordef 1, tol 0.1 ->
EDM goal 2e-4) whose gradient (central differences) and line search are evaluated as ONE batched launch
each through ``Datasets.stat_sum_batch``; ``covariance`` is a batched finite-difference Hesse.
``stat_profile`` / ``stat_surface`` (fit.py:351-519) hand all scan points to a single launch.
"""
import itertools
import logging

import numpy as np

from .parameter import store_covariance

log = logging.getLogger(__name__)

__all__ = ["Fit", "FitResult", "OptimizeResult", "CovarianceResult"]


class _Likelihood:
    """modeling/likelihood.py + iminuit.py:23-31: factors in, joint stat out (single or batched)."""

    def __init__(self, datasets):
        from ..datasets import Datasets, MapDataset

        if hasattr(datasets, "_fit_engine"):  # a dataset with its own statistic (MapDatasetOnOff: WStat)
            pass
        elif isinstance(datasets, MapDataset):
            datasets = Datasets([datasets])
        elif not isinstance(datasets, Datasets) and not hasattr(datasets, "_get_engine"):
            datasets = Datasets(list(datasets))
        self.datasets = datasets
        self.engine = datasets._fit_engine() if hasattr(datasets, "_fit_engine") else datasets._get_engine()
        self.free = [p for p in self.engine.parameters if not p.frozen]
        self.cols = np.array([self.engine.parameters.index(p) for p in self.free], dtype=int)
        self.has_priors = any(p.prior is not None for p in self.engine.parameters)
        self.n_calls = 0
        self.n_launches = 0

    @property
    def factors(self):
        return np.array([p.factor for p in self.free], dtype=float)

    def values_from_factors(self, factors):
        """[B, n_free] factors -> [B, n_par] values via Parameter.inverse_transform (parameter.py:564-574)."""
        factors = np.atleast_2d(np.asarray(factors, dtype=float))
        vals = np.tile(self.engine.values(), (factors.shape[0], 1))
        for j, p in enumerate(self.free):
            vals[:, self.cols[j]] = p.inverse_transform(factors[:, j])
        return vals

    def set_factors(self, factors):
        for p, f in zip(self.free, factors):
            p.factor = f

    def _prior_terms(self, vals):
        """Prior statistic of every row of ``vals`` [B, n_par] (datasets/core.py:266-271: added to the joint stat)."""
        total = np.zeros(vals.shape[0])
        for col, par in enumerate(self.engine.parameters):
            if par.prior is not None:
                total += par.prior.evaluate_many(vals[:, col])
        return total

    def stat_values(self, vals):
        """Joint statistic (likelihood + priors) of full parameter-value vectors ``vals`` [B, n_par], one launch."""
        out = np.atleast_1d(self.engine.stat_sum(vals)).astype(float)
        if self.has_priors:
            out = out + self._prior_terms(np.atleast_2d(vals))
        return out

    def fcn(self, factors):
        self.set_factors(factors)
        self.n_calls += 1
        self.n_launches += 1
        value = self.engine.stat_sum()
        if self.has_priors:
            value += float(self._prior_terms(self.engine.values()[np.newaxis, :])[0])
        return value

    def fcn_batch(self, factors, chunk=512, base=None):
        """``base``: the point the rows of ``factors`` are small departures from (finite differences).  It rides
        along as vector 0 of every launch, so that the library evaluates only the tiles on which a row differs
        from it (the items equal to vector 0's are shared, see ``dedup_*_kernel``); its value is not returned."""
        vals = self.values_from_factors(factors)
        out = np.empty(vals.shape[0], dtype=float)
        lead = None if base is None else self.values_from_factors(np.asarray(base, dtype=float)[np.newaxis, :])
        for i in range(0, vals.shape[0], chunk):
            block = vals[i:i + chunk]
            if lead is not None:
                out[i:i + chunk] = np.atleast_1d(self.engine.stat_sum(np.concatenate([lead, block])))[1:]
            else:
                out[i:i + chunk] = np.atleast_1d(self.engine.stat_sum(block))
            self.n_launches += 1
        if self.has_priors:
            out += self._prior_terms(vals)
        self.n_calls += vals.shape[0]
        return out


class OptimizeResult:
    def __init__(self, parameters, total_stat, nfev, success, message, trace=None, backend="bfgs-b200", method="bfgs"):
        self.parameters = parameters
        self.total_stat = total_stat
        self.nfev = nfev
        self.success = success
        self.message = message
        self.trace = trace
        self.backend = backend
        self.method = method

    def __repr__(self):
        return (f"OptimizeResult(backend={self.backend!r}, success={self.success}, message={self.message!r}, "
                f"nfev={self.nfev}, total_stat={self.total_stat:.4f})")


class CovarianceResult:
    def __init__(self, matrix, success, message, backend="hesse-b200"):
        self.matrix = matrix
        self.success = success
        self.message = message
        self.backend = backend


class FitResult:
    def __init__(self, optimize_result=None, covariance_result=None):
        self.optimize_result = optimize_result
        self.covariance_result = covariance_result

    @property
    def parameters(self):
        return self.optimize_result.parameters

    @property
    def total_stat(self):
        return self.optimize_result.total_stat

    @property
    def success(self):
        ok = self.optimize_result.success
        if self.covariance_result is not None:
            ok = ok and self.covariance_result.success
        return ok

    @property
    def nfev(self):
        return self.optimize_result.nfev

    def __repr__(self):
        return f"FitResult({self.optimize_result!r})"


def _bounds(like):
    lo = np.array([p.factor_min if not np.isnan(p.min) else -np.inf for p in like.free])
    hi = np.array([p.factor_max if not np.isnan(p.max) else np.inf for p in like.free])
    return lo, hi


def _steps(like):
    """Initial finite-difference steps in factor space: the parameter error if known, else 1 % (Minuit's default)."""
    steps = []
    for p in like.free:
        if p.error > 0:
            s = abs(p.error / p.scale)
        else:
            s = 0.01 * max(abs(p.factor), 1e-3) if p.factor != 0 else 0.01
        steps.append(s)
    return np.array(steps)


def _gradient(like, x, f0, h, lo, hi):
    """Central differences with all 2 n points in one launch; also returns the diagonal second derivatives."""
    n = len(x)
    pts = np.tile(x, (2 * n, 1))
    hp = np.minimum(h, hi - x)
    hm = np.minimum(h, x - lo)
    hp = np.where(hp <= 0, 0.0, hp)
    hm = np.where(hm <= 0, 0.0, hm)
    for i in range(n):
        pts[2 * i, i] += hp[i]
        pts[2 * i + 1, i] -= hm[i]
    f = like.fcn_batch(pts, base=x)
    fp, fm = f[0::2], f[1::2]
    with np.errstate(divide="ignore", invalid="ignore"):
        g = (fp - fm) / (hp + hm)  # one-sided when hp or hm is 0 (the point itself is evaluated: f0)
        g2 = 2 * (hm * fp + hp * fm - (hp + hm) * f0) / (hp * hm * (hp + hm))
    one_sided = (hp == 0) | (hm == 0)
    g = np.where(np.isfinite(g), g, 0.0)
    g2 = np.where(one_sided | ~np.isfinite(g2), np.nan, g2)
    return g, g2


def _minimize(like, tol=0.1, max_iter=200, store_trace=False):
    """BFGS with batched gradient + batched line search; stops on Minuit's EDM criterion."""
    x = like.factors.copy()
    n = len(x)
    lo, hi = _bounds(like)
    x = np.clip(x, lo, hi)
    edm_goal = 0.002 * tol * 1.0
    trace = []
    f = like.fcn(x)
    if n == 0:
        return x, f, True, "no free parameters", trace
    h = _steps(like)
    g, g2 = _gradient(like, x, f, h, lo, hi)
    # initial inverse Hessian from the diagonal curvature (what Migrad seeds from)
    diag = np.where(np.isfinite(g2) & (g2 > 0), 1.0 / np.where(g2 > 0, g2, 1.0), h * h)
    H = np.diag(diag)
    alphas = np.array([1.0, 0.5, 0.25, 0.1, 0.03, 0.01, 1e-3, 1.5, 2.0, 4.0])
    message, success = "maximum number of iterations reached", False
    span = np.where(np.isfinite(hi - lo), hi - lo, 1.0)
    for it in range(max_iter):
        # Active set: a component sitting on a limit whose descent direction points out of the box is blocked.  It
        # leaves the gradient, the metric and the EDM (Minuit gets the same effect from its sin-transform, where
        # the internal derivative vanishes on a limit); it is released as soon as its gradient turns inward.
        blocked = ((x <= lo + 1e-12 * span) & (g > 0)) | ((x >= hi - 1e-12 * span) & (g < 0))
        gp = np.where(blocked, 0.0, g)
        Hp = H.copy()
        Hp[blocked, :] = 0.0
        Hp[:, blocked] = 0.0
        edm = 0.5 * float(gp @ Hp @ gp)
        if store_trace:
            trace.append((it, f, x.copy()))
        if edm < edm_goal and it > 0:
            success, message = True, f"Optimization terminated successfully (EDM = {edm:.3g})."
            break
        p = -Hp @ gp
        if not np.all(np.isfinite(p)) or gp @ p >= 0:
            H = np.diag(diag)
            Hp = np.diag(np.where(blocked, 0.0, diag))
            p = -Hp @ gp
        cand = np.clip(x[np.newaxis, :] + alphas[:, np.newaxis] * p[np.newaxis, :], lo, hi)
        fc = like.fcn_batch(cand)
        fc = np.where(np.isfinite(fc), fc, np.inf)
        best = int(np.argmin(fc))
        if not fc[best] < f:
            # shrink further before giving up
            small = np.array([3e-4, 1e-4, 1e-5, 1e-6])
            cand = np.clip(x[np.newaxis, :] + small[:, np.newaxis] * p[np.newaxis, :], lo, hi)
            fc = np.where(np.isfinite(fc2 := like.fcn_batch(cand)), fc2, np.inf)
            best = int(np.argmin(fc))
            if not fc[best] < f:
                if edm < 10 * edm_goal:
                    success, message = True, f"Optimization terminated: no further decrease (EDM = {edm:.3g})."
                else:
                    message = f"line search failed (EDM = {edm:.3g})"
                break
        x_new, f_new = cand[best], float(fc[best])
        # refresh steps from the current metric (like Minuit: step ~ 0.1 sigma..)
        h = np.maximum(0.1 * np.sqrt(np.abs(np.diag(H))), 1e-8 * np.maximum(np.abs(x_new), 1.0))
        g_new, g2 = _gradient(like, x_new, f_new, h, lo, hi)
        s, y = x_new - x, g_new - g
        # the metric learns from the free components only (a clipped step says nothing about the curvature there)
        s = np.where(blocked, 0.0, s)
        y = np.where(blocked, 0.0, y)
        sy = float(s @ y)
        if sy > 1e-12 * np.linalg.norm(s) * np.linalg.norm(y):
            rho = 1.0 / sy
            I = np.eye(n)
            H = (I - rho * np.outer(s, y)) @ H @ (I - rho * np.outer(y, s)) + rho * np.outer(s, s)
        x, f, g = x_new, f_new, g_new
    like.set_factors(x)
    return x, f, success, message, trace


def _hesse(like, x, f0):
    """Finite-difference Hessian in factor space, batched.  Central differences (1 + 2n + 2n(n-1) points) where the
    box leaves room on both sides; a parameter on (or within a step of) a limit uses the inward one-sided stencils
    f(x), f(x + a), f(x + 2a) and, for the mixed terms, the forward formula, so no point leaves the limits."""
    n = len(x)
    lo, hi = _bounds(like)
    h = _steps(like)
    # refine steps so that the stat changes by ~errordef along each axis
    g, g2 = _gradient(like, x, f0, h, lo, hi)
    good = np.isfinite(g2) & (g2 > 0)
    h = np.where(good, np.sqrt(2.0 * 0.1 / np.where(good, g2, 1.0)), h)
    room_p, room_m = hi - x, x - lo
    central = (room_p >= h) & (room_m >= h)
    # one-sided: direction with more room, step such that x + 2 a stays inside
    sign = np.where(room_p >= room_m, 1.0, -1.0)
    room = np.maximum(room_p, room_m)
    a = np.where(central, h, sign * np.minimum(h, 0.5 * room))
    a = np.where(a == 0, _steps(like), a)
    pts = []
    for i in range(n):
        for k in ((+1, -1) if central[i] else (1, 2)):
            p = x.copy()
            p[i] += k * a[i]
            pts.append(p)
    pairs = list(itertools.combinations(range(n), 2))
    for i, j in pairs:
        if central[i] and central[j]:
            combos = ((1, 1), (1, -1), (-1, 1), (-1, -1))
        else:
            combos = ((1, 1),)
        for si, sj in combos:
            p = x.copy()
            p[i] += si * a[i]
            p[j] += sj * a[j]
            pts.append(p)
    f = like.fcn_batch(np.array(pts), base=x) if pts else np.zeros(0)
    like.set_factors(x)
    hess = np.zeros((n, n))
    f1 = np.zeros(n)  # f(x + a_i e_i)
    for i in range(n):
        if central[i]:
            hess[i, i] = (f[2 * i] - 2 * f0 + f[2 * i + 1]) / (a[i] * a[i])
        else:
            hess[i, i] = (f0 - 2 * f[2 * i] + f[2 * i + 1]) / (a[i] * a[i])
        f1[i] = f[2 * i]
    k = 2 * n
    for i, j in pairs:
        if central[i] and central[j]:
            fpp, fpm, fmp, fmm = f[k: k + 4]
            hess[i, j] = hess[j, i] = (fpp - fpm - fmp + fmm) / (4 * a[i] * a[j])
            k += 4
        else:
            hess[i, j] = hess[j, i] = (f[k] - f1[i] - f1[j] + f0) / (a[i] * a[j])
            k += 1
    return hess


def _brentq(f, xa, xb, xtol=2e-12, rtol=8.881784197001252e-16, maxiter=100):
    """Brent's root bracketing as scipy.optimize.brentq implements it (what utils/roots.py:find_roots calls with
    nbin=1).  Returns (root, iterations, converged); (nan, 0, False) when f(xa) and f(xb) do not bracket a root."""
    xpre, xcur = float(xa), float(xb)
    xblk = fblk = spre = scur = 0.0
    fpre, fcur = f(xpre), f(xcur)
    if not (np.isfinite(fpre) and np.isfinite(fcur)):
        return np.nan, 0, False
    if fpre == 0:
        return xpre, 0, True
    if fcur == 0:
        return xcur, 0, True
    if np.sign(fpre) == np.sign(fcur):
        return np.nan, 0, False
    for it in range(1, maxiter + 1):
        if fpre != 0 and fcur != 0 and np.sign(fpre) != np.sign(fcur):
            xblk, fblk = xpre, fpre
            spre = scur = xcur - xpre
        if abs(fblk) < abs(fcur):
            xpre, xcur, xblk = xcur, xblk, xcur
            fpre, fcur, fblk = fcur, fblk, fcur
        delta = (xtol + rtol * abs(xcur)) / 2
        sbis = (xblk - xcur) / 2
        if fcur == 0 or abs(sbis) < delta:
            return xcur, it, True
        if abs(spre) > delta and abs(fcur) < abs(fpre):
            if xpre == xblk:
                stry = -fcur * (xcur - xpre) / (fcur - fpre)
            else:
                dpre = (fpre - fcur) / (xpre - xcur)
                dblk = (fblk - fcur) / (xblk - xcur)
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre))
            if 2 * abs(stry) < min(abs(spre), 3 * abs(sbis) - delta):
                spre, scur = scur, stry
            else:
                spre = scur = sbis
        else:
            spre = scur = sbis
        xpre, fpre = xcur, fcur
        xcur += scur if abs(scur) > delta else (delta if sbis > 0 else -delta)
        fcur = f(xcur)
        if not np.isfinite(fcur):
            return np.nan, it, False
    return xcur, maxiter, False


class Fit:
    """Fit class (modeling/fit.py:81).  ``backend`` is informational: the optimiser is the built-in batched BFGS."""

    def __init__(self, backend="minuit", optimize_opts=None, covariance_opts=None, confidence_opts=None,
                 store_trace=False):
        self.backend = backend
        self.optimize_opts = dict(optimize_opts or {})
        self.covariance_opts = dict(covariance_opts or {})
        self.confidence_opts = dict(confidence_opts or {})
        self.store_trace = store_trace
        self._last_like = None

    def run(self, datasets):
        """Optimize then estimate the covariance (fit.py:147-181)."""
        opt = self.optimize(datasets)
        cov = self.covariance(datasets, optimize_result=opt)
        return FitResult(optimize_result=opt, covariance_result=cov)

    def optimize(self, datasets):
        """fit.py:183-244: autoscale the parameters, minimise the joint stat over the free factors."""
        like = _Likelihood(datasets)
        self._last_like = like
        params = like.datasets.parameters
        params.check_limits()
        if len(like.free) == 0:
            raise ValueError("No free parameters for fitting")
        for p in like.engine.parameters:
            p.autoscale()
        tol = self.optimize_opts.get("tol", 0.1)
        max_iter = self.optimize_opts.get("max_iter", 200)
        x, f, success, message, trace = _minimize(like, tol=tol, max_iter=max_iter, store_trace=self.store_trace)
        params.check_limits()
        return OptimizeResult(parameters=params, total_stat=f, nfev=like.n_calls, success=success, message=message,
                              trace=trace if self.store_trace else None)

    def covariance(self, datasets, optimize_result=None):
        """fit.py:246-300: Hesse in factor space, rescaled to values; errors written to the parameters."""
        like = _Likelihood(datasets)
        x = like.factors
        f0 = like.fcn(x)
        hess = _hesse(like, x, f0)
        scale = np.array([p.scale for p in like.free])
        success, message = True, "Hesse terminated successfully."
        try:
            cov_f = 2.0 * np.linalg.inv(hess)  # errordef = 1 on 2 log L: cov = 2 H^-1
            if not np.all(np.diag(cov_f) > 0):
                raise np.linalg.LinAlgError("non-positive variance")
        except np.linalg.LinAlgError as exc:
            success, message = False, f"Hesse failed: {exc}"
            cov_f = np.full_like(hess, np.nan)
        cov = cov_f * np.outer(scale, scale)
        if success:
            for p, var in zip(like.free, np.diag(cov)):
                p.error = float(np.sqrt(var))
            store_covariance(like.free, cov)  # what Models.covariance assembles (fit.py:172-176, 286-292)
        self._last_like = like
        return CovarianceResult(matrix=cov, success=success, message=message)

    def confidence(self, datasets, parameter, sigma=1, reoptimize=True):
        """fit.py:302-348 with the scipy backend's method (modeling/scipy.py:50-130): on each side of the best fit
        the root of ``stat(x) - stat(best) - sigma**2`` between the best-fit factor and ``conf_min`` / ``conf_max``,
        the other free parameters re-optimised at every trial value when ``reoptimize``.  Returns ``errn``, ``errp``
        (in parameter units), ``success_*``, ``message_*``, ``nfev_*`` (solver iterations) and ``stat_null``."""
        like = _Likelihood(datasets)
        parameter = like.datasets.parameters[parameter] if isinstance(parameter, str) else parameter
        if len(like.free) <= 1:
            reoptimize = False
        opts = dict(self.confidence_opts)
        opts.pop("backend", None)
        rtol = opts.get("rtol", 8.881784197001252e-16)
        xtol = opts.get("xtol", 2e-12)
        maxiter = opts.get("maxiter", 100)
        all_pars = list(like.engine.parameters)
        result = {}
        for upper, suffix in ((False, "errn"), (True, "errp")):
            saved = [(p, p.value, p.frozen) for p in all_pars]
            try:
                stat_null = like.fcn(like.factors)
                x_best = parameter.factor
                bound = (parameter.conf_max if upper else parameter.conf_min) / parameter.scale

                def ts_diff(factor):
                    for p, v, _ in saved:  # every trial starts from the best fit: a failed one cannot spoil the next
                        p.value = v
                    parameter.factor = factor
                    if reoptimize:
                        parameter.frozen = True
                        inner = _Likelihood(like.datasets)
                        _, value, _, _, _ = _minimize(inner, tol=self.optimize_opts.get("tol", 0.1),
                                                      max_iter=self.optimize_opts.get("max_iter", 200))
                    else:
                        value = like.fcn(like.factors)
                    return value - stat_null - sigma ** 2

                root, iterations, _ = _brentq(ts_diff, x_best, bound, xtol=xtol, rtol=rtol, maxiter=maxiter)
                ok = bool(np.isfinite(root))
                result[suffix] = float(np.abs(root - x_best) * parameter.scale)
                result["nfev_" + suffix] = iterations
                result["success_" + suffix] = ok
                result["message_" + suffix] = ("Confidence terminated successfully." if ok else
                                               "Confidence estimation failed. Try to set the parameter.min/max by hand.")
                result["stat_null"] = stat_null
            finally:  # parameters.restore_status()
                for p, v, fr in saved:
                    p.value, p.frozen = v, fr
        result["success"] = result["success_errn"] and result["success_errp"]
        return result

    def stat_profile(self, datasets, parameter, reoptimize=False):
        """fit.py:351-422: the joint stat on ``parameter.scan_values``, all points in one launch."""
        like = _Likelihood(datasets)
        parameter = like.datasets.parameters[parameter] if isinstance(parameter, str) else parameter
        values = np.asarray(parameter.scan_values, dtype=float)
        if reoptimize:
            stats, results = [], []
            saved = [(p, p.value, p.frozen) for p in like.engine.parameters]
            try:
                parameter.frozen = True
                for v in values:
                    parameter.value = v
                    res = self.optimize(like.datasets)
                    stats.append(res.total_stat)
                    results.append(res)
            finally:
                for p, v, fr in saved:
                    p.value, p.frozen = v, fr
            return {f"{parameter.name}_scan": values, "stat_scan": np.array(stats), "fit_results": results}
        base = like.engine.values()
        col = like.engine.parameters.index(parameter)
        grid = np.tile(base, (len(values), 1))
        grid[:, col] = values
        stats = like.stat_values(grid)
        return {f"{parameter.name}_scan": values, "stat_scan": stats, "fit_results": []}

    def stat_surface(self, datasets, x, y, reoptimize=False):
        """fit.py:424-519: joint stat on the (x, y) grid, one launch (chunks of 512 vectors) unless ``reoptimize``."""
        like = _Likelihood(datasets)
        pars = like.datasets.parameters
        x = pars[x] if isinstance(x, str) else x
        y = pars[y] if isinstance(y, str) else y
        xv, yv = np.asarray(x.scan_values, float), np.asarray(y.scan_values, float)
        if reoptimize:  # fit.py:480-500: x and y frozen at each grid point, the other parameters re-fitted
            stats, results = [], []
            saved = [(p, p.value, p.frozen) for p in like.engine.parameters]
            try:
                x.frozen = y.frozen = True
                for a, b in itertools.product(xv, yv):
                    for p, v, _ in saved:
                        if p is not x and p is not y:
                            p.value = v
                    x.value, y.value = a, b
                    res = self.optimize(like.datasets)
                    stats.append(res.total_stat)
                    results.append(res)
            finally:
                for p, v, fr in saved:
                    p.value, p.frozen = v, fr
            shape = (len(xv), len(yv))
            return {f"{x.name}_scan": xv, f"{y.name}_scan": yv, "stat_scan": np.array(stats).reshape(shape),
                    "fit_results": np.array(results, dtype=object).reshape(shape)}
        base = like.engine.values()
        cx, cy = like.engine.parameters.index(x), like.engine.parameters.index(y)
        grid = np.tile(base, (len(xv) * len(yv), 1))
        for k, (a, b) in enumerate(itertools.product(xv, yv)):
            grid[k, cx], grid[k, cy] = a, b
        stats = np.empty(len(grid))
        for i in range(0, len(grid), 512):
            stats[i:i + 512] = like.stat_values(grid[i:i + 512])
        return {f"{x.name}_scan": xv, f"{y.name}_scan": yv, "stat_scan": stats.reshape(len(xv), len(yv)),
                "fit_results": []}
