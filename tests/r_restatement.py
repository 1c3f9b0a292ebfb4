"""Second, independent restatement of the reference's hot path -- a line-by-line transliteration of the R code into
pure Python (small cases only), written from the R sources WITHOUT looking at oracle/hb_oracle.c.

TEST INFRASTRUCTURE ONLY.  Purpose: parity of the C oracle with the reference is "unpinned" (the reference ships no
tests and R is not installed here), i.e. it rests on code reading.  This file is a second reading of
the same R code by a different route (interpreted loops that keep R's statement order, 1-based indices converted at
the array accesses only); tests/test_restatement_crosscheck.py replays one decision tape through both restatements and
requires identical accept/reject sequences and outlier sets and chains within 1e-10.  A disagreement means one of the
two readings of the R code is wrong.

The random draws come from the decision tape in the order the R code consumes them:
  calc_prop            R/internals.R:115-219     u_snooker, D, partner ids, id, zt[d], lambda[d*], gamma choice, zeta[d*]
  metropolis_acceptance R/internals.R:221-247    u_accept
  check_outlier        R/internals.R:71-87       replacement chains
  multi-try selection  R/dream_parallel.R:297-307 mt_pick
Where base R accumulates in long double (sum, colSums, colMeans, mean, sd) numpy.longdouble is used.
"""
import math
import sys

import numpy as np
from scipy import stats

LOG_XMIN = math.log(sys.float_info.min)          # log(.Machine$double.xmin)
LD = np.longdouble


def r_sum(v):
    return float(np.sum(np.asarray(v, dtype=LD)))


def r_mean(v):                                    # R's mean(): long double sum / n, then one correction pass
    v = np.asarray(v, dtype=LD)
    m = np.sum(v) / LD(v.size)
    m = m + np.sum(v - m) / LD(v.size)
    return float(m)


def r_sd(v):                                      # sd(): sqrt(var), two-pass with n - 1
    v = np.asarray(v, dtype=LD)
    n = v.size
    m = np.sum(v) / LD(n)
    m = m + np.sum(v - m) / LD(n)
    return float(np.sqrt(np.sum((v - m) ** 2) / LD(n - 1)))


def quantile7(x, p):                              # stats::quantile, default type 7 (its own branch in quantile.default)
    x = np.sort(np.asarray(x, dtype=np.float64))
    n = x.size
    index = 1 + max(n - 1, 0) * p                 # 1-based
    lo, hi = math.floor(index), math.ceil(index)
    qs = x[lo - 1]
    if index > lo and x[hi - 1] != qs:
        h = index - lo
        qs = (1 - h) * qs + h * x[hi - 1]
    return qs


# ---- targets (R/test_mixture.R:16, R/test_t_distribution.R:19-46, R/test_HydroModel.R:17-36) ----------------------
def mixture(x):
    x = float(np.asarray(x).reshape(-1)[0])
    return 1 / 6 * stats.norm.pdf(x, -8, 1) + 5 / 6 * stats.norm.pdf(x, 10, 1)


_TD = {}


def t_distribution(x):
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    d = x.size
    if d not in _TD:
        A = 0.5 * np.eye(d) + 0.5
        i = np.arange(1, d + 1)
        C = A * np.sqrt(np.outer(i, i))
        _TD[d] = stats.multivariate_t(loc=np.zeros(d), shape=C, df=60)     # mvtnorm::dmvt(x, sigma = C, df = 60, log = TRUE)
    return float(_TD[d].logpdf(x))


def HydroModel(forcing, param):
    if param < 0:
        raise ValueError("Argument 'param' has to be >= zero!")
    Dt, S0 = 1.0, 1.0
    Y = np.empty(len(forcing))
    for i in range(len(forcing)):
        S_Dt = (forcing[i] * Dt + S0) / (1 + param * Dt)
        Y[i] = param * S_Dt
        S0 = S_Dt
    return Y


def abc_distfun(sim, obs):                        # examples/script_examples.R:494
    return np.abs(sim - obs)


def fun_lik(sim, obs):                            # examples/script_examples.R:503-507
    n = len(sim)
    return -n / 2 * math.log(r_sum((sim - obs) ** 2))


# ---- R/internals.R ---------------------------------------------------------------------------------------------------
def calc_ll(res, lik, obs=None, abc_rho=None, abc_e=None, glue_shape=None, lik_fun=None):
    res = np.atleast_1d(np.asarray(res, dtype=np.float64))
    with np.errstate(divide="ignore", invalid="ignore"):
        if lik == 1:
            out = math.log(res[0]) if res[0] > 0 else (-math.inf if res[0] == 0 else math.nan)
        elif lik == 2:
            out = float(res[0])
        elif lik == 21:
            m = len(res)
            abc_e = np.atleast_1d(np.asarray(abc_e, dtype=np.float64))
            out = -m / 2 * math.log(2 * math.pi) - r_sum(np.log(abc_e)) - 0.5 * r_sum((abc_rho(res, obs) / abc_e) ** 2)
        elif lik == 22:
            out = float(np.min(np.atleast_1d(abc_e) - abc_rho(res, obs)))
        elif lik == 31:
            frac = 1 - r_sum((res - obs) ** 2) / r_sum((obs - r_mean(obs)) ** 2)
            v = max(frac, 0.0)
            out = glue_shape * (math.log(v) if v > 0 else -math.inf)
        elif lik == 99:
            out = lik_fun(res, obs)
        else:
            raise ValueError("Argument 'lik' has to be one of {1,2,21,22}!")
    out = max(out, LOG_XMIN)
    if not math.isfinite(out):
        raise FloatingPointError("Calculated log-likelihood is not finite!")
    return out


def prior_pdf(x, par_info, lik):
    if lik == 22:
        lp = 0.0
    else:
        prior = par_info.get("prior", "flat")
        if prior == "flat":
            lp = 0.0
        elif prior == "uniform":
            mn = np.broadcast_to(np.asarray(par_info["min"], dtype=np.float64), x.shape)
            mx = np.broadcast_to(np.asarray(par_info["max"], dtype=np.float64), x.shape)
            terms = [(-math.log(b - a) if a <= v <= b else -math.inf) for v, a, b in zip(x, mn, mx)]   # dunif(log = T)
            lp = -math.inf if any(t == -math.inf for t in terms) else r_sum(terms)
        elif prior == "normal":
            lp = float(stats.multivariate_normal(mean=np.atleast_1d(par_info["mu"]),
                                                 cov=np.atleast_2d(par_info["cov"])).logpdf(x))
        else:
            raise ValueError("user prior functions are not restated")
    lp = max(lp, LOG_XMIN)
    if not math.isfinite(lp):
        raise FloatingPointError("Calculated log-prior is not finite")
    return lp


def check_outlier(dens, x, N, draw_new_states):
    """dens [w, N] (rows ceiling(i/2)..i of lpost); returns (x, last row of dens, outliers 0-based ascending)."""
    proxy = np.array([float(np.sum(dens[:, c].astype(LD)) / LD(dens.shape[0])) for c in range(N)])   # colMeans: long double sum / n, no second pass
    q1, q3 = quantile7(proxy, 0.25), quantile7(proxy, 0.75)
    iqr = abs(q3 - q1)
    outliers = np.nonzero(proxy < q1 - 2 * iqr)[0]
    if len(outliers) > 0:
        new_states = draw_new_states(len(outliers))                         # sample((1:N)[-outliers], n_out), 0-based
        x[outliers, :] = x[new_states, :]
        dens[-1, outliers] = dens[-1, new_states]
    return x, dens[-1, :].copy(), outliers


def bound_par(par, mn, mx, handle):
    par_out = par
    if mn is not None and mx is not None and handle is not None:
        if par < mn or par > mx:
            if handle == "bound":
                par_out = max(min(par, mx), mn)
            elif handle == "reflect":
                par_out = par
                while par_out < mn or par_out > mx:
                    if par_out < mn:
                        par_out = mn + abs(mn - par_out)
                    if par_out > mx:
                        par_out = mx - abs(mx - par_out)
            elif handle == "fold":
                par_out = par
                while par_out < mn or par_out > mx:
                    if par_out < mn:
                        par_out = mx - abs(mn - par_out)
                    if par_out > mx:
                        par_out = mn + abs(mx - par_out)
            else:
                raise ValueError("Value 'bound' of argument list 'par.info' must be one of {'bound', 'reflect', 'fold'}!")
    return par_out


def orth_proj(a, b, p):
    if np.all(a == b):
        q = p
    else:
        u = a - b
        q = a + u * r_sum((p - a) * u) / r_sum(u * u)
    if not np.all(np.isfinite(q)):
        raise FloatingPointError("Result of orthogonal projection is not finite")
    return q


class Draws:
    """The draws of one calc_prop call, in the order R consumes them (one tape record)."""

    def __init__(self, tape, row, ch):
        self.t, self.row, self.ch = tape, row, ch

    def u_snooker(self):
        return 1.0 if self.t.u_snooker is None else float(self.t.u_snooker[self.row, self.ch])

    def D(self):
        return int(self.t.D[self.row, self.ch])

    def partners(self, n):
        return [int(v) for v in self.t.partners[self.row, self.ch, :n]]

    def id(self):
        return int(self.t.id[self.row, self.ch])

    def zt(self):
        return self.t.zt[self.row, self.ch, :]

    def lam(self, n):
        return self.t.lambda_[self.row, self.ch, :n]

    def g_is_one(self):
        return bool(self.t.g_is_one[self.row, self.ch])

    def zeta(self, n):
        return self.t.zeta[self.row, self.ch, :n]

    def g_snooker(self):
        return float(self.t.g_snooker[self.row, self.ch])


def calc_prop(j, x, d, nc, delta, CR, nCR, c_val, c_star, p_g, beta0, par_info, past_sample, z, psnooker, dr):
    """j 0-based.  Returns (xp, id 0-based)."""
    snooker = dr.u_snooker() < psnooker                                    # :119
    D = dr.D()                                                             # :123
    if past_sample:
        if snooker:
            s = dr.partners(3)                                             # :127
            r1, r2, r3 = z[s[0]], z[s[1]], z[s[2]]
        else:
            s = dr.partners(2 * D)                                         # :134
            r1, r2 = z[s[:D]], z[s[D:2 * D]]
    else:
        s = dr.partners(2 * D)                                             # :142  sample((1:nc)[-j], 2D)
        assert j not in s and len(set(s)) == 2 * D
        r1, r2 = x[s[:D]], x[s[D:2 * D]]

    if past_sample and snooker:
        id_ = dr.id()                                                      # :153
        g = dr.g_snooker()                                                 # :156
        zeta = dr.zeta(d)                                                  # :158 rnorm(d, sd = c_star) -- the tape holds the scaled value
        rp2 = orth_proj(x[j], r1, r2)
        rp3 = orth_proj(x[j], r1, r3)
        dx = zeta + g * (rp2 - rp3)                                        # :163
    else:
        dx = np.zeros(d)
        id_ = dr.id()                                                      # :173
        zt = dr.zt()                                                       # :175
        A = np.nonzero(zt < CR[id_])[0]
        d_star = len(A)
        if d_star == 0:
            A = np.array([int(np.argmin(zt))])
            d_star = 1
        lam = dr.lam(d_star)                                               # :188
        gamma_d = 2.38 / math.sqrt(2 * D * d_star)
        g = 1.0 if dr.g_is_one() else gamma_d                              # :192
        zeta = dr.zeta(d_star)                                             # :194
        jump_diff = np.array([r_sum(r1[:, k] - r2[:, k]) for k in A])     # colSums(r1[,A] - r2[,A])
        dx[A] = zeta + (1 + lam) * g * jump_diff
        dx[A] = dx[A] * beta0

    xp = x[j] + dx
    if not np.all(np.isfinite(xp)):
        raise FloatingPointError("Calculated proposal is not finite!")
    if par_info.get("min") is not None and par_info.get("max") is not None:
        mn = np.broadcast_to(np.asarray(par_info["min"], dtype=np.float64), (d,))
        mx = np.broadcast_to(np.asarray(par_info["max"], dtype=np.float64), (d,))
        xp = np.array([bound_par(xp[k], mn[k], mx[k], par_info.get("bound")) for k in range(d)])
    return xp, id_


def metropolis_acceptance(p_xp, p_x, lik, mt, p_zp=None, u=None):
    if lik == 22:
        p_acc = max(1 if p_xp >= p_x else 0, 1 if p_xp >= 0 else 0)
        return p_acc == 1
    if mt > 1:
        p1 = r_sum(np.exp(p_xp))
        p2 = r_sum(np.exp(p_zp))
        if p2 == 0:
            return False
        return min(1.0, p1 / p2) > u
    p_acc = min(max(-100.0, p_xp - p_x), 0.0)
    return p_acc > math.log(u)


# ---- R/gelman_rubin.R:21-56 -------------------------------------------------------------------------------------------
def R_stat(x):
    t, d, n = x.shape
    lo = math.ceil(t / 2) - 1                      # rows ceiling(t/2):t, 0-based start
    out = np.empty(d)
    for k in range(d):
        y = x[lo:, k, :]                           # [rows, n]
        x_mean = np.array([2 / (t - 2) * r_sum(y[:, c]) for c in range(n)])
        W = 2 / (n * (t - 2)) * r_sum((y - x_mean[None, :]) ** 2)
        x_mm = r_mean(x_mean)
        B = 1 / (2 * (n - 1)) * r_sum((x_mean - x_mm) ** 2)
        sigma = (t - 2) / t * W + 2 * B
        out[k] = math.sqrt((n + 1) / n * sigma / W - (t - 2) / (n * t))
    return out


# ---- drivers ------------------------------------------------------------------------------------------------------------
def _eval(fun, x, dots):
    return np.atleast_1d(fun(x, *dots))


def _out_t(t, burnin, thin):
    return int((1 - burnin) * t / thin + 1) if (burnin > 0 or thin > 1) else t


def dream_parallel(fun, dots, lik, par_info, nc, t, d, tape, z0, burnin=0, adapt=0.1, updateInterval=10, delta=3,
                   c_val=0.1, c_star=1e-12, nCR=3, p_g=0.2, beta0=1, thin=1, outlier_check=True, obs=None, abc_rho=None,
                   abc_e=None, glue_shape=None, lik_fun=None, past_sample=False, m0=None, archive_update=None, psnooker=0,
                   mt=1, checkConvergence=False, run=0):
    """R/dream_parallel.R:155-455.  z0 is what prior_sample(par.info, d, m0) returned (initial = 'user')."""
    if not past_sample and nc <= delta * 2:
        raise ValueError("Argument 'nc' must be > 'delta' * 2!")
    if past_sample and m0 is None:
        m0 = 10 * d
    if not past_sample and m0 is None:
        m0 = nc
    if past_sample and archive_update is None:
        archive_update = 10
    if (mt > 1 or psnooker > 0) and lik in (21, 22):
        raise ValueError("ABC method cannot be combined with MT-DREAM or snooker updating")
    n_sets = 2 * mt - 1 if mt > 1 else 1
    ch0 = run * nc
    J = np.zeros(nCR); n_id = np.zeros(nCR)
    n_acc = 0; n_eval = 0
    CR = np.arange(1, nCR + 1) / nCR
    pCR = np.ones(nCR) / nCR
    out_t = _out_t(t, burnin, thin)
    chain = np.full((out_t, d + 3, nc), np.nan)
    n_ar = t // updateInterval + 1
    out_CR = np.full((n_ar, nCR + 1), np.nan)
    out_AR = np.full((n_ar, 2), np.nan)
    out_rstat = np.full((max(out_t - 50, 0), d), np.nan) if checkConvergence else None
    outl = []

    z = np.array(z0, dtype=np.float64).reshape(m0, d)
    xt = z[m0 - nc:m0].copy()
    lp = np.array([prior_pdf(xt[c], par_info, lik) for c in range(nc)])
    res_fun = np.array([_eval(fun, xt[c], dots) for c in range(nc)])
    fx = np.full((out_t, nc, res_fun.shape[1]), np.nan)
    ll = np.array([calc_ll(res_fun[c], lik, obs, abc_rho, abc_e, glue_shape, lik_fun) for c in range(nc)])
    lpost = np.full((t, nc), np.nan)
    lpost[0] = lp + ll
    chain[0, :d, :] = xt.T; chain[0, d, :] = lp; chain[0, d + 1, :] = ll; chain[0, d + 2, :] = lpost[0]
    out_AR[0, 0] = out_CR[0, 0] = nc
    out_CR[0, 1:] = pCR
    fx[0] = res_fun
    ind_out = 1
    i_ar = 1
    accept_log = np.zeros((t - 1, nc), np.uint8)
    J_adapt = J.copy()
    floored = 0

    def props(x_around, set0, n_rep):
        xs, ids = [], []
        for m in range(n_rep):                                              # replicate(mt, lapply(1:nc, ...)): try-major
            for j in range(nc):
                dr = Draws(tape, (i - 2) * n_sets + set0 + m, ch0 + j)
                xp_j, id_j = calc_prop(j, x_around, d, nc, delta, CR, nCR, c_val, c_star, p_g, beta0, par_info, past_sample, z,
                                       psnooker, dr)
                xs.append(xp_j); ids.append(id_j)
        return np.array(xs), np.array(ids)

    for i in range(2, t + 1):
        if i > burnin * t and i % thin == 0:
            ind_out += 1
        dx = np.zeros((nc, d))
        xp, id_ = props(xt, 0, mt)
        lp_xp = np.array([prior_pdf(r, par_info, lik) for r in xp])
        floored += int((lp_xp == LOG_XMIN).sum())
        res_fun_t = np.array([_eval(fun, r, dots) for r in xp])
        ll_xp = np.array([calc_ll(r, lik, obs, abc_rho, abc_e, glue_shape, lik_fun) for r in res_fun_t])
        lpost_xp = lp_xp + ll_xp

        if mt > 1:
            samp_ind = np.array([int(tape.mt_pick[i - 2, ch0 + j]) for j in range(nc)])     # sample(1:mt, 1, prob = p) :306
            samp_ind = samp_ind * nc + np.arange(nc)
            zt = xp[samp_ind]; ll_zt = ll_xp[samp_ind]; lp_zt = lp_xp[samp_ind]; lpost_zt = lpost_xp[samp_ind]
            zt_resfun = res_fun_t[samp_ind]
            id_ = id_[samp_ind]
            zp, _ = props(zt, mt, mt - 1)
            lp_zp = np.array([prior_pdf(r, par_info, lik) for r in zp])
            res_fun_zp = np.array([_eval(fun, r, dots) for r in zp])
            ll_zp = np.array([calc_ll(r, lik, obs, abc_rho, abc_e, glue_shape, lik_fun) for r in res_fun_zp])
            lpost_zp = np.concatenate([lp_zp + ll_zp, lpost[i - 2]])
            grp = [np.arange(mt) * nc + j for j in range(nc)]
            accept = np.array([metropolis_acceptance(lpost_xp[grp[j]], None, lik, mt, lpost_zp[grp[j]],
                                                     float(tape.u_accept[i - 2, ch0 + j])) for j in range(nc)])
        else:
            accept = np.array([metropolis_acceptance(lpost_xp[j], lpost[i - 2, j], lik, 1, None,
                                                     float(tape.u_accept[i - 2, ch0 + j])) for j in range(nc)])
        accept_log[i - 2] = accept
        if accept.any():
            src_x, src_lp, src_ll, src_lpost, src_res = (zt, lp_zt, ll_zt, lpost_zt, zt_resfun) if mt > 1 else \
                (xp, lp_xp, ll_xp, lpost_xp, res_fun_t)
            dx[accept] = src_x[accept] - xt[accept]
            xt[accept] = src_x[accept]
            lp[accept] = src_lp[accept]; ll[accept] = src_ll[accept]
            lpost[i - 1, accept] = src_lpost[accept]
            res_fun[accept] = src_res[accept]
        if (~accept).any():
            lpost[i - 1, ~accept] = lpost[i - 2, ~accept]

        for q in id_:
            n_id[q] += 1
        n_acc += int(accept.sum())
        n_eval += len(accept)
        std_x = np.array([r_sd(xt[:, k]) for k in range(d)])
        with np.errstate(divide="ignore", invalid="ignore"):
            jump_dist = np.array([r_sum((dx[j] / std_x) ** 2) for j in range(nc)])
        for j in range(nc):
            J[id_[j]] += jump_dist[j]
        if i <= adapt * t:
            J_adapt = J.copy()                                             # J is only ever read inside the adaptation period

        if archive_update is not None and i % archive_update == 0:
            z = np.vstack([z, xt])

        if i > burnin * t and i % thin == 0:
            r = ind_out - 1
            chain[r, :d, :] = xt.T; chain[r, d, :] = lp; chain[r, d + 1, :] = ll; chain[r, d + 2, :] = lpost[i - 1]
            fx[r] = res_fun
            if checkConvergence and ind_out > 50:                         # out_x := chain[, 1:d, ]
                out_rstat[ind_out - 51] = R_stat(chain[:ind_out, :d, :])

        if i <= adapt * t and i % updateInterval == 0:
            if (J > 0).any():
                with np.errstate(divide="ignore", invalid="ignore"):
                    pCR = J / n_id
                pCR[np.isnan(pCR)] = 1 / nCR
                pCR = pCR / r_sum(pCR)
            if not past_sample and outlier_check:
                e = i // updateInterval - 1
                w0 = math.ceil(i / 2) - 1

                def draw(n_out, e=e):
                    return np.array([int(v) for v in tape.outlier_new[e, ch0:ch0 + n_out]])
                dens = lpost[w0:i].copy()
                xt, last, outliers = check_outlier(dens, xt, nc, draw)
                lpost[i - 1] = last
                outl += [(i, int(c)) for c in outliers]

        if i % updateInterval == 0:
            i_ar += 1
            out_AR[i_ar - 1, 1] = n_acc / n_eval
            out_AR[i_ar - 1, 0] = out_CR[i_ar - 1, 0] = out_CR[i_ar - 2, 0] + n_eval
            n_acc = 0; n_eval = 0
            out_CR[i_ar - 1, 1:] = pCR

    return dict(chain=chain, fx=fx, outlier=outl, AR=out_AR, CR=out_CR, R_stat=out_rstat, accept=accept_log,
                xt=xt, lp=lp, ll=ll, lpost=lpost[t - 1], pCR=pCR, J=J, J_adapt=J_adapt, n_id=n_id, floored=floored)


def dream(fun, dots, lik, par_info, nc, t, d, tape, x0, burnin=0, adapt=0.1, updateInterval=10, delta=3, c_val=0.1,
          c_star=1e-12, nCR=3, p_g=0.2, beta0=1, thin=1, outlier_check=True, obs=None, abc_rho=None, abc_e=None,
          glue_shape=None, lik_fun=None, checkConvergence=False, run=0):
    """R/dream.R:126-316 (sequential sweep)."""
    if nc <= delta * 2:
        raise ValueError("Argument 'nc' must be > 'delta' * 2!")
    ch0 = run * nc
    J = np.zeros(nCR); n_id = np.zeros(nCR); n_acc = np.zeros(nCR)
    CR = np.arange(1, nCR + 1) / nCR
    pCR = np.ones(nCR) / nCR
    out_t = _out_t(t, burnin, thin)
    chain = np.full((out_t, d + 3, nc), np.nan)
    out_AR = np.full((out_t, nCR), np.nan); out_CR = np.full((out_t, nCR), np.nan)
    out_rstat = np.full((max(out_t - 50, 0), d), np.nan) if checkConvergence else None
    outl = []
    xt = np.array(x0, dtype=np.float64).reshape(nc, d).copy()
    lp = np.array([prior_pdf(xt[c], par_info, lik) for c in range(nc)])
    res_fun = np.array([_eval(fun, xt[c], dots) for c in range(nc)])
    fx = np.full((out_t, nc, res_fun.shape[1]), np.nan)
    ll = np.array([calc_ll(res_fun[c], lik, obs, abc_rho, abc_e, glue_shape, lik_fun) for c in range(nc)])
    lpost = np.full((t, nc), np.nan)
    lpost[0] = lp + ll
    chain[0, :d, :] = xt.T; chain[0, d, :] = lp; chain[0, d + 1, :] = ll; chain[0, d + 2, :] = lpost[0]
    out_AR[0] = 0                                                         # rep(0, 3) at :193, SURVEY Appendix C: rep(0, nCR)
    out_CR[0] = pCR
    fx[0] = res_fun
    ind_out = 1
    accept_log = np.zeros((t - 1, nc), np.uint8)
    J_adapt = J.copy()

    for i in range(2, t + 1):
        if i > burnin * t and i % thin == 0:
            ind_out += 1
        for j in range(nc):
            dr = Draws(tape, i - 2, ch0 + j)
            xp, id_t = calc_prop(j, xt, d, nc, delta, CR, nCR, c_val, c_star, p_g, beta0, par_info, False, None, 0, dr)
            lp_xp = prior_pdf(xp, par_info, lik)
            res_fun_t = _eval(fun, xp, dots)
            ll_xp = calc_ll(res_fun_t, lik, obs, abc_rho, abc_e, glue_shape, lik_fun)
            lpost_xp = lp_xp + ll_xp
            accept = metropolis_acceptance(lpost_xp, lpost[i - 2, j], lik, 1, None, float(tape.u_accept[i - 2, ch0 + j]))
            accept_log[i - 2, j] = accept
            if accept:
                dx = xp - xt[j]
                xt[j] = xp
                lp[j] = lp_xp; ll[j] = ll_xp
                lpost[i - 1, j] = lpost_xp
                res_fun[j] = res_fun_t
                n_acc[id_t] += 1
            else:
                lpost[i - 1, j] = lpost[i - 2, j]
                dx = np.zeros(d)
            n_id[id_t] += 1
            std_x = np.array([r_sd(xt[:, k]) for k in range(d)])
            with np.errstate(divide="ignore", invalid="ignore"):
                J[id_t] += r_sum((dx / std_x) ** 2)
        if i <= adapt * t:
            J_adapt = J.copy()

        if i > burnin * t and i % thin == 0:
            r = ind_out - 1
            chain[r, :d, :] = xt.T; chain[r, d, :] = lp; chain[r, d + 1, :] = ll; chain[r, d + 2, :] = lpost[i - 1]
            fx[r] = res_fun
            with np.errstate(divide="ignore", invalid="ignore"):
                out_AR[r] = n_acc / n_id
            out_CR[r] = pCR
            if checkConvergence and ind_out > 50:
                out_rstat[ind_out - 51] = R_stat(chain[:ind_out, :d, :])

        if i <= adapt * t and i % updateInterval == 0:
            if (J > 0).any():
                with np.errstate(divide="ignore", invalid="ignore"):
                    pCR = J / n_id
                pCR[np.isnan(pCR)] = 1 / nCR
                pCR = pCR / r_sum(pCR)
            if outlier_check:
                e = i // updateInterval - 1
                w0 = math.ceil(i / 2) - 1

                def draw(n_out, e=e):
                    return np.array([int(v) for v in tape.outlier_new[e, ch0:ch0 + n_out]])
                dens = lpost[w0:i].copy()
                xt, last, outliers = check_outlier(dens, xt, nc, draw)
                lpost[i - 1] = last
                outl += [(i, int(c)) for c in outliers]

    return dict(chain=chain, fx=fx, outlier=outl, AR=out_AR, CR=out_CR, R_stat=out_rstat, accept=accept_log,
                xt=xt, lp=lp, ll=ll, lpost=lpost[t - 1], pCR=pCR, J=J, J_adapt=J_adapt, n_id=n_id)
