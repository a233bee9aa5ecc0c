"""Restatement of mpfit (MINPACK-1 ``lmdif`` + Markwardt's box limits).  TEST INFRASTRUCTURE.

The reference reaches this solver through pyspeckit:
    multi_v_fit.py:393 / :206 / :468  ``pcube.fiteach(...)``
      -> pyspeckit ``Cube.fiteach`` -> ``Specfit.multifit`` -> ``SpectralModel.fitter`` -> ``mpfit.mpfit``
pyspeckit (pin ``pyspeckit>=1.0.4``, pyproject.toml:22) is NOT vendored in /root/reference and cannot be
installed here, so this file restates the published algorithm:
  * MINPACK-1 (More', Garbow, Hillstrom 1980): lmdif, qrfac, qrsolv, lmpar, fdjac2, covar
  * C. Markwardt's MPFIT extensions as ported to Python (Rivers / Koposov), in the form pyspeckit ships
    (``pyspeckit/mpfit/mpfit.py``): two-sided parameter limits with "pegged" parameters whose Jacobian
    column is zeroed when the gradient points out of the box, step truncation by ``alpha``,
    snapping to the limit, and ``perror = sqrt(diag(covar))`` with zero for pegged parameters.
Defaults are mpfit's: ftol = xtol = gtol = 1e-10, maxiter = 200, factor = 100, forward differences with
relative step sqrt(machep).  [RECALLED -- PARITY UNPINNED: no golden vector from the reference exists for
the solver.  Cross-checked against scipy's real MINPACK in tests/test_oracle_mpfit.py: same minimum as lmdif on
interior solutions (<= 0.05 sigma, chi2 to 1e-6), same chi2 as the bounded trf solver where a parameter is pegged]

Supported subset: every parameter free, optional lower/upper limits per parameter, no ties, no
``mpmaxstep``, one-sided automatic derivatives -- exactly what SpectralModel.fitter builds for MUFASA
(BaseModel.py:69-80 + the ``limitedmin/limitedmax/minpars/maxpars`` kwargs of multi_v_fit.py:374-383).
"""
import numpy as np

MACHEP = np.finfo(np.float64).eps          # mpfit machar.machep
DWARF = np.finfo(np.float64).tiny          # mpfit machar.minnum


def enorm(vec):
    return np.sqrt(np.sum(vec * vec))


class MPFitResult(object):
    __slots__ = ("params", "perror", "covar", "fnorm", "status", "niter", "nfev", "errmsg")

    def __init__(self):
        self.params = None
        self.perror = None
        self.covar = None
        self.fnorm = None
        self.status = 0
        self.niter = 0
        self.nfev = 0
        self.errmsg = ""


def qrfac(a, pivot=True):
    """Householder QR with column pivoting (MINPACK qrfac, mpfit layout)."""
    m, n = a.shape
    acnorm = np.zeros(n)
    for j in range(n):
        acnorm[j] = enorm(a[:, j])
    rdiag = acnorm.copy()
    wa = rdiag.copy()
    ipvt = np.arange(n)
    minmn = min(m, n)
    for j in range(minmn):
        if pivot:
            rmax = np.max(rdiag[j:])
            kmax = np.nonzero(rdiag[j:] == rmax)[0]
            if len(kmax) > 0:
                kmax = kmax[0] + j
                if kmax != j:
                    ipvt[j], ipvt[kmax] = ipvt[kmax], ipvt[j]
                    rdiag[kmax] = rdiag[j]
                    wa[kmax] = wa[j]
        lj = ipvt[j]
        ajj = a[j:, lj]
        ajnorm = enorm(ajj)
        if ajnorm == 0:
            break
        if a[j, lj] < 0:
            ajnorm = -ajnorm
        ajj = ajj / ajnorm
        ajj[0] = ajj[0] + 1
        a[j:, lj] = ajj
        if j + 1 < n:
            for k in range(j + 1, n):
                lk = ipvt[k]
                ajk = a[j:, lk]
                if a[j, lj] != 0:
                    a[j:, lk] = ajk - ajj * np.sum(ajk * ajj) / a[j, lj]
                    if pivot and rdiag[k] != 0:
                        temp = a[j, lk] / rdiag[k]
                        rdiag[k] = rdiag[k] * np.sqrt(max(1. - temp ** 2, 0.))
                        temp = rdiag[k] / wa[k]
                        if (0.05 * temp * temp) <= MACHEP:
                            rdiag[k] = enorm(a[j + 1:, lk])
                            wa[k] = rdiag[k]
        rdiag[j] = -ajnorm
    return a, ipvt, rdiag, acnorm


def qrsolv(r, ipvt, diag, qtb, sdiag):
    n = r.shape[1]
    for j in range(n):
        r[j:n, j] = r[j, j:n]
    x = np.diagonal(r).copy()
    wa = qtb.copy()
    for j in range(n):
        l = ipvt[j]
        if diag[l] == 0:
            break
        sdiag[j:] = 0
        sdiag[j] = diag[l]
        qtbpj = 0.
        for k in range(j, n):
            if sdiag[k] == 0:
                break
            if np.abs(r[k, k]) < np.abs(sdiag[k]):
                cotan = r[k, k] / sdiag[k]
                sine = 0.5 / np.sqrt(.25 + .25 * cotan * cotan)
                cosine = sine * cotan
            else:
                tang = sdiag[k] / r[k, k]
                cosine = 0.5 / np.sqrt(.25 + .25 * tang * tang)
                sine = cosine * tang
            r[k, k] = cosine * r[k, k] + sine * sdiag[k]
            temp = cosine * wa[k] + sine * qtbpj
            qtbpj = -sine * wa[k] + cosine * qtbpj
            wa[k] = temp
            if n > k + 1:
                temp = cosine * r[k + 1:n, k] + sine * sdiag[k + 1:n]
                sdiag[k + 1:n] = -sine * r[k + 1:n, k] + cosine * sdiag[k + 1:n]
                r[k + 1:n, k] = temp
        sdiag[j] = r[j, j]
        r[j, j] = x[j]
    nsing = n
    wh = np.nonzero(sdiag == 0)[0]
    if len(wh) > 0:
        nsing = wh[0]
        wa[nsing:] = 0
    if nsing >= 1:
        wa[nsing - 1] = wa[nsing - 1] / sdiag[nsing - 1]
        for j in range(nsing - 2, -1, -1):
            sum0 = np.sum(r[j + 1:nsing, j] * wa[j + 1:nsing])
            wa[j] = (wa[j] - sum0) / sdiag[j]
    x[ipvt] = wa
    return r, x, sdiag


def lmpar(r, ipvt, diag, qtb, delta, x, sdiag, par=0.):
    n = r.shape[1]
    nsing = n
    wa1 = qtb.copy()
    rdiag = np.abs(np.diagonal(r))
    rthresh = np.max(rdiag) * MACHEP
    wh = np.nonzero(rdiag < rthresh)[0]
    if len(wh) > 0:
        nsing = wh[0]
        wa1[wh[0]:] = 0
    if nsing >= 1:
        for j in range(nsing - 1, -1, -1):
            wa1[j] = wa1[j] / r[j, j]
            if j - 1 >= 0:
                wa1[0:j] = wa1[0:j] - r[0:j, j] * wa1[j]
    x[ipvt] = wa1
    it = 0
    wa2 = diag * x
    dxnorm = enorm(wa2)
    fp = dxnorm - delta
    if fp <= 0.1 * delta:
        return r, 0., x, sdiag
    parl = 0.
    if nsing >= n:
        wa1 = diag[ipvt] * wa2[ipvt] / dxnorm
        wa1[0] = wa1[0] / r[0, 0]
        for j in range(1, n):
            sum0 = np.sum(r[0:j, j] * wa1[0:j])
            wa1[j] = (wa1[j] - sum0) / r[j, j]
        temp = enorm(wa1)
        parl = ((fp / delta) / temp) / temp
    for j in range(n):
        sum0 = np.sum(r[0:j + 1, j] * qtb[0:j + 1])
        wa1[j] = sum0 / diag[ipvt[j]]
    gnorm = enorm(wa1)
    paru = gnorm / delta
    if paru == 0:
        paru = DWARF / min(delta, 0.1)
    par = max(par, parl)
    par = min(par, paru)
    if par == 0:
        par = gnorm / dxnorm
    while True:
        it += 1
        if par == 0:
            par = max(DWARF, paru * 0.001)
        temp = np.sqrt(par)
        wa1 = temp * diag
        r, x, sdiag = qrsolv(r, ipvt, wa1, qtb, sdiag)
        wa2 = diag * x
        dxnorm = enorm(wa2)
        temp = fp
        fp = dxnorm - delta
        if (np.abs(fp) <= 0.1 * delta) or ((parl == 0) and (fp <= temp) and (temp < 0)) or (it == 10):
            break
        wa1 = diag[ipvt] * wa2[ipvt] / dxnorm
        for j in range(n - 1):
            wa1[j] = wa1[j] / sdiag[j]
            wa1[j + 1:n] = wa1[j + 1:n] - r[j + 1:n, j] * wa1[j]
        wa1[n - 1] = wa1[n - 1] / sdiag[n - 1]
        temp = enorm(wa1)
        parc = ((fp / delta) / temp) / temp
        if fp > 0:
            parl = max(parl, par)
        if fp < 0:
            paru = min(paru, par)
        par = max(parl, par + parc)
    return r, par, x, sdiag


def calc_covar(rr, ipvt, tol=1.e-14):
    n = rr.shape[0]
    r = rr.copy()
    l = -1
    tolr = tol * np.abs(r[0, 0])
    for k in range(n):
        if np.abs(r[k, k]) <= tolr:
            break
        r[k, k] = 1. / r[k, k]
        for j in range(k):
            temp = r[k, k] * r[j, k]
            r[j, k] = 0.
            r[0:j + 1, k] = r[0:j + 1, k] - temp * r[0:j + 1, j]
        l = k
    if l >= 0:
        for k in range(l + 1):
            for j in range(k):
                temp = r[j, k]
                r[0:j + 1, j] = r[0:j + 1, j] + temp * r[0:j + 1, k]
            temp = r[k, k]
            r[0:k + 1, k] = temp * r[0:k + 1, k]
    wa = np.repeat([r[0, 0]], n)
    for j in range(n):
        jj = ipvt[j]
        sing = j > l
        for i in range(j + 1):
            if sing:
                r[i, j] = 0.
            ii = ipvt[i]
            if ii > jj:
                r[ii, jj] = r[i, j]
            if ii < jj:
                r[jj, ii] = r[i, j]
        wa[jj] = r[j, j]
    for j in range(n):
        r[0:j + 1, j] = r[j, 0:j + 1]
        r[j, j] = wa[j]
    return r


def _fdjac2(fcn, x, fvec, qulim, ulim, res):
    """Forward-difference Jacobian (mpfit.fdjac2, one-sided, relative step sqrt(machep))."""
    eps = np.sqrt(MACHEP)
    m, n = len(fvec), len(x)
    fjac = np.zeros([m, n], dtype=float)
    h = eps * np.abs(x)
    h[h == 0] = eps
    mask = qulim & (x > ulim - h)
    h[mask] = -h[mask]
    for j in range(n):
        xp = x.copy()
        xp[j] = xp[j] + h[j]
        fp = fcn(xp)
        res.nfev += 1
        fjac[:, j] = (fp - fvec) / h[j]
    return fjac


def mpfit(fcn, xall, limited=None, limits=None, ftol=1.e-10, xtol=1.e-10, gtol=1.e-10,
          maxiter=200, factor=100.):
    """Minimise sum(fcn(p)**2).  ``fcn(p) -> deviates``.  Returns an MPFitResult.

    status: 0 = bad input (e.g. start outside limits), 1-4 MINPACK convergence, 5 = maxiter,
            6-8 tolerances too small, -16 = non-finite."""
    res = MPFitResult()
    xall = np.asarray(xall, dtype=float).copy()
    n = len(xall)
    if limited is None:
        limited = np.zeros((n, 2), dtype=bool)
        limits = np.zeros((n, 2))
    limited = np.asarray(limited, dtype=bool)
    limits = np.asarray(limits, dtype=float)
    res.params = xall.copy()

    wh = np.nonzero((limited[:, 0] & (xall < limits[:, 0])) | (limited[:, 1] & (xall > limits[:, 1])))[0]
    if len(wh) > 0:
        res.errmsg = 'ERROR: parameters are not within PARINFO limits'
        return res
    wh = np.nonzero(limited[:, 0] & limited[:, 1] & (limits[:, 0] >= limits[:, 1]))[0]
    if len(wh) > 0:
        res.errmsg = 'ERROR: PARINFO parameter limits are not consistent'
        return res

    qulim, ulim = limited[:, 1], limits[:, 1]
    qllim, llim = limited[:, 0], limits[:, 0]
    qanylim = bool(np.any(qulim) or np.any(qllim))

    x = xall.copy()
    fvec = fcn(x)
    res.nfev += 1
    m = len(fvec)
    if m < n:
        res.errmsg = 'ERROR: number of parameters must not exceed data'
        return res
    fnorm = enorm(fvec)
    fnorm1 = None

    par = 0.
    niter = 1
    qtf = x * 0.
    status = 0
    diag = None
    delta = 0.
    xnorm = 0.
    fjac = None
    ipvt = None
    nlpeg = nupeg = 0
    whlpeg = whupeg = None

    while True:
        res.params = x.copy()
        status = 2
        fjac = _fdjac2(fcn, x, fvec, qulim, ulim, res)

        if qanylim:
            whlpeg = np.nonzero(qllim & (x == llim))[0]
            nlpeg = len(whlpeg)
            whupeg = np.nonzero(qulim & (x == ulim))[0]
            nupeg = len(whupeg)
            for i in range(nlpeg):
                sum0 = np.sum(fvec * fjac[:, whlpeg[i]])
                if sum0 > 0:
                    fjac[:, whlpeg[i]] = 0
            for i in range(nupeg):
                sum0 = np.sum(fvec * fjac[:, whupeg[i]])
                if sum0 < 0:
                    fjac[:, whupeg[i]] = 0

        fjac, ipvt, wa1, wa2 = qrfac(fjac, pivot=True)

        if niter == 1:
            diag = wa2.copy()
            diag[diag == 0] = 1.
            wa3 = diag * x
            xnorm = enorm(wa3)
            delta = factor * xnorm
            if delta == 0.:
                delta = factor

        wa4 = fvec.copy()
        for j in range(n):
            lj = ipvt[j]
            temp3 = fjac[j, lj]
            if temp3 != 0:
                fj = fjac[j:, lj]
                wj = wa4[j:]
                wa4[j:] = wj - fj * np.sum(fj * wj) / temp3
            fjac[j, lj] = wa1[j]
            qtf[j] = wa4[j]
        fjac = fjac[0:n, 0:n]
        fjac = fjac[:, ipvt].copy()

        if not np.all(np.isfinite(fjac)):
            res.errmsg = 'ERROR: parameter or function value(s) have become infinite'
            status = -16
            break

        gnorm = 0.
        if fnorm != 0:
            for j in range(n):
                l = ipvt[j]
                if wa2[l] != 0:
                    sum0 = np.sum(fjac[0:j + 1, j] * qtf[0:j + 1]) / fnorm
                    gnorm = max(gnorm, np.abs(sum0 / wa2[l]))
        if gnorm <= gtol:
            status = 4
            break
        if maxiter == 0:
            status = 5
            break

        diag = np.where(diag > wa2, diag, wa2)

        while True:
            fjac, par, wa1, wa2 = lmpar(fjac, ipvt, diag, qtf, delta, wa1, wa2, par=par)
            wa1 = -wa1
            if not qanylim:
                alpha = 1.
                wa2 = x + wa1
            else:
                alpha = 1.
                if nlpeg > 0:
                    wa1[whlpeg] = np.clip(wa1[whlpeg], 0., np.max(wa1))
                if nupeg > 0:
                    wa1[whupeg] = np.clip(wa1[whupeg], np.min(wa1), 0.)
                dwa1 = np.abs(wa1) > MACHEP
                whl = np.nonzero((dwa1 & qllim) & ((x + wa1) < llim))[0]
                if len(whl) > 0:
                    t = (llim[whl] - x[whl]) / wa1[whl]
                    alpha = min(alpha, np.min(t))
                whu = np.nonzero((dwa1 & qulim) & ((x + wa1) > ulim))[0]
                if len(whu) > 0:
                    t = (ulim[whu] - x[whu]) / wa1[whu]
                    alpha = min(alpha, np.min(t))
                wa1 = wa1 * alpha
                wa2 = x + wa1
                sgnu = (ulim >= 0) * 2. - 1.
                sgnl = (llim >= 0) * 2. - 1.
                ulim1 = ulim * (1 - sgnu * MACHEP) - (ulim == 0) * MACHEP
                llim1 = llim * (1 + sgnl * MACHEP) + (llim == 0) * MACHEP
                wh = np.nonzero(qulim & (wa2 >= ulim1))[0]
                if len(wh) > 0:
                    wa2[wh] = ulim[wh]
                wh = np.nonzero(qllim & (wa2 <= llim1))[0]
                if len(wh) > 0:
                    wa2[wh] = llim[wh]
            wa3 = diag * wa1
            pnorm = enorm(wa3)
            if niter == 1:
                delta = min(delta, pnorm)

            res.params = wa2.copy()
            wa4 = fcn(wa2)
            res.nfev += 1
            fnorm1 = enorm(wa4)
            if not np.isfinite(fnorm1):
                res.errmsg = 'ERROR: parameter or function value(s) have become infinite'
                status = -16
                break

            actred = -1.
            if (0.1 * fnorm1) < fnorm:
                actred = -(fnorm1 / fnorm) ** 2 + 1.

            for j in range(n):
                wa3[j] = 0
                wa3[0:j + 1] = wa3[0:j + 1] + fjac[0:j + 1, j] * wa1[ipvt[j]]
            temp1 = enorm(alpha * wa3) / fnorm
            temp2 = (np.sqrt(alpha * par) * pnorm) / fnorm
            prered = temp1 * temp1 + (temp2 * temp2) / 0.5
            dirder = -(temp1 * temp1 + temp2 * temp2)

            ratio = 0.
            if prered != 0:
                ratio = actred / prered

            if ratio <= 0.25:
                if actred >= 0:
                    temp = .5
                else:
                    temp = .5 * dirder / (dirder + .5 * actred)
                if ((0.1 * fnorm1) >= fnorm) or (temp < 0.1):
                    temp = 0.1
                delta = temp * min(delta, pnorm / 0.1)
                par = par / temp
            else:
                if (par == 0) or (ratio >= 0.75):
                    delta = pnorm / .5
                    par = .5 * par

            if ratio >= 0.0001:
                x = wa2
                wa2 = diag * x
                fvec = wa4
                xnorm = enorm(wa2)
                fnorm = fnorm1
                niter += 1

            status = 0
            if (np.abs(actred) <= ftol) and (prered <= ftol) and (0.5 * ratio <= 1):
                status = 1
            if delta <= xtol * xnorm:
                status = 2
            if (np.abs(actred) <= ftol) and (prered <= ftol) and (0.5 * ratio <= 1) and (status == 2):
                status = 3
            if status != 0:
                break
            if niter >= maxiter:
                status = 5
            if (np.abs(actred) <= MACHEP) and (prered <= MACHEP) and (0.5 * ratio <= 1):
                status = 6
            if delta <= MACHEP * xnorm:
                status = 7
            if gnorm <= MACHEP:
                status = 8
            if status != 0:
                break
            if ratio >= 0.0001:
                break
        if status != 0:
            break

    res.status = status
    res.niter = niter
    res.params = x.copy()
    if status > 0:
        # mpfit (nprint>0) re-evaluates at the final point, then reports max(fnorm, fnorm1)**2
        fvec = fcn(x)
        res.nfev += 1
        fnorm = enorm(fvec)
    if fnorm is not None and fnorm1 is not None:
        fnorm = max(fnorm, fnorm1)
    res.fnorm = fnorm ** 2
    if status > 0 and fjac is not None and ipvt is not None:
        cv = calc_covar(fjac[0:n, 0:n], ipvt[0:n])
        res.covar = cv
        d = np.diagonal(cv).copy()
        perror = np.zeros(n)
        wh = d >= 0
        perror[wh] = np.sqrt(d[wh])
        res.perror = perror
    return res
