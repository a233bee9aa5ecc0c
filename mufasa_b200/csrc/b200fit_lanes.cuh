// b200fit_lanes.cuh -- lane-parallel Levenberg-Marquardt step: P lanes of a warp cooperate on ONE pixel.
//
// Same algorithm, same tests and same constants as the serial state machine of b200fit_common.h (lm_adopt_first /
// lm_update / lm_propose / lm_par / lm_errors -- that version stays the readable restatement and is what the
// host emulation in tests/host_emul runs); here lane j of a P-lane group owns parameter j: p_j, the trial p_j,
// its scaling, its gradient entry and ROW j of the symmetric J^T J.  Scalars (chi2, trust radius, lambda, ...)
// are replicated, so all lanes of a group compute the same scalars.  CONTROL FLOW IS WARP-UNIFORM: every collective
// (shuffle, ballot) is executed by all 32 lanes with the constant full mask and segment width P, and a group whose
// branch is not taken simply discards the result.  (The first version let the groups of a warp diverge and used
// per-group masks: a non-constant mask makes nvcc wrap every shuffle in MATCH.ANY / WARPSYNC.COLLECTIVE /
// reconvergence code -- ~20 % of the executed instructions, profiles/r1j -- and the divergent lmpar loop cost the
// maximum trip count of the four groups anyway.)  Each loop over parameters of the serial code becomes one
// step plus a log2(P)-deep shuffle reduction, the Cholesky factorisation is right-looking with one pivot
// broadcast and P-1-j rank-1 broadcasts per column: the dependent-instruction chain of one LM step shrinks from
// ~20k instructions (thread per pixel, rolled loops) to ~2.5k, which is what bounds the late rounds of a fit,
// when only the few slowly converging pixels are still active (profiles/r1g_*).
#pragma once
#include "b200fit_common.h"

namespace b200fit {

// 1 / sqrt(x) and sqrt(x) for x > 0 finite: the hardware's ~20-bit seed (MUFU.RSQ64H) and two Newton steps, ~10
// branch-free instructions.  CUDA's rsqrt() / sqrt() are correctly rounded and handle every special case, at ~60 / ~30
// instructions with slow-path branches -- 9 % of the solver's executed instructions for the pivots of the Cholesky
// factorisation alone (profiles/r1o_solve2c_source_summary.txt), and each branch a break in the instruction stream
// of a kernel whose top stall is instruction fetch.  The result is within ~1 ulp; the factorisation and the norms it
// feeds are not sensitive to that (the test-only host emulation keeps the correctly rounded library calls).
__device__ __forceinline__ double rsqrt_pos(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double h = 0.5 * x;
    double e = fma(-h * r, r, 0.5);      // 0.5 (1 - x r^2)
    r = fma(r, e, r);
    e = fma(-h * r, r, 0.5);
    r = fma(r, e, r);
    return r;
}
__device__ __forceinline__ double sqrt_nonneg(double x) {   // x >= 0 finite (sums of squares)
    const double r = rsqrt_pos(fmax(x, 1e-300));
    double s = x * r;
    s = fma(fma(-s, s, x), 0.5 * r, s);  // one Heron correction of the product
    return s;
}

template <int P>
struct LaneGroup {
    static constexpr unsigned FULLMASK = 0xffffffffu;
    static constexpr unsigned GMASK = (P == 32) ? 0xffffffffu : ((1u << P) - 1u);
    int gl;          // lane index inside the group = parameter index
    int base;        // first lane of the group
    __device__ __forceinline__ LaneGroup() {
        const int lane = threadIdx.x & 31;
        gl = lane & (P - 1);
        base = lane & ~(P - 1);
    }
    // all of these must be reached by all 32 lanes of the warp
    __device__ __forceinline__ double sum(double v) const {
#pragma unroll
        for (int o = P / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o, P);
        return v;
    }
    __device__ __forceinline__ double max(double v) const {
#pragma unroll
        for (int o = P / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULLMASK, v, o, P));
        return v;
    }
    __device__ __forceinline__ double min(double v) const {
#pragma unroll
        for (int o = P / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULLMASK, v, o, P));
        return v;
    }
    __device__ __forceinline__ double bcast(double v, int src) const { return __shfl_sync(FULLMASK, v, src, P); }
    __device__ __forceinline__ unsigned ballot(bool pred) const {
        return (__ballot_sync(FULLMASK, pred) >> base) & GMASK;
    }
    __device__ __forceinline__ bool any_in_warp(bool pred) const { return __any_sync(FULLMASK, pred); }
};

// Registers of one lane.
template <int P>
struct LaneState {
    double p, pt, diag, b, bt, lo, hi;
    double hd, htd;            // diagonal entries H_jj / Ht_jj of this lane's row
    double H[P], Ht[P];        // row j of J^T J at the current / trial point
    double chi2, chi2t, chi2_last, delta, par, xnorm, pnorm, alpha, sHs;
    int status, iters, nevals, nrej, first, fresh;
    unsigned peg;
};

template <int P>
__device__ __forceinline__ void lane_load(const LaneGroup<P>& g, const LMState<P>& st, const LMConfig& cf,
                                          LaneState<P>& s) {
    const int j = g.gl;
    s.p = st.p[j];
    s.pt = st.pt[j];
    s.diag = st.diag[j];
    s.b = st.b[j];
    s.bt = st.bt[j];
    s.lo = cf.lo[j];
    s.hi = cf.hi[j];
#pragma unroll
    for (int k = 0; k < P; ++k) {
        const int q = (j <= k) ? hidx<P>(j, k) : hidx<P>(k, j);
        s.H[k] = st.H[q];
        s.Ht[k] = st.Ht[q];
    }
    s.hd = st.H[hidx<P>(j, j)];
    s.htd = st.Ht[hidx<P>(j, j)];
    s.chi2 = st.chi2;
    s.chi2t = st.chi2t;
    s.chi2_last = st.chi2_last;
    s.delta = st.delta;
    s.par = st.par;
    s.xnorm = st.xnorm;
    s.pnorm = st.pnorm;
    s.alpha = st.alpha;
    s.sHs = st.sHs;
    s.status = st.status;
    s.iters = st.iters;
    s.nevals = st.nevals;
    s.nrej = st.nrej;
    s.first = st.first;
    s.fresh = st.fresh;
    s.peg = st.peg;
}

// Write back what the next round (or the result writer) reads.  ``accepted``: p / b / H changed.
template <int P>
__device__ __forceinline__ void lane_store(const LaneGroup<P>& g, LMState<P>& st, const LaneState<P>& s,
                                           bool accepted) {
    const int j = g.gl;
    st.pt[j] = s.pt;
    st.diag[j] = s.diag;
    if (accepted) {
        st.p[j] = s.p;
        st.b[j] = s.b;
#pragma unroll
        for (int k = 0; k < P; ++k)
            if (k >= j) st.H[hidx<P>(j, k)] = s.H[k];
    }
    if (j == 0) {
        st.chi2 = s.chi2;
        st.chi2_last = s.chi2_last;
        st.delta = s.delta;
        st.par = s.par;
        st.xnorm = s.xnorm;
        st.pnorm = s.pnorm;
        st.alpha = s.alpha;
        st.sHs = s.sHs;
        st.status = s.status;
        st.iters = s.iters;
        st.nevals = s.nevals;
        st.nrej = s.nrej;
        st.first = s.first;
        st.fresh = s.fresh;
        st.peg = s.peg;
    }
}

// |D x| over the group
template <int P>
__device__ __forceinline__ double lane_dnorm(const LaneGroup<P>& g, double dg, double x) {
    const double t = dg * x;
    return sqrt_nonneg(g.sum(t * t));
}

// Right-looking Cholesky of A = H_free + par*diag^2 (fixed rows/cols -> identity).  Lane i ends up with the
// strictly lower part of row i of L in lrow[0..i-1] (zeros from the diagonal on, which lets the triangular solves
// run without per-lane conditions) and 1/L_ii in linv.  Returns false (group-uniform) if A is not numerically positive definite.
template <int P>
__device__ __forceinline__ bool lane_chol(const LaneGroup<P>& g, const double (&H)[P], double diag, double par,
                                          unsigned fixed, double (&lrow)[P], double& linv) {
    const int i = g.gl;
    const bool fi = (fixed >> i) & 1u;
    double a[P];
#pragma unroll
    for (int k = 0; k < P; ++k) {
        const bool fk = (fixed >> k) & 1u;
        double v = (fi || fk) ? 0.0 : H[k];
        if (k == i) v = fi ? 1.0 : v + par * diag * diag;
        a[k] = v;
    }
    double dorig = 0.0;
#pragma unroll
    for (int k = 0; k < P; ++k)
        if (k == i) dorig = a[k];
    bool ok = true;
    linv = 0.0;
#pragma unroll
    for (int j = 0; j < P; ++j) {
        // the pivot test runs on every lane's own (a[j], dorig) -- only lane j's is the diagonal -- so that one
        // broadcast and one vote replace two broadcasts
        const bool good = a[j] > 1e-14 * dorig;  // false: rank deficient in this direction
        const double dj = good ? a[j] : ((dorig > 0.0) ? dorig : 1.0);
        const double d = g.bcast(dj, j);         // (repaired) updated diagonal, held by lane j
        const unsigned votes = g.ballot(good);   // unconditionally: a collective must not sit behind ``ok &&``
        ok = ok && ((votes >> j) & 1u);
        const double inv = rsqrt_pos(d);         // 1 / L_jj  (d > 0: the pivot was repaired above if it was not)
        const double l = (i > j) ? a[j] * inv : 0.0;   // L_ij below the diagonal; the diagonal lives in linv only
        lrow[j] = l;
        if (i == j) linv = inv;
#pragma unroll
        for (int k = j + 1; k < P; ++k) {
            const double lk = g.bcast(l, k);     // L_kj
            a[k] -= l * lk;                      // meaningful for lanes i >= k
        }
    }
    return ok;
}

// L y = rhs  (column-oriented forward substitution); lane i returns y_i
template <int P>
__device__ __forceinline__ double lane_fwd(const LaneGroup<P>& g, const double (&lrow)[P], double linv, double rhs) {
    const int i = g.gl;
    double acc = rhs;
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const double yj = g.bcast(acc * linv, j);
        acc -= lrow[j] * yj;                     // lrow[j] == 0 for j >= i: lane i's acc is final after step i - 1
    }
    (void)i;
    return __dmul_rn(acc, linv);                 // rounded here, as in the serial restatement (no FMA contraction
                                                 // with the caller's subtraction)
}

// L^T x = y; lane i returns x_i
template <int P>
__device__ __forceinline__ double lane_bwd(const LaneGroup<P>& g, const double (&lrow)[P], double linv, double y) {
    const int i = g.gl;
    double x = 0.0;
#pragma unroll
    for (int jj = 0; jj < P; ++jj) {
        const int j = P - 1 - jj;
        const double s = g.sum(lrow[j] * x);     // lanes i <= j contribute 0: lrow[j] == 0 (and x is still 0)
        if (i == j) x = (y - s) * linv;
    }
    return x;
}

template <int P>
__device__ __forceinline__ double lane_newton_den(const LaneGroup<P>& g, const double (&lrow)[P], double linv,
                                                  double dg, double x, double dxnorm, bool fixedbit) {
    const double q = fixedbit ? 0.0 : dg * dg * x / dxnorm;
    const double w = lane_fwd<P>(g, lrow, linv, q);
    return g.sum(w * w);
}

// MINPACK lmpar on the normal equations (see lm_par).  Returns this lane's step component; updates par.
// ``live``: this group wants a step (false for groups that are done or hold no pixel).  The iteration runs until
// no group of the warp is live; a group that has found its lambda keeps executing the collectives and discards
// the results.  One call site each for the factorisation, the triangular solves and the Newton denominator.
template <int P>
__device__ __forceinline__ double lane_par(const LaneGroup<P>& g, const double (&H)[P], double dg, double rhs,
                                           unsigned fixed, double delta, double& par, bool live) {
    const bool fixedbit = (fixed >> g.gl) & 1u;
    double lrow[P], linv, x = 0.0, xout = 0.0;
    double dxnorm = INFINITY, fp = INFINITY, parl = 0.0, paru = 0.0;
    const double t0 = rhs / dg;
    const double gnorm = sqrt_nonneg(g.sum(t0 * t0));   // |D^-1 b|: the same in every iteration
#pragma unroll 1
    for (int it = 0; it <= 10; ++it) {
        if (!g.any_in_warp(live)) break;
        if (live && it > 0 && par == 0.0) par = fmax(LM_DWARF, 0.001 * paru);
        const bool spd = lane_chol<P>(g, H, dg, (it == 0) ? 0.0 : par, fixed, lrow, linv);
        const double temp = fp;
        const bool solved = (it > 0 || spd);
        const double y = lane_fwd<P>(g, lrow, linv, rhs);
        const double xs = lane_bwd<P>(g, lrow, linv, y);
        const double dxs = lane_dnorm<P>(g, dg, xs);
        // accept tests first: the Newton correction below is needed only by a group that goes on iterating (the
        // common last iteration of a call skips its forward solve)
        if (live) {
            if (solved) {
                x = xs;
                dxnorm = dxs;
                fp = dxnorm - delta;
            }
            if (it == 0 && spd && fp <= 0.1 * delta) {
                par = 0.0;
                xout = x;
                live = false;
            } else if (it > 0 && (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || it == 10)) {
                xout = x;
                live = false;
            }
        }
        if (!g.any_in_warp(live)) break;
        // Newton correction (fp / delta) / |L^-1 (D^2 x / |D x|)|^2
        const double den = lane_newton_den<P>(g, lrow, linv, dg, xs, dxs, fixedbit);
        if (!live) continue;
        const double corr = solved ? (fp / delta) / den : 0.0;
        if (it == 0) {
            parl = corr;                       // 0 when H is not positive definite
            paru = gnorm / delta;
            if (paru == 0.0) paru = LM_DWARF / fmin(delta, 0.1);
            par = fmax(par, parl);
            par = fmin(par, paru);
            if (par == 0.0) par = gnorm / dxnorm;
        } else {
            if (fp > 0.0) parl = fmax(parl, par);
            if (fp < 0.0) paru = fmin(paru, par);
            par = fmax(parl, par + corr);
        }
    }
    return xout;
}

template <int P>
__device__ __forceinline__ unsigned lane_pegged(const LaneGroup<P>& g, const LaneState<P>& s) {
    const double thr = 3.0 * LM_NOISE_REL * sqrt_nonneg((s.hd > 0.0 ? s.hd : 0.0) * s.chi2);
    const bool peg = (s.p <= s.lo && s.b < thr) || (s.p >= s.hi && s.b > -thr);
    return g.ballot(peg);
}

// One solver step of every group of the warp.  ``first_eval``: the group's state holds its very first evaluation
// (adopt it); ``valid``: the group holds a pixel at all.  Returns LM_DONE / LM_CONTINUE per group; ``accepted`` tells
// the caller whether p / b / H changed.  Warp-uniform: call with all 32 lanes.
template <int P>
__device__ __forceinline__ int lane_step(const LaneGroup<P>& g, LaneState<P>& s, const LMConfig& cf, bool valid,
                                         bool& accepted) {
    int done = valid ? LM_CONTINUE : LM_DONE;
    accepted = false;
    // ---- judge the trial evaluation (lm_adopt_first / lm_update) ------------------------------------------
    const bool first_eval = (s.nevals == 0);
    const double trap = g.sum((s.pt - s.p) * (s.b + s.bt));
    const double xnorm_t = lane_dnorm<P>(g, s.diag, s.pt);   // |D p| should the trial point be accepted
    if (valid) {
        const bool finite = (s.chi2t == s.chi2t) && !isinf(s.chi2t);
        if (first_eval) {
            s.nevals = 1;
            if (!finite) {
                s.status = -16;
                done = LM_DONE;
            } else {
                s.chi2 = s.chi2t;
                s.chi2_last = s.chi2t;
                s.b = s.bt;
                s.hd = s.htd;
#pragma unroll
                for (int k = 0; k < P; ++k) s.H[k] = s.Ht[k];
                accepted = true;
            }
        } else {
            s.nevals += 1;
            s.chi2_last = s.chi2t;
            if (!finite) {
                s.status = -16;
                done = LM_DONE;
            } else {
                const double direct = s.chi2 - s.chi2t;
                const double act_abs = (fabs(direct) > 40.0 * LM_NOISE_REL * s.chi2) ? direct : trap;
                double actred = -1.0;
                if (0.01 * s.chi2t < s.chi2) actred = act_abs / s.chi2;
                const double temp1sq = s.alpha * s.alpha * s.sHs / s.chi2;
                const double temp2sq = s.alpha * s.par * s.pnorm * s.pnorm / s.chi2;
                const double prered = temp1sq + temp2sq / 0.5;
                const double dirder = -(temp1sq + temp2sq);
                double ratio = 0.0;
                if (prered != 0.0) ratio = actred / prered;
                if (ratio <= 0.25) {
                    double temp = (actred >= 0.0) ? 0.5 : 0.5 * dirder / (dirder + 0.5 * actred);
                    if (0.01 * s.chi2t >= s.chi2 || temp < 0.1) temp = 0.1;
                    s.delta = temp * fmin(s.delta, s.pnorm / 0.1);
                    s.par = s.par / temp;
                } else if (s.par == 0.0 || ratio >= 0.75) {
                    s.delta = s.pnorm / 0.5;
                    s.par = 0.5 * s.par;
                }
                if (ratio >= 1e-4) {
                    s.p = s.pt;
                    s.b = s.bt;
                    s.hd = s.htd;
#pragma unroll
                    for (int k = 0; k < P; ++k) s.H[k] = s.Ht[k];
                    s.chi2 = s.chi2t;
                    s.xnorm = xnorm_t;
                    s.iters += 1;
                    s.first = 0;
                    s.fresh = 1;
                    accepted = true;
                } else {
                    s.nrej += 1;
                }
                int st = 0;
                if (fabs(actred) <= cf.ftol && prered <= cf.ftol && 0.5 * ratio <= 1.0) st = 1;
                else if (fabs(actred) <= LM_FTOL_NOISE && prered <= LM_FTOL_NOISE && 0.5 * ratio <= 1.0) st = 6;
                if (s.delta <= LM_XTOL * s.xnorm) st = (st == 1) ? 3 : 2;
                if (st != 0) {
                    s.status = st;
                    done = LM_DONE;
                } else if (s.iters >= cf.max_iter || s.nevals >= 2 * cf.max_iter + 2) {
                    s.status = 5;
                    done = LM_DONE;
                } else if (s.delta <= LM_MACHEP * s.xnorm) {
                    s.status = 7;
                    done = LM_DONE;
                }
            }
        }
    }
    // ---- propose the next trial point (lm_propose) ---------------------------------------------------------
    // start of an outer iteration (fresh Jacobian): pegged set, scaling, gradient test -- computed by every group,
    // applied by the groups that are at that point
    const unsigned peg_new = lane_pegged<P>(g, s);
    const bool fresh = (done == LM_CONTINUE) && s.fresh;
    {
        const bool pegbit = (peg_new >> g.gl) & 1u;
        const double fnorm = sqrt_nonneg(s.chi2);
        const double cn = pegbit ? 0.0 : sqrt_nonneg(s.hd > 0.0 ? s.hd : 0.0);
        double diag_new = s.diag;
        if (s.first)
            diag_new = (cn == 0.0) ? 1.0 : cn;
        else if (cn > s.diag)
            diag_new = cn;
        const double gn = (cn != 0.0 && fnorm != 0.0) ? fabs(s.b) / (cn * fnorm) : 0.0;
        const double gnorm = g.max(gn);
        const double xnorm0 = lane_dnorm<P>(g, diag_new, s.p);
        if (fresh) {
            s.peg = peg_new;
            s.diag = diag_new;
            if (s.first) {
                s.xnorm = xnorm0;
                s.delta = LM_FACTOR * s.xnorm;
                if (s.delta == 0.0) s.delta = LM_FACTOR;
            }
            if (gnorm <= LM_GTOL) {
                s.status = 4;
                done = LM_DONE;
            } else {
                s.fresh = 0;
            }
        }
    }
    const bool live = (done == LM_CONTINUE);
    const bool pegbit = (s.peg >> g.gl) & 1u;
    const bool fixedbit = pegbit || !(s.hd > 0.0);   // zero column: no information, leave untouched
    const unsigned fixed = g.ballot(fixedbit);
    const double rhs = fixedbit ? 0.0 : s.b;
    double par = s.par;
    double dl = lane_par<P>(g, s.H, s.diag, rhs, fixed, s.delta, par, live);
    // mpfit: parameters sitting on a limit may only step into the box; the whole step is shortened so that no
    // other parameter crosses a limit
    if (fixedbit) dl = 0.0;
    if (s.p <= s.lo && dl < 0.0) dl = 0.0;
    if (s.p >= s.hi && dl > 0.0) dl = 0.0;
    double a = 1.0;
    if (fabs(dl) > LM_MACHEP) {
        if (s.p + dl < s.lo) a = fmin(a, (s.lo - s.p) / dl);
        if (s.p + dl > s.hi) a = fmin(a, (s.hi - s.p) / dl);
    }
    const double alpha = g.min(a);
    dl *= alpha;
    double pn = s.p + dl;
    const double u1 = s.hi * (1.0 - (s.hi >= 0 ? 1.0 : -1.0) * LM_MACHEP) - (s.hi == 0.0 ? LM_MACHEP : 0.0);
    const double l1 = s.lo * (1.0 + (s.lo >= 0 ? 1.0 : -1.0) * LM_MACHEP) + (s.lo == 0.0 ? LM_MACHEP : 0.0);
    if (pn >= u1) pn = s.hi;
    if (pn <= l1) pn = s.lo;
    const double pnorm = lane_dnorm<P>(g, s.diag, dl);
    // |J s|^2 (dl is zero for every fixed / pegged parameter)
    double row = 0.0;
#pragma unroll
    for (int k = 0; k < P; ++k) row += s.H[k] * g.bcast(dl, k);
    const double sHs = g.sum(dl * row);
    if (live) {
        s.par = par;
        s.pt = pn;
        s.alpha = alpha;
        s.pnorm = pnorm;
        if (s.first) s.delta = fmin(s.delta, s.pnorm);
        s.sHs = sHs;
    }
    return done;
}

// Parameter error of this lane's parameter (see lm_errors).  Warp-uniform.
template <int P>
__device__ __forceinline__ double lane_error(const LaneGroup<P>& g, const LaneState<P>& s, double err2) {
    const unsigned peg = lane_pegged<P>(g, s);
    const double hmax = g.max(s.hd);
    const bool fixedbit = ((peg >> g.gl) & 1u) || !(s.hd > 1e-28 * hmax);
    const unsigned fixed = g.ballot(fixedbit);
    double lrow[P], linv;
    lane_chol<P>(g, s.H, 0.0, 0.0, fixed, lrow, linv);
    double perr = 0.0;
#pragma unroll 1
    for (int j = 0; j < P; ++j) {
        const double y = lane_fwd<P>(g, lrow, linv, (g.gl == j) ? 1.0 : 0.0);
        const double c = g.sum(y * y) * err2;   // (H^-1)_jj = | L^-1 e_j |^2
        if (g.gl == j) perr = (fixedbit || !(c > 0.0) || isinf(c)) ? 0.0 : sqrt(c);
    }
    return perr;
}

}  // namespace b200fit
