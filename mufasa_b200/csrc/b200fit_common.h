// b200fit_common.h -- host/device shared pieces of the B200 per-pixel fitter.
//
//  * hyperfine line tables (restated from mufasa/spec_models/m_constants.py:22-61 for N2H+ (1-0) and from
//    pyspeckit's ammonia_constants 'oneone' for NH3 (1,1), which the reference imports at m_constants.py:11-17)
//  * DeviceProblem: per-call constants derived on the host in fp64
//  * per-component coefficient set-up and the per-channel radiative-transfer + analytic Jacobian assembly
//    (velocity-space restatement of HyperfineModel.py:22-109)
//  * LMState<P>: the Levenberg-Marquardt state machine (damped normal equations + Cholesky, mpfit-style box
//    limits: pegged parameters, step clipping/truncation, snapping; zero error for pegged parameters).
//    The solver is written with ROLLED loops over small arrays on purpose: in the fit kernel it runs SIMT (one
//    lane per in-flight pixel of a warp's pixel group), and the first version of this file -- fully unrolled,
//    16.5k SASS instructions -- spent 70 % of its stall samples waiting for the instruction cache
//    (profiles/r1a_fit2c_source_summary.txt).
//
// Everything here is plain C++ so the same source is compiled by nvcc into the kernels and by g++ into the
// test-only host emulation (tests/host_emul).  It is NOT a CPU fallback: libb200fit.so only exposes CUDA paths.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#define B2_NOINL __host__ __device__ __noinline__
#define B2_ROLL _Pragma("unroll 1")
#else
#define B2_HD inline
#define B2_NOINL inline
#define B2_ROLL
#endif

namespace b200fit {

constexpr int MAXHF_FIT = 18;    // distinct hyperfines used by the fit kernel (NH3: 18, N2H+: 7 after merging)
constexpr int MAXHF_FULL = 18;   // entries of the reference table (NH3: 18, N2H+: 15), original order
constexpr double CKMS = 299792.458;         // BaseModel.py:26
constexpr double H_CGS = 6.62607015e-27;    // BaseModel.py:14
constexpr double KB_CGS = 1.380649e-16;     // BaseModel.py:15
constexpr double TCMB = 2.7315;             // BaseModel.py:13
constexpr float LOG2E_F = 1.4426950408889634f;
constexpr double KAPPA0 = 0.8493218002880191;   // sqrt(0.5*log2(e)):  exp(-u^2/(2 s^2)) = 2^-(KAPPA0*u/s)^2
constexpr float ZCUT2 = 28.0f;                  // gaussians below 2^-28 of their peak are not evaluated
constexpr float ZCUT = 5.291502622129181f;      // sqrt(ZCUT2)  (= 6.23 sigma)

struct LineTable {
    int nfull;
    double nu0_hz;
    double voff_full[MAXHF_FULL];
    double wts_full[MAXHF_FULL];
};

// NH3 (1,1) 'oneone' -- pyspeckit ammonia_constants (not in the reference tree; restated)
static const LineTable kNH3_11 = {
    18, 23.6944955e9,
    {19.8513, 19.3159, 7.88669, 7.46967, 7.35132, 0.460409, 0.322042, -0.0751680, -0.213003, 0.311034,
     0.192266, -0.132382, -0.250923, -7.23349, -7.37280, -7.81526, -19.4117, -19.5500},
    {0.0740740, 0.148148, 0.0925930, 0.166667, 0.0185190, 0.0370370, 0.0185190, 0.0185190, 0.0925930,
     0.0333330, 0.300000, 0.466667, 0.0333330, 0.0925930, 0.0185190, 0.166667, 0.0740740, 0.148148}};

// N2H+ (1-0) 'onezero' -- mufasa/spec_models/m_constants.py:24-26, 30-31, 45-47
static const LineTable kN2HP_10 = {
    15, 93173.7637e6,
    {-7.9930, -7.9930, -7.9930, -0.6112, -0.6112, -0.6112, 0.0000, 0.9533, 0.9533, 5.5371, 5.5371, 5.5371,
     5.9704, 5.9704, 6.9238},
    {0.025957, 0.065372, 0.019779, 0.004376, 0.034890, 0.071844, 0.259259, 0.156480, 0.028705, 0.041361,
     0.013309, 0.056442, 0.156482, 0.028705, 0.037038}};

// Per-call constants, computed on the host in fp64 (see make_device_problem in b200fit_host.h).
struct DeviceProblem {
    int ncomp, nchan, max_iter, weight_mode;
    int nhf;                       // distinct hyperfines, sorted by ascending velocity offset
    int nfull;                     // reference-table entries (original order) for the fp64 model kernels
    int ngroups;                   // groups of neighbouring hyperfines (windowing granularity of the fit kernel)
    int gstart[MAXHF_FIT + 1];     // group g = lines [gstart[g], gstart[g+1])
    int ngroups_full;              // the same for the reference-order table (contiguous runs of neighbouring lines)
    int gstart_full[MAXHF_FULL + 1];
    long long npix;
    double v0, dv, inv_dv;         // ascending velocity axis
    double ftol, signal_cut;
    double lo[8], hi[8];
    // fit kernel (velocity space): line k of a component at velocity v sits at  cen = c1mrho[k] + rho[k]*v,
    // with width sigma*rho[k];  rho[k] = (1 - voff_k/c) * nu0/nu_ref;  kdv[k] = KAPPA0*dv/rho[k]
    double rho[MAXHF_FIT], c1mrho[MAXHF_FIT], kdv[MAXHF_FIT];
    float lw[MAXHF_FIT];           // log2(normalised weight)
    double wfit[MAXHF_FIT];        // normalised weight (fp64 evaluation of the polish phase)
    // Planck terms, linear in channel offset di = i - ic0:  T0_i = T0c + dT0*di
    double ic0, T0c, dT0, Jbg0, Jbg1;
    // fp64 reference-order model (frequency space, HyperfineModel.py:91-107)
    double nu_ref_ghz;             // xarr_i = nu_ref_ghz * (1 - V_i/c)
    double lines_full[MAXHF_FULL]; // (1 - voff/c) * nu0 / 1e9
    double wts_full[MAXHF_FULL];   // tau_wts / sum(tau_wts)
};

// The slice of DeviceProblem the fit kernel's per-evaluation set-up indexes with a lane-dependent subscript
// (kept in shared memory: constant-bank reads with divergent indices serialise).
struct FitTables {
    double rho[MAXHF_FIT], c1mrho[MAXHF_FIT], kdv[MAXHF_FIT];
    double lo[8], hi[8];
    float lw[MAXHF_FIT];
    int gstart[MAXHF_FIT + 1];
};

// ------------------------------------------------------------------------------------------------------------
// Per-component coefficients for one evaluation (fp64 set-up, fp32 use).
struct CompCoef {
    float tau0;        // tau0
    float dJ0, dJ1;    // J(Tex) - J(Tbg), linear in di
    float dJT0, dJT1;  // dJ/dTex, linear in di
    float sA, sB;      // dtau/dv = sA * A,  dtau/dsigma = sB * B   (A = sum g z', B = sum g z'^2)
    float pad;
};

B2_HD void planck_terms(double T, double T0c, double dT0, double& J0, double& J1, double& JT0, double& JT1) {
    // J(T;T0) = T0/(e^x - 1), x = T0/T  (BaseModel.py:176-193); first-order expansion in T0 across the band
    const double x = T0c / T;
    const double ex = exp(x);
    const double em1 = expm1(x);
    const double inv = 1.0 / em1;
    J0 = T0c * inv;
    const double dJdT0 = (em1 - x * ex) * inv * inv;
    J1 = dJdT0 * dT0;
    JT0 = x * x * ex * inv * inv;                                   // dJ/dT
    const double dJTdx = ((2.0 * x * ex + x * x * ex) * em1 - 2.0 * x * x * ex * ex) * inv * inv * inv;
    JT1 = dJTdx / T * dT0;
}

B2_HD void comp_coefs(double T0c, double dT0, double Jbg0, double Jbg1, const double* p4, CompCoef& cc) {
    const double sigma = p4[1], tex = p4[2], tau0 = p4[3];
    double J0, J1, JT0, JT1;
    planck_terms(tex, T0c, dT0, J0, J1, JT0, JT1);
    cc.tau0 = (float)tau0;
    cc.dJ0 = (float)(J0 - Jbg0);
    cc.dJ1 = (float)(J1 - Jbg1);
    cc.dJT0 = (float)JT0;
    cc.dJT1 = (float)JT1;
    cc.sA = (float)(tau0 / (sigma * KAPPA0));
    cc.sB = (float)(tau0 / (sigma * KAPPA0 * KAPPA0));
    cc.pad = 0.f;
}

// Line set-up: z'(i) = ((i - nf) - f) * a  with  a = KAPPA0*dv/(sigma*rho_k);  stored as {nf, -f*a, a, lw}.
struct alignas(16) LineCoef {
    float nf, b, a, lw;
};

// cen/rho/kdv of line k, component velocity v and 1/sigma -> coefficients + the line's channel window [wlo, whi]
B2_HD void line_coef(double c1mrho_k, double rho_k, double kdv_k, float lw_k, double v0, double inv_dv, double v,
                     double inv_sigma, LineCoef& lc, float& wlo, float& whi) {
    const double cen = c1mrho_k + rho_k * v;
    const double ic = (cen - v0) * inv_dv;
    const double nf = rint(ic);
    const double a = kdv_k * inv_sigma;
    lc.nf = (float)nf;
    lc.b = (float)(-(ic - nf) * a);
    lc.a = (float)a;
    lc.lw = lw_k;
    const double hw = (double)ZCUT / a;
    wlo = (float)(ic - hw);
    whi = (float)(ic + hw);
}

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float b2_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#else
inline float b2_ex2(float x) { return exp2f(x); }
#endif

// One gaussian of one hyperfine at channel index fi (float, exact integer):
//   g = w_k exp(-u^2/(2 s_k^2));  accumulates G += g, A += g z', B += g z'^2
B2_HD void gauss_accum(float fi, const LineCoef& lc, float& G, float& A, float& B) {
    const float t = fi - lc.nf;
    const float z = fmaf(t, lc.a, lc.b);
    const float z2 = z * z;
    const float g = b2_ex2(lc.lw - z2);
    G += g;
    A = fmaf(g, z, A);
    B = fmaf(g, z2, B);
}

// Radiative transfer + Jacobian for one channel, background-subtracted form:
//   M_c = dJ_c (1 - e_c) + M_{c-1} e_c,  dJ_c = J(Tex_c) - J(Tbg),  e_c = exp(-tau_c)
template <int NC>
B2_HD void rt_channel(const float* G, const float* A, const float* B, float di, const CompCoef* cc, float& m,
                      float* J) {
    float Mprev = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const float tau = cc[c].tau0 * G[c];
        const float e = b2_ex2(-tau * LOG2E_F);
        const float ome = 1.f - e;
        const float dJ = fmaf(di, cc[c].dJ1, cc[c].dJ0);
        const float dJT = fmaf(di, cc[c].dJT1, cc[c].dJT0);
        const float D = (dJ - Mprev) * e;
#pragma unroll
        for (int q = 0; q < 4 * c; ++q) J[q] *= e;
        J[4 * c + 0] = D * cc[c].sA * A[c];
        J[4 * c + 1] = D * cc[c].sB * B[c];
        J[4 * c + 2] = dJT * ome;
        J[4 * c + 3] = D * G[c];
        Mprev = fmaf(dJ, ome, Mprev * e);
    }
    m = Mprev;
}

// ------------------------------------------------------------------------------------------------------------
// fp64 evaluation (the POLISH phase of a fit: the last iterations of every pixel, from the point where the fp32
// evaluation above stops resolving chi2, run on this model with mpfit's tests at face value -- DESIGN.md 3.1).
// Same velocity-space form; natural units: q = (V_i - cen_k) / (sigma rho_k), g = w_k exp(-q^2 / 2),
//   G = sum g, A = sum g q, B = sum g q^2;  dtau/dv = tau0 A / sigma,  dtau/dsigma = tau0 B / sigma.
// The Planck terms J(T; T0_i) are a quadratic through their exact values at the first, centre and last channel
// (T0 is exactly linear in the channel index; the cubic remainder is ~(band / nu)^3 ~ 2e-12 relative).
constexpr double QCUT2_64 = 80.0;   // gaussians below exp(-40) = 4e-18 of their peak are skipped

struct CompCoef64 {
    double tau0, inv_sigma;
    double dJ[3];     // J(Tex) - J(Tbg) = dJ[0] + di (dJ[1] + di dJ[2])
    double dJT[3];    // dJ/dTex, likewise
};

struct LineCoef64 {
    double ic, a, w;  // q(i) = (i - ic) * a
};

B2_HD void planck_exact(double T, double T0, double& J, double& JT) {   // BaseModel.py:176-193 and its T-derivative
    const double x = T0 / T;
    const double em1 = expm1(x);
    J = T0 / em1;
    JT = x * x * (em1 + 1.0) / (em1 * em1);
}

B2_HD void comp_coefs64(double T0c, double dT0, double ic0, const double* p4, CompCoef64& cc) {
    const double h = (ic0 > 0.5) ? ic0 : 0.5;   // half the band, in channels
    double fj[3], ft[3];
    B2_ROLL
    for (int m = 0; m < 3; ++m) {
        const double T0 = T0c + dT0 * (double)(m - 1) * h;
        double Jx, JTx, Jb, JTb;
        planck_exact(p4[2], T0, Jx, JTx);
        planck_exact(TCMB, T0, Jb, JTb);
        fj[m] = Jx - Jb;
        ft[m] = JTx;
    }
    cc.tau0 = p4[3];
    cc.inv_sigma = 1.0 / p4[1];
    cc.dJ[0] = fj[1];
    cc.dJ[1] = (fj[2] - fj[0]) / (2.0 * h);
    cc.dJ[2] = (fj[2] - 2.0 * fj[1] + fj[0]) / (2.0 * h * h);
    cc.dJT[0] = ft[1];
    cc.dJT[1] = (ft[2] - ft[0]) / (2.0 * h);
    cc.dJT[2] = (ft[2] - 2.0 * ft[1] + ft[0]) / (2.0 * h * h);
}

// line k of a component at velocity v: channel position, scale, weight; the window [wlo, whi] (in channels) outside
// which the gaussian is skipped
B2_HD void line_coef64(double c1mrho_k, double rho_k, double w_k, double v0, double inv_dv, double dv, double v,
                       double inv_sigma, LineCoef64& lc, double& wlo, double& whi) {
    const double cen = c1mrho_k + rho_k * v;
    lc.ic = (cen - v0) * inv_dv;
    lc.a = dv * inv_sigma / rho_k;
    lc.w = w_k;
    const double hw = 8.944271909999159 / lc.a;   // sqrt(QCUT2_64)
    wlo = lc.ic - hw;
    whi = lc.ic + hw;
}

B2_HD void gauss_accum64(double fi, const LineCoef64& lc, double& G, double& A, double& B) {
    const double q = (fi - lc.ic) * lc.a;
    const double q2 = q * q;
    if (q2 < QCUT2_64) {
        const double g = lc.w * exp(-0.5 * q2);
        G += g;
        A = fma(g, q, A);
        B = fma(g, q2, B);
    }
}

template <int NC>
B2_HD void rt_channel64(const double* G, const double* A, const double* B, double di, const CompCoef64* cc, double& m,
                        double* J) {
    double Mprev = 0.0;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const double tau = cc[c].tau0 * G[c];
        const double e = exp(-tau);
        const double ome = -expm1(-tau);
        const double dJ = fma(di, fma(di, cc[c].dJ[2], cc[c].dJ[1]), cc[c].dJ[0]);
        const double dJT = fma(di, fma(di, cc[c].dJT[2], cc[c].dJT[1]), cc[c].dJT[0]);
        const double D = (dJ - Mprev) * e;
        const double ts = cc[c].tau0 * cc[c].inv_sigma;
#pragma unroll
        for (int q = 0; q < 4 * c; ++q) J[q] *= e;
        J[4 * c + 0] = D * ts * A[c];
        J[4 * c + 1] = D * ts * B[c];
        J[4 * c + 2] = dJT * ome;
        J[4 * c + 3] = D * G[c];
        Mprev = fma(dJ, ome, Mprev * e);
    }
    m = Mprev;
}

// ------------------------------------------------------------------------------------------------------------
// Levenberg-Marquardt state machine: MINPACK-style trust region on the NORMAL EQUATIONS.
//   H = J^T J (packed upper triangle), b = J^T r with r = y - M, chi2 = r^T r   (UNWEIGHTED sums: the per-pixel
//   weight 1/err^2 is a common factor -- it cancels in the step and is applied to chi2 / covariance at the end).
// The damped system (H + par D^2) delta = b is solved by Cholesky; ``par`` is chosen so that |D delta| matches the
// trust radius (More' 1978, the lmpar iteration, restated for Cholesky factors instead of QR).  Box limits follow
// mpfit: pegged parameters (on a limit, gradient pointing out) are frozen, parameters on a limit may only step
// inwards, the step is shortened so no parameter crosses a limit, and landing on a limit snaps to it exactly.
// Every evaluation returns chi2 AND the normal equations at the trial point (fused), so an accepted step needs
// no second pass.
constexpr int LM_CONTINUE = 0, LM_DONE = 1;
constexpr double LM_NOISE_REL = 5e-7;     // relative evaluation noise of a chi2 built from an fp32 model
constexpr double LM_MACHEP = 2.220446049250313e-16;
constexpr double LM_DWARF = 2.2250738585072014e-308;
constexpr double LM_XTOL = 1e-10, LM_GTOL = 1e-10, LM_FACTOR = 100.0;
// MINPACK's termination test 6 ("ftol is too small: no further reduction in the sum of squares is possible") fires
// when both the actual and the predicted relative reduction are below the machine precision of the function.
// Here the function is chi2 of an fp32 model: its precision is ~1e-7, not 2e-16.  Without this test a pixel whose
// chi2 has stalled at that level keeps taking noise-driven steps along any flat valley until maxiter (200), where
// mpfit -- evaluating in fp64 -- stopped with status 1 long before.  3e-8 is the largest floor that leaves the parity
// statistics against the oracle unchanged; status 6 counts as converged, as in MINPACK.
constexpr double LM_FTOL_NOISE = 3e-8;

// Per-call solver constants (limits live in shared memory in the kernel).
struct LMConfig {
    const double* lo;
    const double* hi;
    double ftol;
    int max_iter;
};

template <int P>
struct LMState {
    static constexpr int NH = P * (P + 1) / 2;
    double p[P], pt[P], diag[P], b[P], bt[P];
    double H[NH], Ht[NH];
    double chi2, chi2t, chi2_last, delta, par, xnorm, pnorm, alpha, sHs;
    int status, iters, nevals, nrej, first, fresh;
    unsigned peg;
    int mode;   // 0: fp32 evaluation (noise-aware tests); 1: polish phase, fp64 evaluation, mpfit's tests at face value
};

template <int P>
B2_HD int hidx(int i, int j) {   // packed upper triangle, i <= j
    return i * P - (i * (i - 1)) / 2 + (j - i);
}

template <int P>
B2_HD double hget(const double* H, int i, int j) {
    return H[(i <= j) ? hidx<P>(i, j) : hidx<P>(j, i)];
}

B2_HD int lidx(int i, int k) {   // packed lower triangle, k <= i
    return (i * (i + 1)) / 2 + k;
}

// Cholesky factor (packed lower, L; reciprocal diagonal in Linv) of A = H_free + par*diag^2; fixed rows/cols are
// replaced by identity.  Returns false if A is not (numerically) positive definite.
// ONE out-of-line instance, fully unrolled inside (the factor lives in registers while it is built): unrolled
// code is ~3x fewer instructions than rolled loops with index arithmetic, and a single instance keeps the
// kernel's instruction footprint small.
template <int P>
B2_NOINL bool lm_chol(const double* __restrict__ H, const double* __restrict__ dg, double par, unsigned fixed,
                      double* __restrict__ Lout, double* __restrict__ Linv) {
    bool ok = true;
    double L[P * (P + 1) / 2];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const bool fj = (fixed >> j) & 1u;
        double d = fj ? 1.0 : H[hidx<P>(j, j)] + par * dg[j] * dg[j];
        const double d0 = d;
#pragma unroll
        for (int k = 0; k < j; ++k) d -= L[lidx(j, k)] * L[lidx(j, k)];
        if (!(d > 1e-14 * d0)) {   // rank deficient in this direction
            ok = false;
            d = (d0 > 0.0) ? d0 : 1.0;
        }
        const double s = sqrt(d);
        const double inv = 1.0 / s;
        L[lidx(j, j)] = s;
        Linv[j] = inv;
#pragma unroll
        for (int i = j + 1; i < P; ++i) {
            const bool fi = (fixed >> i) & 1u;
            double v = (fi || fj) ? 0.0 : H[hidx<P>(j, i)];
#pragma unroll
            for (int k = 0; k < j; ++k) v -= L[lidx(i, k)] * L[lidx(j, k)];
            L[lidx(i, j)] = v * inv;
        }
    }
#pragma unroll
    for (int q = 0; q < P * (P + 1) / 2; ++q) Lout[q] = L[q];
    return ok;
}

template <int P>
B2_NOINL void lm_fwd(const double* __restrict__ L, const double* __restrict__ Linv, const double* __restrict__ rhs,
                     double* __restrict__ yout) {   // L y = rhs
    double y[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
        double v = rhs[i];
#pragma unroll
        for (int k = 0; k < i; ++k) v -= L[lidx(i, k)] * y[k];
        y[i] = v * Linv[i];
    }
#pragma unroll
    for (int i = 0; i < P; ++i) yout[i] = y[i];
}

template <int P>
B2_NOINL void lm_bwd(const double* __restrict__ L, const double* __restrict__ Linv, const double* __restrict__ yin,
                     double* __restrict__ xout) {   // L^T x = y
    double x[P];
#pragma unroll
    for (int ii = 0; ii < P; ++ii) {
        const int i = P - 1 - ii;
        double v = yin[i];
#pragma unroll
        for (int k = i + 1; k < P; ++k) v -= L[lidx(k, i)] * x[k];
        x[i] = v * Linv[i];
    }
#pragma unroll
    for (int i = 0; i < P; ++i) xout[i] = x[i];
}

template <int P>
B2_HD double lm_dnorm(const double* dg, const double* x) {
    double s = 0.0;
    B2_ROLL
    for (int j = 0; j < P; ++j) s += dg[j] * x[j] * dg[j] * x[j];
    return sqrt(s);
}

// |L^-1 (D^2 x / |D x|)|^2 : the Newton correction denominator of lmpar
template <int P>
B2_HD double lm_newton_den(const double* L, const double* Linv, const double* dg, const double* x, double dxnorm,
                           unsigned fixed) {
    double q[P], w[P];
    B2_ROLL
    for (int j = 0; j < P; ++j) q[j] = ((fixed >> j) & 1u) ? 0.0 : dg[j] * dg[j] * x[j] / dxnorm;
    lm_fwd<P>(L, Linv, q, w);
    double wn = 0.0;
    B2_ROLL
    for (int j = 0; j < P; ++j) wn += w[j] * w[j];
    return wn;
}

// MINPACK lmpar on normal equations.  rhs = b with fixed entries zeroed.  Returns delta-step in x, updates par.
template <int P>
B2_HD void lm_par(const double* H, const double* dg, const double* rhs, unsigned fixed, double delta, double& par,
                  double* x) {
    double L[P * (P + 1) / 2], Linv[P], y[P];
    double dxnorm = INFINITY, fp = INFINITY, parl = 0.0, paru = 0.0;
    // it == 0: Gauss-Newton direction (par = 0); it >= 1: the lmpar iteration
    B2_ROLL
    for (int it = 0;; ++it) {
        if (it > 0 && par == 0.0) par = fmax(LM_DWARF, 0.001 * paru);
        const bool spd = lm_chol<P>(H, dg, (it == 0) ? 0.0 : par, fixed, L, Linv);
        const double temp = fp;
        if (it > 0 || spd) {
            lm_fwd<P>(L, Linv, rhs, y);
            lm_bwd<P>(L, Linv, y, x);
            dxnorm = lm_dnorm<P>(dg, x);
            fp = dxnorm - delta;
        }
        if (it == 0) {
            if (spd && fp <= 0.1 * delta) {
                par = 0.0;
                return;
            }
            if (spd) parl = (fp / delta) / lm_newton_den<P>(L, Linv, dg, x, dxnorm, fixed);
            double gn = 0.0;
                    B2_ROLL
                    for (int j = 0; j < P; ++j) {
                const double t = rhs[j] / dg[j];
                gn += t * t;
            }
            const double gnorm = sqrt(gn);
            paru = gnorm / delta;
            if (paru == 0.0) paru = LM_DWARF / fmin(delta, 0.1);
            par = fmax(par, parl);
            par = fmin(par, paru);
            if (par == 0.0) par = gnorm / dxnorm;
            continue;
        }
        if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || it == 10) break;
        const double parc = (fp / delta) / lm_newton_den<P>(L, Linv, dg, x, dxnorm, fixed);
        if (fp > 0.0) parl = fmax(parl, par);
        if (fp < 0.0) paru = fmin(paru, par);
        par = fmax(parl, par + parc);
    }
}

template <int P>
B2_HD unsigned lm_pegged(const LMState<P>& s, const LMConfig& cf) {
    // mpfit: a parameter sitting on a limit whose chi2 gradient points out of the box gets its Jacobian column
    // zeroed.  d(chi2)/dp_j = -2 b_j:  lower limit & b_j < 0, or upper limit & b_j > 0.
    // The gradient of an fp32-evaluated model carries noise of relative size ~LM_NOISE_REL in MINPACK's cosine
    // measure |b_j| / (|J_j| |r|); a parameter on a limit is released only by a gradient that is significant.
    unsigned peg = 0;
    B2_ROLL
    for (int j = 0; j < P; ++j) {
        const double hjj = s.H[hidx<P>(j, j)];
        const double thr = s.mode ? 0.0 : 3.0 * LM_NOISE_REL * sqrt((hjj > 0.0 ? hjj : 0.0) * s.chi2);
        if (s.p[j] <= cf.lo[j] && s.b[j] < thr) peg |= 1u << j;
        if (s.p[j] >= cf.hi[j] && s.b[j] > -thr) peg |= 1u << j;
    }
    return peg;
}

// Start: returns false (status 0) when the guess is not finite or outside the box (mpfit: "parameters are not
// within PARINFO limits" -> pyspeckit raises -> pixel not fitted).
template <int P>
B2_HD bool lm_init(LMState<P>& s, const double* guess, const LMConfig& cf) {
    bool ok = true;
    B2_ROLL
    for (int j = 0; j < P; ++j) {
        const double g = guess[j];
        s.p[j] = g;
        s.pt[j] = g;
        s.diag[j] = 1.0;
        if (!(g >= cf.lo[j] && g <= cf.hi[j])) ok = false;   // also catches NaN
    }
    s.par = 0.0;
    s.delta = 0.0;
    s.xnorm = 0.0;
    s.pnorm = 0.0;
    s.alpha = 1.0;
    s.sHs = 0.0;
    s.iters = 0;
    s.nevals = 0;
    s.nrej = 0;
    s.status = 0;
    s.peg = 0;
    s.first = 1;
    s.fresh = 1;
    s.mode = 0;
    s.chi2 = 0.0;
    s.chi2t = 0.0;
    s.chi2_last = 0.0;
    return ok;
}

// After the very first evaluation (stored in the *t slots): adopt it as the current point.
template <int P>
B2_HD int lm_adopt_first(LMState<P>& s) {
    s.nevals = 1;
    if (!(s.chi2t == s.chi2t) || isinf(s.chi2t)) {
        s.status = -16;
        return LM_DONE;
    }
    s.chi2 = s.chi2t;
    s.chi2_last = s.chi2t;
    B2_ROLL
    for (int j = 0; j < P; ++j) s.b[j] = s.bt[j];
    B2_ROLL
    for (int q = 0; q < LMState<P>::NH; ++q) s.H[q] = s.Ht[q];
    return LM_CONTINUE;
}

// Polish phase.  A pixel the fp32 phase declared converged (status 1-4 or 6) goes on from its current point with
// fp64 evaluations and mpfit's tests at face value: the point is re-evaluated (pt = p), the result adopted, and the
// iteration continues with the trust region, damping and scaling it had, until mpfit's own tests fire.
template <int P>
B2_HD bool lm_wants_polish(const LMState<P>& s) {
    return s.mode == 0 && (s.status == 1 || s.status == 2 || s.status == 3 || s.status == 4 || s.status == 6);
}

template <int P>
B2_HD void lm_begin_polish(LMState<P>& s) {
    s.mode = 1;
    s.status = 0;
    B2_ROLL
    for (int j = 0; j < P; ++j) s.pt[j] = s.p[j];
}

// After the fp64 re-evaluation at the current point (stored in the *t slots).
template <int P>
B2_HD int lm_adopt_polish(LMState<P>& s) {
    s.nevals += 1;
    if (!(s.chi2t == s.chi2t) || isinf(s.chi2t)) {
        s.status = -16;
        return LM_DONE;
    }
    s.chi2 = s.chi2t;
    s.chi2_last = s.chi2t;
    B2_ROLL
    for (int j = 0; j < P; ++j) s.b[j] = s.bt[j];
    B2_ROLL
    for (int q = 0; q < LMState<P>::NH; ++q) s.H[q] = s.Ht[q];
    s.fresh = 1;
    return LM_CONTINUE;
}

// Propose the next trial point s.pt.  Returns LM_DONE when a convergence test fires (s.status set).
template <int P>
B2_HD int lm_propose(LMState<P>& s, const LMConfig& cf) {
    double rhs[P], dl[P];
    if (s.fresh) {   // start of an outer iteration: new Jacobian => new pegged set, scaling, gradient test
        s.peg = lm_pegged(s, cf);
        double gnorm = 0.0;
        const double fnorm = sqrt(s.chi2);
            B2_ROLL
            for (int j = 0; j < P; ++j) {
            const double hjj = s.H[hidx<P>(j, j)];
            const double cn = ((s.peg >> j) & 1u) ? 0.0 : sqrt(hjj > 0.0 ? hjj : 0.0);
            if (s.first)
                s.diag[j] = (cn == 0.0) ? 1.0 : cn;
            else if (cn > s.diag[j])
                s.diag[j] = cn;
            if (cn != 0.0 && fnorm != 0.0) gnorm = fmax(gnorm, fabs(s.b[j]) / (cn * fnorm));
        }
        if (s.first) {
            s.xnorm = lm_dnorm<P>(s.diag, s.p);
            s.delta = LM_FACTOR * s.xnorm;
            if (s.delta == 0.0) s.delta = LM_FACTOR;
        }
        if (gnorm <= LM_GTOL) {
            s.status = 4;
            return LM_DONE;
        }
        s.fresh = 0;
    }
    unsigned fixed = s.peg;
    B2_ROLL
    for (int j = 0; j < P; ++j) {
        if (!(s.H[hidx<P>(j, j)] > 0.0)) fixed |= 1u << j;   // zero column: no information, leave untouched
        rhs[j] = ((fixed >> j) & 1u) ? 0.0 : s.b[j];
    }
    double par = s.par;
    lm_par<P>(s.H, s.diag, rhs, fixed, s.delta, par, dl);
    s.par = par;
    // mpfit: parameters sitting on a limit may only step into the box, and the whole step is shortened so that
    // no other parameter crosses a limit
    double alpha = 1.0;
    B2_ROLL
    for (int j = 0; j < P; ++j) {
        const double pj = s.p[j], lo = cf.lo[j], hi = cf.hi[j];
        double d = dl[j];
        if ((fixed >> j) & 1u) d = 0.0;
        if (pj <= lo && d < 0.0) d = 0.0;
        if (pj >= hi && d > 0.0) d = 0.0;
        dl[j] = d;
        if (fabs(d) > LM_MACHEP) {
            if (pj + d < lo) alpha = fmin(alpha, (lo - pj) / d);
            if (pj + d > hi) alpha = fmin(alpha, (hi - pj) / d);
        }
    }
    B2_ROLL
    for (int j = 0; j < P; ++j) {
        const double lo = cf.lo[j], hi = cf.hi[j];
        dl[j] *= alpha;
        double pn = s.p[j] + dl[j];
        // snap to the limit when the step lands on it (mpfit ulim1/llim1 rule)
        const double u1 = hi * (1.0 - (hi >= 0 ? 1.0 : -1.0) * LM_MACHEP) - (hi == 0.0 ? LM_MACHEP : 0.0);
        const double l1 = lo * (1.0 + (lo >= 0 ? 1.0 : -1.0) * LM_MACHEP) + (lo == 0.0 ? LM_MACHEP : 0.0);
        if (pn >= u1) pn = hi;
        if (pn <= l1) pn = lo;
        s.pt[j] = pn;
    }
    s.alpha = alpha;
    s.pnorm = lm_dnorm<P>(s.diag, dl);
    if (s.first) s.delta = fmin(s.delta, s.pnorm);
    // |J s|^2 with the pegged columns zeroed (dl is already zero for every fixed / pegged parameter)
    double sHs = 0.0;
    {
        int q = 0;
        B2_ROLL
        for (int i = 0; i < P; ++i) {
            const double di = dl[i];
            sHs += s.H[q++] * di * di;
            B2_ROLL
            for (int j = i + 1; j < P; ++j) sHs += 2.0 * s.H[q++] * di * dl[j];
        }
    }
    s.sHs = sHs;
    return LM_CONTINUE;
}

// Judge the trial evaluation (chi2t/Ht/bt filled), MINPACK lmdif / mpfit bookkeeping.  Returns LM_DONE when finished.
template <int P>
B2_HD int lm_update(LMState<P>& s, const LMConfig& cf) {
    s.nevals += 1;
    s.chi2_last = s.chi2t;
    const bool finite = (s.chi2t == s.chi2t) && !isinf(s.chi2t);
    if (!finite) {
        s.status = -16;
        return LM_DONE;
    }
    // Actual reduction.  chi2 is built from an fp32 model, so the difference chi2 - chi2t cannot be resolved below
    // ~LM_NOISE_REL*chi2, while the gradients b = J^T r stay accurate to ~1e-6 of their own size.  For small
    // reductions use the trapezoid rule along the step,  chi2(p) - chi2(p+s) = 2 int_0^1 s.b(t) dt ~ s.(b + bt),
    // whose error is third order in the step.
    const double direct = s.chi2 - s.chi2t;
    double trap = 0.0;
    B2_ROLL
    for (int j = 0; j < P; ++j) trap += (s.pt[j] - s.p[j]) * (s.b[j] + s.bt[j]);
    const double act_abs = (s.mode || fabs(direct) > 40.0 * LM_NOISE_REL * s.chi2) ? direct : trap;
    double actred = -1.0;
    if (0.01 * s.chi2t < s.chi2) actred = act_abs / s.chi2;
    // mpfit applies alpha a second time to |J s| (temp1) -- kept
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
            B2_ROLL
            for (int j = 0; j < P; ++j) {
            s.p[j] = s.pt[j];
            s.b[j] = s.bt[j];
        }
            B2_ROLL
            for (int q = 0; q < LMState<P>::NH; ++q) s.H[q] = s.Ht[q];
        s.chi2 = s.chi2t;
        s.xnorm = lm_dnorm<P>(s.diag, s.p);
        s.iters += 1;
        s.first = 0;
        s.fresh = 1;
    } else {
        s.nrej += 1;
    }
    int st = 0;
    if (fabs(actred) <= cf.ftol && prered <= cf.ftol && 0.5 * ratio <= 1.0) st = 1;
    else if (!s.mode && fabs(actred) <= LM_FTOL_NOISE && prered <= LM_FTOL_NOISE && 0.5 * ratio <= 1.0) st = 6;
    if (s.delta <= LM_XTOL * s.xnorm) st = (st == 1) ? 3 : 2;
    if (st != 0) {
        s.status = st;
        return LM_DONE;
    }
    if (s.iters >= cf.max_iter || s.nevals >= 2 * cf.max_iter + 2) {
        s.status = 5;
        return LM_DONE;
    }
    if (s.delta <= LM_MACHEP * s.xnorm) {
        s.status = 7;
        return LM_DONE;
    }
    return LM_CONTINUE;
}

// chi2 AT THE RETURNED PARAMETERS.  (mpfit's own fnorm is max(fnorm(final), fnorm1(last trial))**2 -- the chi2 of a
// rejected step when the fit ends on one.  MUFASA never reads it: pyspeckit stores no per-pixel chi2 and
// UltraCube.get_chisq / get_rss re-evaluate the model at parcube, UltraCube.py:1384-1565.)
template <int P>
B2_HD double lm_reported_chi2(const LMState<P>& s) {
    return s.chi2;
}

// Parameter errors: sqrt(diag(err^2 (H_FF)^-1)), zero for pegged parameters (mpfit calc_covar on the
// column-zeroed Jacobian); 0 as well for directions without sensitivity.
template <int P>
B2_HD void lm_errors(const LMState<P>& s, const LMConfig& cf, double err2, double* perr) {
    unsigned fixed = lm_pegged(s, cf);
    double hmax = 0.0;
    B2_ROLL
    for (int j = 0; j < P; ++j) hmax = fmax(hmax, s.H[hidx<P>(j, j)]);
    B2_ROLL
    for (int j = 0; j < P; ++j)
        if (!(s.H[hidx<P>(j, j)] > 1e-28 * hmax)) fixed |= 1u << j;
    double L[P * (P + 1) / 2], Linv[P], rhs[P], y[P];
    B2_ROLL
    for (int j = 0; j < P; ++j) rhs[j] = 0.0;
    lm_chol<P>(s.H, rhs /* zero scaling: par = 0 anyway */, 0.0, fixed, L, Linv);
    // (H^-1)_jj = | L^-1 e_j |^2
    B2_ROLL
    for (int j = 0; j < P; ++j) {
            B2_ROLL
            for (int k = 0; k < P; ++k) rhs[k] = (k == j) ? 1.0 : 0.0;
        lm_fwd<P>(L, Linv, rhs, y);
        double c = 0.0;
            B2_ROLL
            for (int k = 0; k < P; ++k) c += y[k] * y[k];
        c *= err2;
        perr[j] = (((fixed >> j) & 1u) || !(c > 0.0) || isinf(c)) ? 0.0 : sqrt(c);
    }
}

}  // namespace b200fit
