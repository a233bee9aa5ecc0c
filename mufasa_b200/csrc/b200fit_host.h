// b200fit_host.h -- host-side derivation of the per-call device constants (fp64).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <utility>
#include <vector>

#include "../../include/b200fit.h"
#include "b200fit_common.h"

namespace b200fit {

inline const LineTable* line_table(int model_id) {
    if (model_id == B200FIT_MODEL_NH3_11) return &kNH3_11;
    if (model_id == B200FIT_MODEL_N2HP_10) return &kN2HP_10;
    return nullptr;
}

// Returns 0 or a B200FIT_E_* code.
inline int make_device_problem(const b200fit_problem& in, DeviceProblem& out) {
    const LineTable* lt = line_table(in.model_id);
    if (!lt) return B200FIT_E_MODEL;
    if (in.ncomp != 1 && in.ncomp != 2) return B200FIT_E_ARG;
    if (in.nchan < 8 || in.nchan > B200FIT_MAX_NCHAN) return B200FIT_E_ARG;
    if (!(in.dv > 0.0) || !std::isfinite(in.v0) || in.npix < 0) return B200FIT_E_ARG;
    std::memset(&out, 0, sizeof(out));
    out.ncomp = in.ncomp;
    out.nchan = in.nchan;
    out.max_iter = in.max_iter > 0 ? in.max_iter : 200;
    out.weight_mode = in.weight_mode;
    out.npix = in.npix;
    out.v0 = in.v0;
    out.dv = in.dv;
    out.inv_dv = 1.0 / in.dv;
    out.ftol = in.ftol > 0 ? in.ftol : 1e-10;
    out.signal_cut = in.signal_cut;
    for (int j = 0; j < 8; ++j) {
        out.lo[j] = in.lo[j];
        out.hi[j] = in.hi[j];
    }
    const double nu0_ghz = lt->nu0_hz / 1e9;
    const double nu_ref = in.nu_ref_ghz > 0 ? in.nu_ref_ghz : nu0_ghz;
    out.nu_ref_ghz = nu_ref;
    // reference-order table (HyperfineModel.py:91-95)
    double wsum = 0.0;
    for (int k = 0; k < lt->nfull; ++k) wsum += lt->wts_full[k];
    out.nfull = lt->nfull;
    for (int k = 0; k < lt->nfull; ++k) {
        out.lines_full[k] = (1 - lt->voff_full[k] / CKMS) * lt->nu0_hz / 1e9;
        out.wts_full[k] = lt->wts_full[k] / wsum;
    }
    // contiguous runs of neighbouring lines in the reference-order table (window tests of the fp64 model kernels)
    out.ngroups_full = 0;
    for (int k = 0; k < lt->nfull; ++k)
        if (k == 0 || std::fabs(lt->voff_full[k] - lt->voff_full[k - 1]) > 3.0) out.gstart_full[out.ngroups_full++] = k;
    for (int g = out.ngroups_full; g <= MAXHF_FULL; ++g) out.gstart_full[g] = lt->nfull;
    // distinct hyperfines, ascending velocity offset; equal offsets merged (exact: identical gaussians)
    std::vector<std::pair<double, double>> v;
    for (int k = 0; k < lt->nfull; ++k) {
        bool merged = false;
        for (auto& e : v)
            if (e.first == lt->voff_full[k]) {
                e.second += lt->wts_full[k];
                merged = true;
            }
        if (!merged) v.emplace_back(lt->voff_full[k], lt->wts_full[k]);
    }
    std::sort(v.begin(), v.end());
    out.nhf = (int)v.size();
    for (int k = 0; k < out.nhf; ++k) {
        const double rho = (1.0 - v[k].first / CKMS) * nu0_ghz / nu_ref;
        out.rho[k] = rho;
        out.c1mrho[k] = CKMS * (1.0 - rho);
        out.kdv[k] = KAPPA0 * in.dv / rho;
        out.lw[k] = (float)std::log2(v[k].second / wsum);
        out.wfit[k] = v[k].second / wsum;
    }
    // windowing groups: neighbouring hyperfines closer than GROUP_GAP km/s share one channel window
    const double GROUP_GAP = 3.0;
    out.ngroups = 0;
    for (int k = 0; k < out.nhf; ++k)
        if (k == 0 || v[k].first - v[k - 1].first > GROUP_GAP) out.gstart[out.ngroups++] = k;
    for (int g = out.ngroups; g <= MAXHF_FIT; ++g) out.gstart[g] = out.nhf;
    // Planck terms about the band centre
    out.ic0 = 0.5 * (in.nchan - 1);
    const double vc = in.v0 + in.dv * out.ic0;
    out.T0c = H_CGS * nu_ref * 1e9 * (1.0 - vc / CKMS) / KB_CGS;
    out.dT0 = -H_CGS * nu_ref * 1e9 * in.dv / (CKMS * KB_CGS);
    double J0, J1, JT0, JT1;
    planck_terms(TCMB, out.T0c, out.dT0, J0, J1, JT0, JT1);
    out.Jbg0 = J0;
    out.Jbg1 = J1;
    return 0;
}

}  // namespace b200fit
