"""The kernel's numerical algorithm (shared header b200fit_common.h: velocity-space fp32 model + analytic Jacobian,
windowing, fp64 normal equations, LM state machine) compiled for the host and checked against the CPU oracle.
This is a TEST of the shared source, not a product path: libb200fit.so contains no host implementation."""
import numpy as np

import _tiles as T
from mufasa_b200 import _abi, synth
from oracle import model as M


def test_emulated_evaluation_matches_fp64_oracle():
    """chi2, J^T r and J^T J of one evaluation vs the oracle's fp64 velocity-space model + analytic Jacobian."""
    import ctypes as C
    lib = T.build_emul()
    tile = T.sample_tile(2, nrows=1, xstep=64)
    lo, hi = T.limits_vectors(tile["fittype"], 2)
    g = synth.clamp_guesses(tile["guess2"], lo, hi)
    spec, gg = T.pixel_major(tile, g)
    prob = _abi.make_problem(tile["fittype"], 2, tile["nchan"], 1, tile["v"][0], tile["dv"], lo, hi)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    for pix in range(spec.shape[0]):
        H, b, chi2 = np.zeros(36), np.zeros(8), np.zeros(1)
        assert lib.emul_eval(C.byref(prob), vp(np.ascontiguousarray(spec[pix])), vp(np.ascontiguousarray(gg[pix])),
                             vp(H), vp(b), vp(chi2)) == 0
        m, J = M.model_and_jacobian_vspace(tile["v"], gg[pix], tile["fittype"])
        y = spec[pix, :tile["nchan"]].astype(np.float64)
        r = y - m
        assert abs(chi2[0] - np.sum(r * r)) <= 2e-6 * np.sum(r * r)          # fp32 model, fp64 sums
        bo = J @ r
        Ho = J @ J.T
        scale = np.sqrt(np.diag(Ho))
        assert np.max(np.abs(b - bo) / (scale * np.sqrt(np.sum(r * r)))) < 2e-6
        q = 0
        for i in range(8):
            for j in range(i, 8):
                assert abs(H[q] - Ho[i, j]) <= 5e-6 * scale[i] * scale[j]
                q += 1


def test_emulated_fit_matches_oracle_1c():
    lib = T.build_emul()
    tile = T.sample_tile(2, nrows=2, xstep=16)
    lo, hi = T.limits_vectors(tile["fittype"], 1)
    g = synth.clamp_guesses(tile["guess1"], lo, hi)
    o = T.oracle_fit(tile, 1, lo, hi, g, cores=2)
    spec, gg = T.pixel_major(tile, g)
    prob = _abi.make_problem(tile["fittype"], 1, tile["nchan"], spec.shape[0], tile["v"][0], tile["dv"], lo, hi)
    em = T.emul_fit(lib, prob, spec, gg)
    c = T.criteria(em["params"], em["chi2"], em["status"], o, 4)
    assert c["both"].mean() > 0.9
    ok = (c["par_ok"] & c["chi2_ok"])[c["both"]]
    assert ok.mean() >= 0.95
    assert np.median(c["mz"][c["both"]]) < 5e-3
