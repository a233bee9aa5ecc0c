"""The moment / signal-mask pre-pass of cubefit_gen on the GPU: signals.get_moments (mufasa/signals.py:329-396)
composed from the bit-cube primitives of libb200fit.so (include/b200fit.h, csrc/b200fit_prepass.cu).

Same sequence as the numpy mirror in signals.py -- robust rms, 3 sigma signal cube, v_estimate (erosion with ball(2),
velocity at the peak), optional central window around the mode of the peak velocities (N2H+), opening, velocity
window, disk(3) dilation, moment0/1/2 -- but the boolean cubes never leave the device and every step is one kernel
over resident rows.  Only the 25-bin mode estimate of the peak-velocity MAP runs on the host (a [ny, nx] array).
"""
import ctypes as C

import numpy as np
import torch

MORPH_ERODE_BALL2, MORPH_ERODE_CROSS, MORPH_DILATE_CROSS, MORPH_DILATE_DISK3 = 0, 1, 2, 3


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class BitCubeOps(object):
    """Thin wrappers; all tensors live on ``engine.device``.  rows: [ny*nx, stride] fp32 (full plane, row-major)."""

    def __init__(self, engine, rows, nchan, ny, nx, v0, dv):
        assert rows.shape[0] == ny * nx
        self.eng, self.rows = engine, rows
        self.nchan, self.ny, self.nx, self.v0, self.dv = int(nchan), int(ny), int(nx), float(v0), float(dv)
        self.stride = int(rows.shape[1])
        self.W = self.stride // 32
        self.npix = ny * nx

    def _bits(self):
        return torch.empty((self.npix, self.W), dtype=torch.int32, device=self.eng.device)

    def signal_bits(self, rms, snr_cut):
        out = self._bits()
        e = self.eng
        e._check(e.lib.b200fit_signal_bits(e.ctx, self.nchan, self.npix, _p(self.rows), self.stride, _p(rms),
                                           float(snr_cut), _p(out), e._stream()), "b200fit_signal_bits")
        return out

    def morph(self, op, bits):
        out = self._bits()
        e = self.eng
        e._check(e.lib.b200fit_bits_morph(e.ctx, int(op), self.nchan, self.ny, self.nx, self.stride, _p(bits), _p(out),
                                          e._stream()), "b200fit_bits_morph")
        return out

    def opening(self, bits):
        return self.morph(MORPH_DILATE_CROSS, self.morph(MORPH_ERODE_CROSS, bits))

    def window(self, bits, vcenter, hwidth):
        out = self._bits()
        e = self.eng
        per_pixel = isinstance(vcenter, torch.Tensor)
        e._check(e.lib.b200fit_bits_window(e.ctx, self.nchan, self.npix, self.stride, self.v0, self.dv, _p(bits),
                                           _p(vcenter) if per_pixel else None,
                                           0.0 if per_pixel else float(vcenter), float(hwidth), _p(out), e._stream()),
                 "b200fit_bits_window")
        return out

    def v_at_peak(self, bits_a, bits_b=None, planemask=None):
        out = torch.empty(self.npix, dtype=torch.float64, device=self.eng.device)
        e = self.eng
        e._check(e.lib.b200fit_v_at_peak(e.ctx, self.nchan, self.npix, _p(self.rows), self.stride, _p(bits_a),
                                         _p(bits_b), _p(planemask), self.v0, self.dv, _p(out), e._stream()),
                 "b200fit_v_at_peak")
        return out

    def moments(self, bits, planemask=None):
        m = [torch.empty(self.npix, dtype=torch.float64, device=self.eng.device) for _ in range(3)]
        e = self.eng
        e._check(e.lib.b200fit_masked_moments(e.ctx, self.nchan, self.npix, _p(self.rows), self.stride, _p(bits),
                                              _p(planemask), self.v0, self.dv, _p(m[0]), _p(m[1]), _p(m[2]),
                                              e._stream()), "b200fit_masked_moments")
        return m


def estimate_mode(data, bins=25):
    """signals.estimate_mode (:399-420) on a map."""
    data = data[np.isfinite(data)]
    counts, edges = np.histogram(data, bins=bins)
    k = np.argmax(counts)
    return (edges[k] + edges[k + 1]) / 2


def get_moments(pcube, planemask, window_hwidth=5, linewidth_sigma=True, central_win_hwidth=None, return_rms=False):
    """signals.get_moments for the cube of ``pcube`` with the include-mask ``isfinite(cube) & planemask[ny, nx]``
    (what trim_cube_edge produces).  Returns (m0, m1, m2[, rms]) as [ny, nx] float64 maps."""
    snr_cut, noise_mask_cut = 3, 2.5
    # the 1- and the 2-component driver of one UltraCube ask for the same maps: cached with the device rows
    key = (np.asarray(planemask, bool).tobytes(), float(window_hwidth), bool(linewidth_sigma),
           None if central_win_hwidth is None else float(central_win_hwidth))
    cached = pcube._dev.get("prepass_moments")
    if cached is not None and cached[0] == key:
        out = cached[1]
        return out if return_rms else out[:3]
    eng = pcube.engine
    nchan, ny, nx = pcube.cube.shape
    rows = pcube.rows_on_device()
    v0, dv, _ = pcube.xarr.v0_dv_flip
    ops = BitCubeOps(eng, rows, nchan, ny, nx, v0, dv)
    pm = torch.from_numpy(np.ascontiguousarray(planemask, dtype=np.uint8).ravel()).to(eng.device)
    noise = eng.rms_snr(nchan, rows, planemask=pm, sigma_cut=noise_mask_cut, expand=20)
    rms = noise["rms"]
    sig = ops.signal_bits(rms, snr_cut)                                   # data > 3 rms (raw data)
    eroded = ops.morph(MORPH_ERODE_BALL2, sig)
    v_peak = ops.v_at_peak(eroded, None, pm)                              # v_estimate
    opened = None
    if central_win_hwidth is not None:
        v_mode = estimate_mode(v_peak.cpu().numpy(), bins=25)
        opened = ops.opening(sig)
        tmp = ops.window(opened, float(v_mode), central_win_hwidth)       # get_v_mask about the mode
        v_peak = ops.v_at_peak(eroded, tmp, pm)
    if opened is None:
        opened = ops.opening(sig)
    signal = ops.window(opened, v_peak, window_hwidth)                    # get_v_mask about each pixel's peak
    signal = ops.morph(MORPH_DILATE_DISK3, signal)
    m0, m1, m2 = ops.moments(signal, pm)
    if linewidth_sigma:
        m2 = torch.sqrt(m2)
    packed = torch.stack([m0, m1, m2, rms, noise["peak"]]).cpu().numpy().reshape(5, ny, nx)
    pcube._dev["prepass_peak"] = packed[4]
    pcube._dev["prepass_moments"] = (key, (packed[0], packed[1], packed[2], packed[3]))
    out = pcube._dev["prepass_moments"][1]
    return out if return_rms else out[:3]
