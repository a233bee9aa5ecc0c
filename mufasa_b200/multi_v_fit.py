"""Fit drivers: host-side mirror of mufasa/multi_v_fit.py for the B200 fitter.

Same function names, argument meaning and error behaviour as the reference drivers
  * register_pcube   mufasa/multi_v_fit.py:96-123
  * cubefit_simp     mufasa/multi_v_fit.py:126-208   ("trust my guesses" -- used by every refit)
  * cubefit_gen      mufasa/multi_v_fit.py:211-417   (moments -> guesses -> clamps -> masks -> fiteach)
  * retry_fit / get_start_point / get_vstats / handle_snr / snr_estimate   :456-478, :634-738
but the per-pixel work behind ``pcube.fiteach`` is one launch of the CUDA library (mufasa_b200/pcube.py).
Everything here is per-call control logic on [ny, nx] maps: it stays on the host, as in the reference.
"""
import logging

import numpy as np

from . import guess_refine as g_refine
from . import moment_guess as momgue
from .exceptions import SNRMaskError, StartFitError
from .spec_models.meta_model import MetaModel

logger = logging.getLogger(__name__)


def validate_n_cores(n_cores):
    """utils/multicore.py:5-18 -- kept for API compatibility; the GPU path ignores it."""
    import os
    if n_cores is True or n_cores is None:
        return max(1, (os.cpu_count() or 2) - 1)
    if n_cores is False:
        return 1
    return max(1, min(int(n_cores), os.cpu_count() or 1))


def register_pcube(pcube, mod_info):
    """multi_v_fit.py:96-123: unit check, rest frequency default, radio convention, fitter registration."""
    if hasattr(pcube, "unit"):
        if pcube.unit is None or pcube.unit == '':
            logger.warning("The cube doesn't have a unit. The fitting will proceed assume it has a unit of K")
            pcube.unit = 'K'
        elif pcube.unit != 'K':
            raise ValueError(f"The cube should have units of K rather than {pcube.unit}")
    else:
        logger.warning("The cube doesn't have a unit. The fitting will proceed assume it has a unit of K")
    if pcube.xarr.refX is None or not np.isfinite(pcube.xarr.refX):
        logger.debug("The cube has no reference rest frequency. The rest frequency of the spectral model will be "
                     "used instead")
        pcube.xarr.refX = mod_info.rest_value
    if pcube.xarr.velocity_convention is None:
        pcube.xarr.velocity_convention = 'radio'
    fitter = mod_info.fitter
    pcube.specfit.Registry.add_fitter(mod_info.fittype, fitter, fitter.npars)
    return pcube


def get_vstats(velocities, signal_mask=None):
    """multi_v_fit.py:634-644."""
    m1 = velocities
    mask = np.isfinite(m1)
    if signal_mask is not None:
        mask = np.logical_and(mask, signal_mask)
    m1_good = m1[mask]
    # np.median is the mean of the two middle values; the 50th percentile's lerp can differ from it in the last bit
    v_median = np.median(m1_good)
    v_99p, v_1p = np.percentile(m1_good, [99, 1])
    return v_median, v_99p, v_1p


def get_start_point(maskmap, weight=None, return_xy=True):
    """multi_v_fit.py:647-674.  (The start point does not influence results: use_neighbor_as_guess=False.)"""
    if weight is not None:
        weight = np.nan_to_num(np.asarray(weight) * maskmap)
        try:
            start_from_point = np.unravel_index(weight.argmax(), weight.shape)
            if not maskmap[start_from_point]:
                raise IndexError("The start point is not within maskmap.")
        except IndexError:
            indx_g = np.argwhere(maskmap)
            start_from_point = (indx_g[0, 0], indx_g[0, 1])
    else:
        indx_g = np.argwhere(maskmap)
        start_from_point = (indx_g[0, 0], indx_g[0, 1])
    if return_xy:
        return (int(start_from_point[1]), int(start_from_point[0]))
    return (int(start_from_point[0]), int(start_from_point[1]))


def snr_estimate(pcube, errmap, smooth=True):
    """multi_v_fit.py:677-702: errmap = mad_std(cube, axis=0); peaksnr = nanmax(cube / errmap), smoothed by a
    Gaussian2DKernel(1)."""
    from . import signals
    cube = pcube.cube if hasattr(pcube, 'cube') else pcube
    peak = None
    cache = getattr(pcube, "_dev", None) if errmap is None else None
    if cache is not None and ("snr_estimate", bool(smooth)) in cache:
        # the maps depend on the data cube only, and the PCubes of a refit share their cube's device state: the
        # ~40 masked refits of fit_surroundings each smoothed the same 512 x 512 map again (7 ms apiece)
        peaksnr, errmap, err_mask = cache[("snr_estimate", bool(smooth))]
        return peaksnr.copy(), None, errmap, err_mask
    if errmap is None:
        # per-pixel MAD + peak on the GPU (b200fit_rms_snr): mad_std(cube, axis=0) and nanmax(cube, axis=0)
        r = pcube.rms_robust()
        errmap, peak = r["rms0"], r["peak"]
    err_med = np.nanmedian(errmap)
    err_mask = errmap < err_med * 3.0
    with np.errstate(all="ignore"):
        if peak is None:
            peak = np.nanmax(cube, axis=0)
        peaksnr = peak / errmap      # == nanmax(cube / errmap) for errmap > 0
    if smooth:
        peaksnr = signals.convolve_gauss2d(peaksnr, 1.0)
    if cache is not None:
        cache[("snr_estimate", bool(smooth))] = (peaksnr.copy(), errmap, err_mask)
    return peaksnr, None, errmap, err_mask


def handle_snr(pcube, snr_min, planemask, return_errmask=False, **kwargs):
    """multi_v_fit.py:705-738, including its quirk: with snr_min=None and no signal_cut the fiteach default
    signal_cut (3) stays in force."""
    if snr_min is None:
        snr_min = 0.0 if 'signal_cut' not in kwargs else kwargs['signal_cut']
    elif 'signal_cut' in kwargs:
        if kwargs['signal_cut'] != snr_min:
            if kwargs['signal_cut'] > snr_min:
                snr_min = kwargs['signal_cut']
            else:
                kwargs['signal_cut'] = snr_min
            logger.warning("snr_min and signal_cut provided do not match. The higher value of the two, {}, is "
                           "adopted".format(snr_min))
    else:
        kwargs['signal_cut'] = 0
    peaksnr, err_mask = None, None
    if snr_min > 0 or return_errmask:
        # the reference always computes the map; it only matters when snr_min > 0 (full-cube mad_std is costly)
        peaksnr, _, _, err_mask = snr_estimate(pcube, kwargs.get('errmap', None), smooth=True)
    if snr_min > 0:
        snr_mask = peaksnr > snr_min
        planemask = np.logical_and(planemask, snr_mask)
    if planemask.sum() < 1:
        msg = "The provided snr_min={} results in no valid pixels to fit; fitting terminated.".format(snr_min)
        raise SNRMaskError(msg)
    if return_errmask:
        return peaksnr, planemask, kwargs, err_mask
    return peaksnr, planemask, kwargs


def _cube_data(cube):
    return cube._data if hasattr(cube, '_data') else np.asarray(cube)


def cubefit_simp(cube, pcube, ncomp, guesses, multicore=None, maskmap=None, fittype='nh3_multi_v', snr_min=None,
                 **kwargs):
    """multi_v_fit.py:126-208.  Mutates ``guesses`` in place exactly like the reference (zero velocities -> NaN,
    clamps with +-eps)."""
    for _ in cubefit_simp_steps(cube, pcube, ncomp, guesses, multicore=multicore, maskmap=maskmap, fittype=fittype,
                                snr_min=snr_min, **kwargs):
        pass
    return pcube


def cubefit_simp_steps(cube, pcube, ncomp, guesses, multicore=None, maskmap=None, fittype='nh3_multi_v', snr_min=None,
                       **kwargs):
    """``cubefit_simp`` as a two-step generator: everything up to the ``yield`` is host work on the guesses that does
    not need the cube on the device; the rest (footprint, pixel selection, the fit) does.  A caller preparing several
    fits of one cube runs the first half of all of them while the cube is still uploading (UltraCube.fit_cube)."""
    mod_info = MetaModel(fittype, ncomp)
    Texmin, Texmax = mod_info.Texmin, mod_info.Texmax
    sigmin, sigmax = mod_info.sigmin, mod_info.sigmax
    taumin, taumax = mod_info.taumin, mod_info.taumax
    eps = mod_info.eps

    register_pcube(pcube, mod_info)
    multicore = validate_n_cores(multicore)

    data = _cube_data(cube)
    if hasattr(pcube, "rows_on_device"):
        # start the (asynchronous) upload + layout pass of the first slab; host logic below overlaps it, and the other
        # slabs are queued behind the guesses (PCube.fiteach): host-to-device copies are served first in, first out
        pcube.rows_on_device(wait=False, max_slabs=1)
    planemask = np.any(np.isfinite(guesses), axis=0)
    peaksnr, planemask, kwargs = handle_snr(pcube, snr_min, planemask, **kwargs)

    v_peak_hwidth = 10
    v_guess = guesses[::4]
    v_guess[v_guess == 0] = np.nan
    # (median, 99th, 1st percentile) of the velocity guesses: a GLOBAL quantity -- a partitioned run passes the
    # statistics of the whole map (mufasa_b200/dist.py), SURVEY.md 7.5
    v_median, v_99p, v_1p = kwargs.pop('vstats', None) or get_vstats(v_guess)
    vmax = v_median + v_peak_hwidth
    vmin = v_median - v_peak_hwidth
    if v_99p + sigmax > vmax:
        vmax = v_99p + sigmax
    if v_1p - sigmax < vmin:
        vmin = v_1p - sigmax

    def impose_lim(data, min=None, max=None, eps=0):
        if min is not None:
            mask = data < min
            data[mask] = min + eps
        if max is not None:
            mask = data > max
            data[mask] = max - eps

    impose_lim(guesses[::4], vmin, vmax, eps)
    impose_lim(guesses[1::4], sigmin, sigmax, eps)
    impose_lim(guesses[2::4], Texmin, Texmax, eps)
    impose_lim(guesses[3::4], taumin, taumax, eps)
    yield

    # The reference also removes the pixels without a finite channel here (footprint_mask, multi_v_fit.py:150, :193).
    # On the device that would be a host synchronisation with the cube upload; the fit kernel returns "not fitted"
    # (zeros, has_fit False) for such spectra by itself, so they are left in the mask and the maps come out the same.
    if hasattr(pcube, "rows_on_device"):
        footprint_mask = True
    else:
        footprint_mask = np.any(np.isfinite(data), axis=0)
    if maskmap is not None:
        maskmap = maskmap & planemask & footprint_mask
    else:
        maskmap = planemask & footprint_mask

    kwargs['start_from_point'] = get_start_point(maskmap, weight=None)
    kwargs['multicore'] = multicore
    kwargs['maskmap'] = maskmap
    kwargs['use_neighbor_as_guess'] = False
    kwargs['limitedmax'] = [True, True, True, True] * ncomp
    kwargs['maxpars'] = [vmax, sigmax, Texmax, taumax] * ncomp
    kwargs['limitedmin'] = [True, True, True, True] * ncomp
    kwargs['minpars'] = [vmin, sigmin, Texmin, taumin] * ncomp
    kwargs.setdefault('verbose_level', 0)
    kwargs.setdefault('verbose', False)

    pcube.fiteach(fittype=fittype, guesses=guesses, **kwargs)
    yield


def cubefit_gen(cube, pcube, ncomp=2, paraname=None, modname=None, chisqname=None, guesses=None, errmap11name=None,
                multicore=None, mask_function=None, snr_min=0.0, momedgetrim=True, saveguess=False,
                fittype='nh3_multi_v', **kwargs):
    """multi_v_fit.py:211-417: moment maps -> guesses -> tautex_renorm -> clamps -> masks -> fiteach.

    The moment / rms / S/N pre-pass (signals.get_moments, get_snr) runs on the GPU (mufasa_b200/prepass.py); there
    is no host path for it (its numpy restatement is test infrastructure, oracle/signals.py).

    ``paraname`` / ``modname`` / ``chisqname`` (FITS writers) are outside the fit path and not supported here."""
    from . import signals
    from .utils import interpolate

    mod_info = MetaModel(fittype, ncomp)
    nu0_ghz = mod_info.rest_value / 1e9
    Texmin, Texmax = mod_info.Texmin, mod_info.Texmax
    sigmin, sigmax = mod_info.sigmin, mod_info.sigmax
    taumin, taumax = mod_info.taumin, mod_info.taumax
    eps = mod_info.eps

    register_pcube(pcube, mod_info)

    linewidth_sigma = False
    trim = 3
    # the whole pre-pass on the GPU (mufasa_b200/prepass.py): the include-mask of the trimmed cube is
    # isfinite(cube) & cube_plane, so only the 2-D footprint is trimmed on the host (signals.trim_cube_edge :199-226)
    from . import prepass
    cube_plane = pcube.footprint().copy()
    if np.sum(cube_plane) > 1000:
        cube_plane = signals.trim_edge(cube_plane, trim=trim)
    m0, m1, m2, rms = prepass.get_moments(pcube, cube_plane, window_hwidth=mod_info.window_hwidth,
                                          linewidth_sigma=linewidth_sigma,
                                          central_win_hwidth=mod_info.central_win_hwidth, return_rms=True)
    with np.errstate(all="ignore"):
        peaksnr = pcube._dev["prepass_peak"] / rms

    vmax = np.nanmax(m1) + sigmax / 2
    vmin = np.nanmin(m1) - sigmax / 2

    if 'planemask' in kwargs:
        planemask = kwargs.pop('planemask') & cube_plane
    else:
        planemask = cube_plane
    if 'maskmap' in kwargs:
        planemask = np.logical_and(kwargs['maskmap'], planemask)
    if 'signal_cut' in kwargs:
        logger.warning("signal_cut will be set to zero for pcube.fiteach() in cubefit_gen. The snr_min is used to "
                       "determine which pixels to fit prior to calling fiteach()")
    if snr_min > 0:
        with np.errstate(invalid="ignore"):
            snr_mask = peaksnr > snr_min
        if snr_min >= 3:
            planemask = planemask & snr_mask & np.isfinite(m1)
        else:
            planemask = planemask & snr_mask
    if planemask.sum() < 1:
        raise SNRMaskError("The provided snr_min={} results in no valid pixels to fit; fitting "
                           "terminated".format(snr_min))

    mmm = interpolate.iter_expand(np.array([m0, m1, m2]), mask=planemask)
    m0, m1, m2 = mmm[0], mmm[1], mmm[2]

    gg = momgue.moment_guesses(m1, m2, ncomp, sigmin=sigmin, moment0=m0, linetype=mod_info.linetype,
                               mom0_floor=None)
    if guesses is None:
        guesses = gg
    else:
        guesses[guesses == 0] = np.nan
        gmask = np.isfinite(guesses)
        mom_mask = np.isfinite(gg)
        guesses_smooth = guesses.copy()
        for i in range(guesses_smooth.shape[0]):
            guesses_smooth[i] = signals.convolve_gauss2d(guesses[i], 5.0, boundary='extend')
        guesses[~gmask] = guesses_smooth[~gmask]
        guesses[~mom_mask] = np.nan
        gmask = np.isfinite(guesses)
        guesses[~gmask] = gg[~gmask]
        with np.errstate(invalid="ignore"):
            gmask = guesses[1::4] < sigmin
        guesses[1::4][gmask] = guesses_smooth[1::4][gmask]

    guesses[3::4], guesses[2::4] = g_refine.tautex_renorm(guesses[3::4], guesses[2::4], tau_thresh=taumin,
                                                          nu=nu0_ghz)
    with np.errstate(invalid="ignore"):
        guesses[::4][guesses[::4] > vmax] = vmax
        guesses[::4][guesses[::4] < vmin] = vmin
        guesses[1::4][guesses[1::4] > sigmax] = sigmax
        guesses[1::4][guesses[1::4] < sigmin] = sigmin + eps
        guesses[2::4][guesses[2::4] > Texmax] = Texmax
        guesses[2::4][guesses[2::4] < Texmin] = Texmin
        guesses[3::4][guesses[3::4] > taumax] = taumax
        guesses[3::4][guesses[3::4] < taumin] = taumin

    kwargs['maskmap'] = planemask & cube_plane & np.all(np.isfinite(guesses), axis=0)
    kwargs['maskmap'] = kwargs['maskmap'] & match_pcube_mask(pcube, kwargs['maskmap'])
    mask_size = np.sum(kwargs['maskmap'])
    if mask_size < 1:
        logger.warning("maskmap has no pixels, no fitting will be performed")
        return pcube
    elif np.all(np.isfinite(guesses), axis=0).sum() < 1:
        logger.warning("guesses has no pixels, no fitting will be performed")
        return pcube

    multicore = validate_n_cores(multicore)
    start_point = get_start_point(kwargs['maskmap'], weight=peaksnr)
    kwargs['start_from_point'] = start_point
    kwargs_core = dict(fittype=fittype, guesses=guesses,
                       limitedmax=[True, True, True, True] * ncomp, maxpars=[vmax, sigmax, Texmax, taumax] * ncomp,
                       limitedmin=[True, True, True, True] * ncomp, minpars=[vmin, sigmin, Texmin, taumin] * ncomp,
                       multicore=multicore, use_neighbor_as_guess=False, guesses_are_moments=False, signal_cut=0)
    kwargs = {**kwargs, **kwargs_core}
    kwargs.setdefault('verbose_level', 0)
    kwargs.setdefault('verbose', False)
    try:
        pcube.fiteach(**kwargs)
    except (ValueError, AssertionError) as e:
        logger.warning("The starting fit position is somehow not among the valid: " + str(e))
        retry_fit(pcube, kwargs, start_point, peaksnr=peaksnr)
    if paraname is not None:
        if getattr(pcube, "_deferred", None) is not None:
            pcube._save_after_fit = (paraname, ncomp)     # batched fit: the maps exist after finish_fit (UltraCube)
        else:
            save_pcube(pcube, paraname, ncomp=ncomp)
    if modname is not None or chisqname is not None:
        raise NotImplementedError("model-cube / legacy chi-square writers are not part of the fit path")
    return pcube


def make_header(ndim, ref_header):
    """multi_v_fit.py:481-520: a fresh header carrying the WCS / beam cards of ``ref_header`` (dict-like of
    keyword -> value | (value, comment), or None)."""
    from datetime import datetime
    from .utils import fits_io
    hdr = fits_io.Header()
    hdr.set('DATE', datetime.now().strftime('%Y-%m-%dT%H:%M:%S'), "Written by mufasa_b200")
    if ndim == 3:
        keylist = ['OBJECT', 'BMAJ', 'BMIN', 'BPA', 'WCSAXES', 'CRPIX1', 'CRPIX2', 'CRPIX3', 'CDELT1', 'CDELT2',
                   'CDELT3', 'CUNIT1', 'CUNIT2', 'CUNIT3', 'CTYPE1', 'CTYPE2', 'CTYPE3', 'CRVAL1', 'CRVAL2', 'CRVAL3',
                   'LONPOLE', 'LATPOLE', 'WCSNAME', 'MJDREF', 'RADESYS', 'EQUINOX', 'SPECSYS']
    else:
        keylist = ['OBJECT', 'BMAJ', 'BMIN', 'BPA', 'WCSAXES', 'CRPIX1', 'CRPIX2', 'CDELT1', 'CDELT2', 'CUNIT1',
                   'CUNIT2', 'CTYPE1', 'CTYPE2', 'CRVAL1', 'CRVAL2', 'LONPOLE', 'LATPOLE', 'WCSNAME', 'MJDREF',
                   'RADESYS', 'EQUINOX', 'SPECSYS']
    if ref_header is not None:
        for key in keylist:
            if key in ref_header:
                v = ref_header[key]
                val, com = v if isinstance(v, tuple) else (v, None)
                hdr.set(key, val, com)
    if 'WCSAXES' in hdr and hdr.value('WCSAXES') != ndim:
        hdr.set('WCSAXES', ndim, 'Number of coordinate axes')
    hdr.history.append("=" * 50)
    hdr.history.append("Written by mufasa_b200 on {}".format(datetime.now().strftime('%Y-%m-%d %H:%M:%S')))
    hdr.comments.append("=" * 50)
    hdr.comments.append("MUFASA citation: Chen et al. (2020), DOI: 10.3847/1538-4357/ab7378")
    return hdr


def save_pcube(pcube, savename, ncomp, npara=4, header_note=None):
    """multi_v_fit.py:523-582: ``concatenate([parcube, errcube])`` as one FITS image with the PLANEi cards that name
    the planes -- the file ``Region.load_fits`` / ``load_model_fit`` read back."""
    from .utils import fits_io
    hdr = make_header(ndim=3, ref_header=getattr(pcube, 'header', None))
    if header_note is None:
        header_note = 'Parameter maps derived from {}-comp model fits'.format(ncomp)
    hdr.set('NOTES', header_note)
    hdr.set('CDELT3', 1, '[unitless] Coordinate increment at reference point')
    hdr.set('CTYPE3', 'FITPAR', 'fitted parameters')
    hdr.set('CRVAL3', 0, '[unitless] Coordinate value at reference point')
    hdr.set('CRPIX3', 1, '[unitless] Pixel coordinate of reference point')
    hdr.set('CUNIT3', 'None', '[unitless] Units of coordinate increment and value')
    names = (('VELOCITY', '[km/s] vlsr'), ('SIGMA', '[km/s] velocity dispersion'),
             ('TEX', '[K] excitation temperature'), ('TAU', '[unitless] peak optical depth'))
    for err in (False, True):
        for i in range(ncomp):
            for k, (nm, com) in enumerate(names):
                plane = ((ncomp if err else 0) + i) * npara + k
                label = ('e' if err else '') + nm + ('' if ncomp == 1 else '_{}'.format(i + 1))
                note = ('estimated error of ' if err else '') + com + ('' if ncomp == 1 else
                                                                       ' of component {}'.format(i + 1))
                hdr.set('PLANE{}'.format(plane), label, note)
    fits_io.write(savename, np.concatenate([pcube.parcube, pcube.errcube]), hdr)
    logger.debug("{} saved.".format(savename))


def match_pcube_mask(pcube, maskmap):
    """multi_v_fit.py:439-453: pixels pyspeckit itself would accept (finite in its 2-D plane)."""
    plane = pcube.footprint() if hasattr(pcube, "footprint") else np.any(np.isfinite(pcube.cube), axis=0)
    return plane & np.asarray(maskmap, dtype=bool)


def retry_fit(pcube, kwargs, old_start_point, peaksnr=None):
    """multi_v_fit.py:456-478: drop the failed start pixel, restart from the first masked pixel, once."""
    kwargs['maskmap'][old_start_point[1], old_start_point[0]] = False
    if not kwargs['maskmap'].any():
        raise StartFitError("The first two attempts to fit a pixel did not yield a fit.")
    kwargs['start_from_point'] = get_start_point(kwargs['maskmap'], weight=None)
    try:
        pcube.fiteach(**kwargs)
    except (ValueError, AssertionError) as e:
        msg = "The first two attempts to fit a pixel did not yield a fit. Likely due to a lack of signal or poor " \
              "guesses. Message from the fitter: " + str(e)
        logger.error(msg)
        raise StartFitError(msg)
