"""MetaModel: per-fittype information the fit drivers need (mirror of mufasa/spec_models/meta_model.py:67-150).

Same attribute names, values and error behaviour; the ``fitter`` it carries is a light descriptor (npars,
parnames) because the model itself is evaluated by the CUDA library, not by a Python callable."""
import numpy as np

from ..exceptions import FitTypeError
from . import m_constants as mc

# deprecated module names the reference still forwards (ammonia_multiv.py:24-30, n2hp_multiv.py:23-29)
FITTYPE_ALIASES = {"ammonia_multiv": "nh3_multi_v", "n2hp_multiv": "n2hp_multi_v"}


class FitterInfo(object):
    """Stands in for the pyspeckit SpectralModel built by BaseModel.multi_v_model_generator (BaseModel.py:43-81):
    parameter count/order [vlsr, sigma, tex, tau] x ncomp."""

    def __init__(self, fittype, n_comp):
        self.fittype = fittype
        self.npars = 4 * n_comp
        self.parnames = [x for ln in range(n_comp) for x in
                         ('vlsr%d' % ln, 'sigma%d' % ln, 'tex%d' % ln, 'tau%d' % ln)]
        self.fitunit = 'Hz'


class MetaModel(object):
    def __init__(self, fittype, ncomp):
        fittype = FITTYPE_ALIASES.get(fittype, fittype)
        self.fittype = fittype
        self.ncomp = ncomp
        self.adoptive_snr_masking = True
        self.window_hwidth = 5
        self.central_win_hwidth = None
        self.Texmin = 3.0
        self.Texmax = 40
        self.sigmin = 0.07
        self.sigmax = 2.5
        self.taumax = 30.0
        self.taumin = 0.1
        self.eps = 0.001
        if self.fittype == 'nh3_multi_v':
            consts = mc.nh3_constants
            self.linetype = 'nh3'
            self.linename = 'oneone'
        elif self.fittype == 'n2hp_multi_v':
            consts = mc.n2hp_constants
            self.linetype = 'n2hp'
            self.linename = 'onezero'
            self.taumax = 40.0
            self.window_hwidth = 4
            self.central_win_hwidth = 3.5
        else:
            raise FitTypeError("\'{}\' is an invalid fittype".format(fittype))
        self.fitter = FitterInfo(self.fittype, ncomp)
        self.linenames = [self.linename]
        self.rest_value = consts['freq_dict'][self.linename]          # Hz (the reference carries an astropy Quantity)
        self.voffs = consts['voff_lines_dict'][self.linename]
        self.tau_wts = consts['tau_wts_dict'][self.linename]
        self.voff_at_peak = self.voffs[int(np.argmax(self.tau_wts))]

    def limits(self, vmin, vmax):
        """minpars / maxpars vectors in the order fiteach expects (multi_v_fit.py:376-379)."""
        lo = [vmin, self.sigmin, self.Texmin, self.taumin] * self.ncomp
        hi = [vmax, self.sigmax, self.Texmax, self.taumax] * self.ncomp
        return lo, hi
