"""Error types at the fit boundary; same names and bases as mufasa/exceptions.py:6-29."""


class SNRMaskError(Exception):
    """SNR mask has no valid pixel (raised when snr_min leaves nothing to fit; multi_v_fit.py:280-282, :731-733)."""
    pass


class FitTypeError(LookupError):
    """The fittype is not one of 'nh3_multi_v' / 'n2hp_multi_v' (meta_model.py:135)."""
    pass


class StartFitError(Exception):
    """Fitting failed from the beginning (multi_v_fit.py:456-478)."""
    pass
