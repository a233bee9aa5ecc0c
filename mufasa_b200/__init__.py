"""mufasa_b200 -- B200-native per-pixel spectral fitter behind MUFASA's fit_cube / fiteach interface."""
__version__ = "0.1.0"
