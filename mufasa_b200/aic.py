"""AIC / AICc / relative likelihood on maps: host mirror of mufasa/aic.py:136-201 (the same arithmetic runs per pixel
inside the CUDA ``aicc`` kernel; these serve callers that hold rss / N maps already)."""
import numpy as np


def AIC(rss, p, N):
    N = np.asarray(N, dtype=float)
    N[N == 0] = np.nan          # in place, like the reference (aic.py:155-157)
    return N * np.log(rss / N) + 2 * p


def AICc(rss, p, N):
    top = 2 * p * (p + 1)
    bottom = N - p - 1
    return AIC(rss, p, N) + top / bottom


def likelihood(aiccA, aiccB):
    """ln(L_A / L_B) = -(AICc_A - AICc_B) / 2."""
    return -1.0 * (aiccA - aiccB) / 2.0
