"""UltraCube.get_masked_moment (mufasa/UltraCube.py:1568-1655), orders 0 / 1 / 2, on the host (no GPU): the mask
recipe (model > 0, low-signal pixels borrow the channels of the strong ones, spectral dilation) and spectral_cube's
moment definitions, against a direct evaluation."""
import numpy as np
from scipy import ndimage as nd

from mufasa_b200.cube import SpecCube
from mufasa_b200.UltraCube import get_masked_moment


def test_masked_moments_orders_0_1_2():
    rng = np.random.default_rng(0)
    v = np.linspace(-5, 5, 101)
    cen = rng.normal(0, 1, (1, 4, 5))
    amp = rng.uniform(0.5, 2, (1, 4, 5))
    model = np.exp(-0.5 * ((v[:, None, None] - cen) / 0.5) ** 2) * amp
    model[model < 1e-3] = 0
    model[:, 0, 0] = np.nan                                    # no fit in this pixel
    data = (model + rng.normal(0, 0.01, model.shape)).astype(np.float32)
    data[:, 0, 0] = rng.normal(0, 0.01, v.size)
    data[10, 2, 2] = np.nan
    cube = SpecCube(data, v)
    expand = 4
    # the mask, restated
    with np.errstate(invalid="ignore"):
        mask = model > 0
        peak = np.nanmax(np.where(np.isfinite(model), model, -np.inf), axis=0)
        peak[~np.isfinite(peak)] = np.nan
        med = np.nanmedian(peak)
        high = peak > med
        spec = np.any(model > 0.1 * med, axis=(1, 2))
    low = mask.copy()
    low[:, high] = False
    low[spec, :] = True
    mask[:, ~high] = low[:, ~high]
    mask = nd.binary_dilation(mask, structure=np.ones((expand, 1, 1), bool)) & np.isfinite(data)
    w = np.where(mask, data, 0.0).astype(np.float64)
    s0 = w.sum(0)
    m0 = s0 * (v[1] - v[0])
    with np.errstate(all="ignore"):
        m1 = (w * v[:, None, None]).sum(0) / s0
        m2 = (w * (v[:, None, None] - m1) ** 2).sum(0) / s0
    for order, want in ((0, m0), (1, m1), (2, m2)):
        got = get_masked_moment(cube, model.copy(), order=order, expand=expand)
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
    # sanity: first moment ~ line centre, second ~ sigma^2 for the strong pixels
    ok = high & np.isfinite(m1)
    assert np.nanmax(np.abs(m1[ok] - cen[0][ok])) < 0.05
    assert np.nanmax(np.abs(m2[ok] - 0.25)) < 0.05
