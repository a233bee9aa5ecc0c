"""Header arithmetic of the regridding steps (mirror of mufasa/utils/fits_utils.py: downsample_header :11-60,
get_pixel_mapping :63-151).

astropy.wcs is absent from this image.  The two places the reference maps pixels between headers
(convolve_sky_byfactor's reproject, convolve_tools.py:72-79; guess_from_cnvpara's regrid, guess_refine.py:301-314)
always relate a header to its own ``downsample_header`` copy: same projection, same CRVAL, scaled CDELT and shifted
CRPIX.  For such a pair the world coordinates cancel and the map is linear in pixels -- that case is implemented,
anything else raises NotImplementedError.
"""
import numpy as np

_SAME = ("CTYPE1", "CTYPE2", "CRVAL1", "CRVAL2", "CROTA2", "PC1_1", "PC1_2", "PC2_1", "PC2_2", "CD1_2", "CD2_1")


def downsample_header(header, factor, axis):
    """Copy of ``header`` with CDELT{axis} * factor and CRPIX{axis} moved so that the sky position of the grid's
    outer edge is unchanged (FITS axis numbering)."""
    header = header.copy()
    cd, cp = 'CDELT{0:d}'.format(axis), 'CRPIX{0:d}'.format(axis)
    scalefactor = 1. / factor
    header[cp] = (header[cp] - 1) * scalefactor + scalefactor / 2. + 0.5
    header[cd] = header[cd] * factor
    return header


def linear_pixel_map(header_in, header_out):
    """((ay, by), (ax, bx)): 0-based pixel of ``header_in`` = a * (0-based pixel of ``header_out``) + b."""
    for k in _SAME:
        va, vb = header_in.get(k), header_out.get(k)
        if va != vb and not (isinstance(va, float) and isinstance(vb, float) and np.isclose(va, vb, rtol=1e-12, atol=0)):
            raise NotImplementedError("pixel mapping between different projections needs a WCS library "
                                      "(%s: %r vs %r)" % (k, va, vb))
    out = []
    for ax in (2, 1):
        a = header_out['CDELT%d' % ax] / header_in['CDELT%d' % ax]
        b = (header_in['CRPIX%d' % ax] - 1) - a * (header_out['CRPIX%d' % ax] - 1)
        out.append((float(a), float(b)))
    return tuple(out)


def get_pixel_mapping(header1, header2):
    """grid[2, n, m]: the (y, x) pixel coordinates in ``header1`` of every pixel of ``header2``."""
    (ay, by), (ax, bx) = linear_pixel_map(header1, header2)
    ny, nx = int(header2['NAXIS2']), int(header2['NAXIS1'])
    yy, xx = np.indices((ny, nx), dtype=np.float64)
    return np.array([ay * yy + by, ax * xx + bx])
