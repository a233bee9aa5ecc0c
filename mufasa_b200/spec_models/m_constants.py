"""Molecular constants on the host side (mirror of mufasa/spec_models/m_constants.py).

The arithmetic copies of these tables live in the CUDA library (csrc/b200fit_common.h); this module only serves
host logic that needs them (rest frequency, velocity offset of the brightest hyperfine, moment windows).
N2H+ (1-0): m_constants.py:22-61.  NH3 (1,1): pyspeckit ammonia_constants 'oneone' (imported by the reference
at m_constants.py:11-17, not part of its tree).
"""

nh3_constants = dict(
    line_names=['oneone'],
    freq_dict={'oneone': 23.6944955e9},
    voff_lines_dict={'oneone': [19.8513, 19.3159, 7.88669, 7.46967, 7.35132, 0.460409, 0.322042,
                                -0.0751680, -0.213003, 0.311034, 0.192266, -0.132382, -0.250923,
                                -7.23349, -7.37280, -7.81526, -19.4117, -19.5500]},
    tau_wts_dict={'oneone': [0.0740740, 0.148148, 0.0925930, 0.166667, 0.0185190, 0.0370370,
                             0.0185190, 0.0185190, 0.0925930, 0.0333330, 0.300000, 0.466667,
                             0.0333330, 0.0925930, 0.0185190, 0.166667, 0.0740740, 0.148148]},
)

n2hp_constants = dict(
    line_names=['onezero'],
    freq_dict={'onezero': 93173.7637e6},
    voff_lines_dict={'onezero': [-7.9930, -7.9930, -7.9930, -0.6112, -0.6112, -0.6112, 0.0000, 0.9533,
                                 0.9533, 5.5371, 5.5371, 5.5371, 5.9704, 5.9704, 6.9238]},
    tau_wts_dict={'onezero': [0.025957, 0.065372, 0.019779, 0.004376, 0.034890, 0.071844, 0.259259,
                              0.156480, 0.028705, 0.041361, 0.013309, 0.056442, 0.156482, 0.028705,
                              0.037038]},
)
