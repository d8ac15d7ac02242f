"""ac_dsp_b200 -- B200-native (sm_100a) engine for the FIR / CIC hot path of hlslibs/ac_dsp.

The product is the C-ABI shared library (include/b200dsp.h, ac_dsp_b200/csrc/); this package is the thin
host-side mirror of the reference's class templates.  Importing it never touches the CPU oracle.
"""
from ._lib import B2dError, FTYPES, INTERLEAVED, PLANAR, load, lib_path  # noqa: F401
from .filters import (Comm, ac_cic_dec_full, ac_cic_intr_full, ac_fir_const_coeffs, ac_fir_load_coeffs,  # noqa: F401
                      ac_fir_prog_coeffs, ac_fir_reg_share, ac_fixed, ac_intg_dump, ac_mv_avg, ac_poly_dec, ac_poly_intr, cic_intr_fir_cascade, shard_channels)
