"""B200-native drop-in for the reference's BPtrain_Sigmoid training step and Wav2LPS_be LPS front end.

The product is the C-ABI shared library ``libggd_b200.so`` (headers in ``include/``), built from
``csrc/`` for sm_100a.  This package only mirrors the reference's operator interface on top of it
(``BP_GPU`` = Train_code_ML_GGD/BP_GPU.h:45-70, ``Wav2LPS`` = Wav2LPS_be) through ctypes; there is no
CPU fallback -- every entry point raises when the library or an sm_100 GPU is missing.
"""
from .bp_gpu import BP_GPU, GGDError, load_library, library_path, FLAG_UNFUSED_UPDATE, FLAG_NO_GRAPH, \
    FLAG_KEEP_DEBUG, FLAG_PIN_HOST, PREC_BF16X3, PREC_FP32_SIMT  # noqa: F401
from .wav2lps import Wav2LPS, LPSError, lps_nframes, FLAG_BIG_ENDIAN, FLAG_ZSCORE, FLAG_EXACT, \
    FLAG_PFILE, FLAG_ACCUM_NORM  # noqa: F401
