"""Drop-in replacement at the reference's import path.

The reference's notebooks do `from src.mshds_extractor import extract_mshds_features`
(/root/reference/notebooks/01_feature_extraction_setup.ipynb:31); `src/` is a namespace package there, so placing this
file on sys.path ahead of the reference's is all a maintainer has to do.  Same function name, same
arguments, same 26-column DataFrame -- computed by the B200 CUDA library instead of Praat.
"""
from robust_speech_analysis_framework_b200.mshds_extractor import (  # noqa: F401
    FEATURE_NAMES,
    extract_mshds_features,
    extract_mshds_from_pcm,
)
