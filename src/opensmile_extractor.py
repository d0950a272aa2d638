"""Drop-in replacement at the reference's import path.

The reference's notebooks do `from src.opensmile_extractor import extract_opensmile_features`
(/root/reference/src/opensmile_extractor.py:9 runs the external SMILExtract binary once per file).  Same function name and
arguments, same shape of result (feature columns, then 'filename'; files that fail are left out) -- computed by the B200 CUDA
library without the binary: 720 of the 911 columns of Androids.conf.
"""
from robust_speech_analysis_framework_b200.lld_extractor import (  # noqa: F401
    extract_lld_functionals,
    extract_opensmile_features,
    functional_names,
    parse_smile_config,
)
