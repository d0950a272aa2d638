"""Drop-in replacement at the reference's import path for the clip -> session aggregation.

The notebooks do `from src.utils import aggregate_clip_features`
(/root/reference/notebooks/01_feature_extraction_setup.ipynb, call at :992).  Same name, arguments and columns; the reduction
runs in libmshds_b200.so.  The reference's other helper in this module (aggregate_sequence_features, wav2vec2 path) is out
of scope and not provided.
"""
from robust_speech_analysis_framework_b200.utils import aggregate_clip_features  # noqa: F401
