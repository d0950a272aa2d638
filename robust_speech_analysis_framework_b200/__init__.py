"""B200-native (sm_100a) implementation of the MSHDS acoustic feature extractor hot path of
ayushpradhan-dev/robust-speech-analysis-framework (src/mshds_extractor.py).

Layout: csrc/ (hand-written CUDA kernels + the C ABI of include/mshds_b200.h), _lib.py (ctypes binding),
mshds_extractor.py (host mirror of the reference interface), sharding.py (one process per GPU, clip-level sharding),
synth.py (seeded synthetic speech for tests and benchmarks).
"""
from ._lib import FEATURE_NAMES, N_FEATURES, Extractor, MshdsError  # noqa: F401
from .mshds_extractor import extract_mshds_features, extract_mshds_from_pcm  # noqa: F401
