"""Host-side mirror of the reference's clip -> session aggregation (the step right after the extractor).

Reference: /root/reference/src/utils.py:7-58  aggregate_clip_features(clip_features_df, metadata_df) -> DataFrame
  * empty input: prints the warning of :32 and returns an empty DataFrame (:31-33)
  * inner merge of metadata[['filename', 'unique_participant_id']] with the clip features on 'filename' (:36-39), in
    metadata row order; 'filename' dropped (:42)
  * groupby('unique_participant_id').agg(['mean', 'std']) (:49): sessions sorted by id, NaNs skipped, std with ddof = 1
    (NaN for a single clip, cf. notebooks/02_model_evaluation.ipynb:153-155)
  * columns flattened to '<feature>_mean', '<feature>_std' in feature order (:53), id back as the first column (:56)

The merge and the id factorisation stay in pandas (string work); the numeric reduction is ONE call into
libmshds_b200.so (mshds_aggregate_sessions), which reproduces pandas' group_mean / group_var arithmetic bit for bit.
There is no CPU fallback: without the CUDA library / a CUDA device the function raises.
"""
from __future__ import annotations

import numpy as np

from . import mshds_extractor as _mx


def aggregate_clip_features(clip_features_df, metadata_df, device: int = 0):
    """Drop-in for /root/reference/src/utils.py:7 (same name, arguments, columns, NaN conventions)."""
    import pandas as pd

    if clip_features_df.empty:
        print("Warning: Input clip_features_df is empty. Return an empty aggregated DataFrame.")
        return pd.DataFrame()
    metadata_subset = metadata_df[['filename', 'unique_participant_id']]
    merged = pd.merge(metadata_subset, clip_features_df, on='filename').drop(columns=['filename'])
    feature_cols = [c for c in merged.columns if c != 'unique_participant_id']
    codes, uniques = pd.factorize(merged['unique_participant_id'], sort=True)       # sorted ids, -1 for a missing id
    x = merged[feature_cols].to_numpy(dtype=np.float64, na_value=np.nan) if feature_cols else np.zeros((len(merged), 0))
    if feature_cols:
        mean, std = _mx.get_extractor(device).aggregate_sessions(x, codes.astype(np.int32), len(uniques))
    else:
        mean = std = np.zeros((len(uniques), 0))
    data = {'unique_participant_id': np.asarray(uniques)}
    for j, name in enumerate(feature_cols):
        data[f"{name}_mean"] = mean[:, j]
        data[f"{name}_std"] = std[:, j]
    return pd.DataFrame(data)
