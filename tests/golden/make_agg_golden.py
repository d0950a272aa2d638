"""Generates tests/golden/session_agg_golden_v1.npz by running the REFERENCE's own aggregate_clip_features
(/root/reference/src/utils.py:7-58) in the build container.  The reference file cannot travel to the GPU box; the vectors do.

    python tests/golden/make_agg_golden.py
"""
import importlib.util
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/utils.py"


def make_inputs(seed=20260101, n_sessions=23, d=25):
    rng = np.random.default_rng(seed)
    names, sess = [], []
    for s in range(n_sessions):
        k = 1 if s % 7 == 0 else int(rng.integers(2, 14))          # single-clip sessions give NaN std
        for c in range(k):
            names.append(f"{s:02d}_clip{c:03d}.wav")
            sess.append(f"P{(n_sessions - s):02d}_{'PD' if s % 2 else 'HC'}")      # ids NOT in row order
    n = len(names)
    x = rng.normal(size=(n, d)) * np.logspace(-3, 4, d)[None, :] + np.linspace(-50, 5000, d)[None, :]
    x[rng.random((n, d)) < 0.06] = np.nan                          # helper failures
    x[:, 3] = np.where(np.arange(n) % 5 == 0, x[:, 3], np.nan)     # a column that is mostly missing
    order = rng.permutation(n)                                     # clip frame in another order than the metadata
    cols = [f"feat_{j:02d}" for j in range(d)]
    clip_df = pd.DataFrame(x[order], columns=cols)
    clip_df.insert(0, "filename", [names[i] for i in order])
    meta = pd.DataFrame({"filename": names + ["not_extracted.wav"], "unique_participant_id": sess + ["P99_HC"],
                         "label": [0] * (n + 1)})
    return clip_df, meta


if __name__ == "__main__":
    spec = importlib.util.spec_from_file_location("ref_utils", REF)
    ref = importlib.util.module_from_spec(spec)
    sys.modules["tqdm.auto"] = sys.modules.get("tqdm.auto") or __import__("tqdm.auto")
    spec.loader.exec_module(ref)
    clip_df, meta = make_inputs()
    out = ref.aggregate_clip_features(clip_df, meta)
    np.savez_compressed(os.path.join(HERE, "session_agg_golden_v1.npz"),
                        clip_filenames=np.array(clip_df["filename"]), clip_values=clip_df.iloc[:, 1:].to_numpy(dtype=np.float64),
                        clip_columns=np.array(clip_df.columns[1:]),
                        meta_filenames=np.array(meta["filename"]), meta_ids=np.array(meta["unique_participant_id"]),
                        out_ids=np.array(out["unique_participant_id"]), out_columns=np.array(out.columns[1:]),
                        out_values=out.iloc[:, 1:].to_numpy(dtype=np.float64), pandas_version=pd.__version__)
    print(out.shape, "pandas", pd.__version__)
