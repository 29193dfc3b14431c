"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by msr3d_b200/).

numpy restatement of the per-object point-cloud preprocessing in front of the hot path
(SURVEY.md §8f-4): `MSR3DDataset.preprocess_pcd` (data/datasets/msr3d.py:181-215) followed by the
padding of `dataset_wrapper.py:151-158`, with the random draws (rotation, subsample indices) taken
as inputs.  Pinned by tests/golden/preprocess.npz, produced by executing the unmodified reference
functions (tests/golden/gen_golden_preprocess.py)."""
from __future__ import annotations

from typing import Optional

import numpy as np


def preprocess_objects(points: np.ndarray, offsets: np.ndarray, sample_idx: np.ndarray,
                       rot: Optional[np.ndarray], o_max: int):
    """points (total, 6) = xyz + rgb of all objects back to back; offsets (O+1,); sample_idx (O, P) indices
    inside each object (np.random.choice of msr3d.py:200-201); rot (3,3) or None (build_rotate_mat).
    Returns obj_fts (o_max, P, 6) f32, obj_locs (o_max, 6) f32, obj_masks (o_max,) bool."""
    O, P = sample_idx.shape
    fts = np.ones((o_max, P, 6), dtype=np.float32)           # dataset_wrapper.py:155 pads with ones
    locs = np.zeros((o_max, 6), dtype=np.float32)            # :157 pads with zeros
    for i in range(O):
        pcd = points[offsets[i]:offsets[i + 1]].copy()
        if rot is not None:
            pcd[:, :3] = np.matmul(pcd[:, :3], rot.transpose())                      # msr3d.py:190-191
        centre = pcd[:, :3].mean(0)                                                  # :193
        size = pcd[:, :3].max(0) - pcd[:, :3].min(0)                                 # :194
        locs[i] = np.concatenate([centre, size], 0)
        sub = pcd[sample_idx[i]]                                                     # :200-202
        sub[:, :3] = sub[:, :3] - sub[:, :3].mean(0)                                 # :205
        max_dist = np.sqrt((sub[:, :3] ** 2).sum(1)).max()                           # :206
        if max_dist < 1e-6:                                                          # :207-208
            max_dist = 1
        sub[:, :3] = sub[:, :3] / max_dist                                           # :209
        fts[i] = sub
    masks = np.arange(o_max) < O                                                     # dataset_wrapper.py:156
    return fts, locs, masks
