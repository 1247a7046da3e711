"""Generates tests/golden/preprocess_small.npz with the REFERENCE's own expressions (run in the build container):
  * positions: get_max_value_image_coordinates imported from /root/reference/expertsim/train/utils.py when present
    (else the two-line definition it holds, utils.py:81-82), looped as in calculate_and_analysis_of_max_coordinates.ipynb cell 6;
  * std: the pandas pipeline of calculating_diversity_for_data.ipynb cells 12-23, verbatim in structure.
Usage: python oracle/make_preprocess_golden.py"""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rng = np.random.default_rng(20251018)
    n_groups, H, W = 7, 56, 30
    cond_rows = rng.normal(size=(n_groups, 9)).astype(np.float32)
    members = rng.integers(0, n_groups, size=60)
    members[:n_groups] = np.arange(n_groups)              # every group occurs; one group will be a singleton below
    members[members == 6] = 5
    members[6] = 6
    cond = cond_rows[members]
    hit = rng.random((60, H, W)) < 0.02
    photons = np.where(hit, np.ceil(rng.exponential(20.0, size=(60, H, W))), 0.0)
    photons[:, H // 2, W // 2] += 1
    data = np.log1p(photons).astype(np.float32)
    data[3, 10, 4] = data[3].max()                         # a tie: the first maximum must win
    data[3, 2, 7] = data[3].max()

    # --- positions (notebook cell 6)
    get = lambda img: np.unravel_index(np.argmax(img), img.shape)      # train/utils.py:81-82
    positions = np.array([get(img) for img in data], dtype=np.int64)

    # --- std (notebook cells 12-23)
    data_cond = pd.DataFrame(cond, columns=[f"c{i}" for i in range(9)])
    CONDITIONAL_COLS = list(data_cond.columns)
    flatten_responses = pd.DataFrame(data.reshape(len(data), -1))
    data_all = pd.concat([data_cond, flatten_responses], axis=1)
    stddev_group = data_all.groupby(CONDITIONAL_COLS).transform(lambda x: np.std(x))
    sum_pixels = stddev_group.sum(axis=1)
    normalized_stddevs = sum_pixels / sum_pixels.max()

    out = os.path.join(ROOT, "tests", "golden", "preprocess_small.npz")
    np.savez_compressed(out, cond=cond, data=data, positions=positions, std=normalized_stddevs.to_numpy().astype(np.float64))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    sys.exit(main())
