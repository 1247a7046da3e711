"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy) of the reference's two offline preprocessing steps; imported only by
tests/ and by oracle/make_preprocess_golden.py.  Never on the product path.

Follows:
  * expertsim/train/utils.py:81-82 ``get_max_value_image_coordinates`` as looped in
    notebooks/calculate_and_analysis_of_max_coordinates.ipynb cell 6;
  * notebooks/calculating_diversity_for_data.ipynb cells 12-23 (``groupby(CONDITIONAL_COLS).transform(np.std)``,
    ``.sum(axis=1)``, division by the maximum).

Pinned: oracle/make_preprocess_golden.py runs the notebook's own pandas expressions (pandas is the third-party dependency
the arithmetic lives in) on a seeded synthetic set and commits tests/golden/preprocess_small.npz; tests/test_oracle_golden.py
checks this restatement against it.
"""
import numpy as np


def max_coordinates(images: np.ndarray) -> np.ndarray:
    """[N,H,W] -> int64 [N,2]: np.unravel_index(np.argmax(img), img.shape) per image (utils.py:81-82)."""
    out = np.empty((len(images), 2), dtype=np.int64)
    for i, img in enumerate(images):
        out[i] = np.unravel_index(np.argmax(img), img.shape)
    return out


def condition_group_std(cond: np.ndarray, images: np.ndarray) -> np.ndarray:
    """cond [N,K], images [N,H,W] -> float64 [N]: per group of identical conditioning rows, the sum over pixels of the
    population standard deviation over the group's showers, divided by the largest such sum (notebook cells 16-21)."""
    flat = images.reshape(len(images), -1).astype(np.float64)
    groups = {}
    for i, row in enumerate(cond):
        groups.setdefault(row.tobytes(), []).append(i)
    sums = np.empty(len(images), dtype=np.float64)
    for idx in groups.values():
        sums[idx] = np.std(flat[idx], axis=0).sum()       # np.std: ddof = 0
    return sums / sums.max()
