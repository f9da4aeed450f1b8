"""Oracle preprocess (fixed-point restatement of cv2) vs real cv2 (FlappyBirdDQN.py:31-34)."""
import os

import numpy as np

from oracle import flappy_oracle as fo


def test_resize_tables_match_survey():
    t = fo.resize_tables()
    sx, a0, a1, sy, b0, b1 = t
    # weight pairs cycle with period 5
    assert [(int(a0[i]), int(a1[i])) for i in range(5)] == [(1434, 614), (205, 1843), (1024, 1024), (1843, 205), (614, 1434)]
    assert [(int(b0[i]), int(b1[i])) for i in range(5)] == [(614, 1434), (1843, 205), (1024, 1024), (205, 1843), (1434, 614)]
    for i in range(80):
        assert (a0[i], a1[i]) == (a0[i % 5], a1[i % 5]) and (b0[i], b1[i]) == (b0[i % 5], b1[i % 5])
        assert sx[i] == 18 * (i // 5) + [1, 4, 8, 12, 15][i % 5]
        assert sy[i] == 32 * (i // 5) + [2, 9, 15, 21, 28][i % 5]
    assert sx.max() + 1 == 286 and sy.max() + 1 == 509


def test_preprocess_matches_golden_cv2(golden_dir):
    g = np.load(os.path.join(golden_dir, "cv2_preprocess.npz"))
    for f, o in zip(g["frames"], g["obs"]):
        np.testing.assert_array_equal(fo.preprocess(f), o)


def test_preprocess_matches_live_cv2():
    import cv2
    assert cv2.__version__.startswith("4."), cv2.__version__
    rng = np.random.default_rng(11)
    for k in range(6):
        hi = [256, 256, 4, 8, 256, 3][k]
        f = rng.integers(0, hi, (288, 512, 3), dtype=np.uint8)
        if k == 4:
            f *= (rng.random((288, 512, 1)) < 0.03).astype(np.uint8)
        ref = cv2.cvtColor(cv2.resize(f, (80, 80)), cv2.COLOR_BGR2GRAY)
        _, ref = cv2.threshold(ref, 1, 255, cv2.THRESH_BINARY)
        np.testing.assert_array_equal(fo.preprocess(f), ref)
