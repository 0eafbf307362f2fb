"""The batch-assembly oracle (oracle/batch_ref.py) against the golden vectors produced by the REAL reference
(oracle/make_golden_batch.py), and the host mirror's RNG-free logic.  CPU only."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import batch_ref as BR

NAMES = ["rot6d", "rotmat", "rot_pos", "joint_pos", "linear_v", "angular_v", "root_v"]


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLDEN, "batch.npz")))


@pytest.fixture(scope="module")
def mean_std(smpl):
    ms = smpl["mean_std"].copy()
    ms[1, ms[1] == 0] = 1.0
    return ms


def test_rand_rotation_matrix_known_answers(gold):
    for r, full, small in zip(gold["rr_randnums"], gold["rr_full"], gold["rr_small"]):
        np.testing.assert_allclose(BR.rand_rotation_matrix(1.0, r), full, rtol=0, atol=1e-15)
        np.testing.assert_allclose(BR.rand_rotation_matrix(0.25, r), small, rtol=0, atol=1e-15)
        m = BR.rand_rotation_matrix(1.0, r)
        np.testing.assert_allclose(m @ m.T, np.eye(3), atol=1e-12)
        assert np.linalg.det(m) > 0.999


@pytest.mark.parametrize("tag", ["plain", "rot", "rot64", "fps_rot"])
def test_assemble_matches_reference_getitem(gold, mean_std, tag):
    idx, T, freq, t0, rot = [int(v) for v in gold[f"{tag}_meta"]]
    window = gold[f"seq{idx}"][0::freq][t0:t0 + T]
    M = BR.rand_rotation_matrix(1.0, gold[f"{tag}_randnums"]) if rot else None
    out = BR.assemble(window, mean_std, M)
    for n, v in zip(NAMES, out):
        ref = gold[f"{tag}_{n}"]
        assert v.shape == ref.shape and v.dtype == np.float32, n
        if n in ("rot_pos", "joint_pos", "linear_v", "angular_v") or not rot:
            np.testing.assert_array_equal(v, ref, err_msg=n)            # copies and float64 standardisation: bit-exact
        else:
            np.testing.assert_allclose(v, ref, rtol=2e-6, atol=2e-6, err_msg=n)   # fp32 3x3 products (summation order)


def test_change_fps_factor_rule():
    seq = iter([12, 10, 3])
    assert BR.change_fps_factor(96, 8, lambda: next(seq)) == 12          # 96 / 12 = 8 frames: enough
    seq = iter([12, 10, 8, 6, 5, 4, 3, 2, 1, 1])
    assert BR.change_fps_factor(40, 8, lambda: next(seq)) == 5           # 12, 10, 8, 6 leave < 8 frames; 5 leaves 8
    assert BR.change_fps_factor(5, 8, lambda: 2) == 1                    # never enough: the original data


@pytest.mark.parametrize("tag,seed,fps", [("plain", 11, False), ("rot", 12, False), ("rot64", 13, False), ("fps_rot", 14, True)])
def test_host_mirror_draws_the_same_window_as_the_reference(gold, tag, seed, fps):
    """hm_vae_b200.utils_motion_vae keeps the reference's host-side index logic and RNG consumption order: with the generators
    seeded like oracle/make_golden_batch.py seeded them, it crops the window and draws the three numbers the reference used."""
    import random

    from hm_vae_b200 import utils_motion_vae as U

    idx, T, freq, t0, rot = [int(v) for v in gold[f"{tag}_meta"]]
    random.seed(seed)
    np.random.seed(seed)
    window = U.crop_window(gold[f"seq{idx}"], T, fps_aug_flag=fps)
    np.testing.assert_array_equal(window, gold[f"seq{idx}"][0::freq][t0:t0 + T])
    if rot:
        rnd = np.random.uniform(size=(3,))
        np.testing.assert_array_equal(rnd, gold[f"{tag}_randnums"])
        np.testing.assert_allclose(U.rand_rotation_matrix(1.0, rnd), BR.rand_rotation_matrix(1.0, rnd), atol=1e-15)
    assert U.crop_window(gold["seq0"][:5], 8) is None          # too short: the reference draws another sequence
