"""CPU: the preprocessing oracle (oracle/preprocess_np.py) is pinned bit-for-bit against OpenCV itself --
the un-vendored dependency whose 8-bit resize / gray code paths it restates -- and against the repo's cv2-based
host transform (molnextr_b200/preprocess.py), on drawings of assorted sizes including the all-white, the
full-frame and the exact-2x (area path) cases."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from molnextr_b200 import preprocess
from oracle import preprocess_np
from tests.helpers_images import drawings


def test_resize_and_gray_match_cv2_bit_for_bit():
    rng = np.random.default_rng(1)
    for h, w in [(470, 923), (239, 263), (384, 384), (768, 768), (385, 383), (50, 2000), (1536, 1536), (3, 5)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = cv2.resize(img, (384, 384), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(preprocess_np.resize_linear_u8(img, 384, 384), ref), (h, w)
        assert np.array_equal(preprocess_np.rgb2gray_u8(ref), cv2.cvtColor(ref, cv2.COLOR_RGB2GRAY))


def test_transform_matches_the_cv2_host_path():
    for i, img in enumerate(drawings()):
        a = preprocess_np.transform(img)
        b = preprocess.transform(img)
        assert a.shape == b.shape == (3, 384, 384) and a.dtype == b.dtype == np.float32
        assert np.array_equal(a, b), f"drawing {i} {img.shape}"


def test_crop_box_follows_cropwhite():
    img = np.full((40, 60, 3), 255, np.uint8)
    assert preprocess_np.crop_box(img) == (0, 40, 0, 60)
    img[7, 11] = (255, 254, 255)
    img[30, 50] = (0, 0, 0)
    assert preprocess_np.crop_box(img) == (7, 31, 11, 51)


def test_resize_restatement_fuzz_against_cv2():
    """Random source sizes (upscaling, downscaling, extreme aspect ratios, 1-pixel sides) and destination sizes."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 700), st.integers(1, 700), st.sampled_from([(384, 384), (96, 96), (37, 53), (1, 1)]), st.integers(0, 2**31 - 1))
    def check(h, w, dst, seed):
        img = np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = cv2.resize(img, dst, interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(preprocess_np.resize_linear_u8(img, dst[0], dst[1]), ref), (h, w, dst)

    check()
