"""GPU: mnx_preprocess (CropWhite -> Resize -> ToGray -> Normalize on the device) is bit-exact with the
reference's transform as executed by OpenCV on the host (molnextr_b200/preprocess.py = cv2.resize +
cv2.cvtColor + the numpy float32 normalisation) and with the numpy oracle, for drawings of assorted sizes."""
import numpy as np
import pytest
import torch

from molnextr_b200 import preprocess, synth
from oracle import preprocess_np
from tests.helpers_images import drawings

pytestmark = pytest.mark.gpu


def test_device_preprocess_bit_exact():
    from molnextr_b200.engine import Engine
    eng = Engine({"decoder": synth.decoder_state(0, "sensitised"), "encoder": None}, max_batch=16)
    imgs = drawings()
    out = eng.preprocess(imgs).cpu().numpy()
    assert out.shape == (len(imgs), 3, 384, 384)
    for i, im in enumerate(imgs):
        want = preprocess_np.transform(im)
        assert np.array_equal(out[i], want), f"drawing {i} {im.shape}: {(out[i] != want).sum()} of {want.size} values differ"
        try:
            import cv2  # noqa: F401
            assert np.array_equal(out[i], preprocess.transform(im))
        except ImportError:
            pass
    # a second call with another batch composition reuses the crop-box scratch correctly
    out2 = eng.preprocess(imgs[3:7]).cpu().numpy()
    assert np.array_equal(out2, out[3:7])
    eng.close()
