"""CPU oracle of the image preprocessing: a numpy restatement of what
`get_transforms(384, 384, augment=False)` does to one RGB uint8 image (MolNexTR/dataset.py:158-185,
MolNexTR/data_aug.py:98-143 CropWhite; albumentations 1.1.0 Resize / ToGray / Normalize call cv2.resize,
cv2.cvtColor and three numpy float32 operations).

TEST INFRASTRUCTURE ONLY.  OpenCV is a dependency of the reference that is not vendored in it; its 8-bit
code paths are restated here from the published algorithm (modules/imgproc/src/resize.cpp: 11-bit fixed-point
bilinear coefficients, HResizeLinear / VResizeLinear; color_yuv: 15-bit RGB2GRAY) and PINNED bit-for-bit
against cv2 itself (4.13, present in this image and on the GPU box) by tests/test_preprocess_oracle.py."""
from __future__ import annotations

import numpy as np

MEAN = np.array([0.485, 0.456, 0.406], np.float32)
STD = np.array([0.229, 0.224, 0.225], np.float32)


def crop_box(img: np.ndarray):
    """CropWhite.update_params (data_aug.py:106-136): tight box of the pixels != (255,255,255)."""
    nonwhite = (img != 255).any(axis=2)
    if not nonwhite.any():
        return 0, img.shape[0], 0, img.shape[1]
    rows, cols = np.flatnonzero(nonwhite.any(axis=1)), np.flatnonzero(nonwhite.any(axis=0))
    return int(rows[0]), int(rows[-1]) + 1, int(cols[0]), int(cols[-1]) + 1


def resize_linear_u8(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR) for HxWxC uint8."""
    sh, sw = src.shape[:2]
    if sw == 2 * dw and sh == 2 * dh:      # exact 2x decimation: resize() switches INTER_LINEAR to the fast area path
        s = src.astype(np.int32)
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)

    def coords(dn, sn):
        scale = 1.0 / (dn / sn)
        f = ((np.arange(dn, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int32)
        return s, (f - s.astype(np.float32)).astype(np.float32)

    sx, fx = coords(dw, sw)
    lo, hi = sx < 0, sx >= sw - 1           # the x fraction is zeroed at the borders; rows are clipped instead
    fx[lo], sx[lo] = 0, 0
    fx[hi], sx[hi] = 0, sw - 1
    sy, fy = coords(dh, sh)
    fix = lambda c: np.rint(c * np.float32(2048)).astype(np.int32)      # saturate_cast<short>(c * 2048): cvRound
    a0, a1, b0, b1 = fix(np.float32(1) - fx), fix(fx), fix(np.float32(1) - fy), fix(fy)
    S = src.astype(np.int32)
    Hp = S[:, sx] * a0[None, :, None] + S[:, np.minimum(sx + 1, sw - 1)] * a1[None, :, None]
    r0, r1 = Hp[np.clip(sy, 0, sh - 1)], Hp[np.clip(sy + 1, 0, sh - 1)]
    out = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def rgb2gray_u8(img: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(img, cv2.COLOR_RGB2GRAY) for uint8: 15-bit fixed point."""
    r, g, b = (img[..., i].astype(np.int32) for i in range(3))
    return ((r * 9798 + g * 19235 + b * 3735 + (1 << 14)) >> 15).astype(np.uint8)


def transform(image: np.ndarray, size: int = 384, pad: int = 50) -> np.ndarray:
    """RGB uint8 HxWx3 -> float32 (3, size, size), normalised."""
    t, b, l, r = crop_box(image)
    crop = image[t:b, l:r]
    padded = np.full((crop.shape[0] + 2 * pad, crop.shape[1] + 2 * pad, 3), 255, np.uint8)
    padded[pad:pad + crop.shape[0], pad:pad + crop.shape[1]] = crop
    gray = rgb2gray_u8(resize_linear_u8(padded, size, size))
    img = np.repeat(gray[:, :, None], 3, axis=2).astype(np.float32)
    img = (img - MEAN * 255.0) * (1.0 / (STD * 255.0))
    return np.ascontiguousarray(img.transpose(2, 0, 1))
