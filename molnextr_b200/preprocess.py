"""Image preprocessing of the reference's inference path, restated with cv2 only.

`molnextr.transform = get_transforms(384, 384, augment=False)` (MolNexTR/model.py:48,
MolNexTR/dataset.py:158-185) = CropWhite(pad=50) -> Resize(384, 384, INTER_LINEAR) -> ToGray ->
Normalize(ImageNet) -> ToTensorV2.  albumentations is not vendored by the reference and is not
installed here; the four steps are simple enough to state directly.
This is CPU code upstream of the accelerated path (SURVEY.md section 8f-1 lists a GPU version as
the next row)."""
from __future__ import annotations

import numpy as np

MEAN = np.array([0.485, 0.456, 0.406], np.float32)
STD = np.array([0.229, 0.224, 0.225], np.float32)


def crop_white(img: np.ndarray, value=(255, 255, 255), pad: int = 50) -> np.ndarray:
    """Tight crop of the non-white pixels, then a constant white border (data_aug.py:98-143)."""
    assert img.ndim == 3 and img.shape[2] == 3
    h, w = img.shape[:2]
    nonwhite = (img != np.asarray(value, img.dtype).reshape(1, 1, 3)).sum(axis=2) > 0
    rows, cols = nonwhite.sum(axis=1), nonwhite.sum(axis=0)
    if rows.sum() == 0:
        top = left = 0
        bottom, right = h, w
    else:
        top = int(np.argmax(rows > 0))
        bottom = h - int(np.argmax(rows[::-1] > 0))
        left = int(np.argmax(cols > 0))
        right = w - int(np.argmax(cols[::-1] > 0))
    img = img[top:bottom, left:right]
    if pad > 0:
        import cv2
        img = cv2.copyMakeBorder(img, pad, pad, pad, pad, cv2.BORDER_CONSTANT, value=value)
    return img


def transform(image: np.ndarray, size: int = 384) -> np.ndarray:
    """RGB uint8 HxWx3 -> float32 (3, size, size), normalised."""
    import cv2
    img = crop_white(np.ascontiguousarray(image))
    img = cv2.resize(img, (size, size), interpolation=cv2.INTER_LINEAR)
    gray = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY)
    img = cv2.cvtColor(gray, cv2.COLOR_GRAY2RGB).astype(np.float32)
    img = (img - MEAN * 255.0) * (1.0 / (STD * 255.0))
    return np.ascontiguousarray(img.transpose(2, 0, 1))
