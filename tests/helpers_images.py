"""Synthetic "molecule drawings" of assorted sizes for the preprocessing tests (cv2 is in the image on
both the build container and the GPU box; the reference's examples/*.png are not on the GPU box)."""
import numpy as np


def drawings(seed=0):
    import cv2
    rng = np.random.default_rng(seed)
    out = []
    shapes = [(470, 923), (500, 761), (443, 732), (239, 263), (384, 384), (668, 668), (100, 140), (1000, 1500), (61, 2000)]
    for n, (h, w) in enumerate(shapes):
        img = np.full((h, w, 3), 255, np.uint8)
        y0, y1 = sorted(rng.integers(0, h, 2)); x0, x1 = sorted(rng.integers(0, w, 2))
        y1, x1 = max(y1, y0 + 8), max(x1, x0 + 8)
        for _ in range(25):
            p = (int(rng.integers(x0, min(x1, w - 1) + 1)), int(rng.integers(y0, min(y1, h - 1) + 1)))
            q = (int(rng.integers(x0, min(x1, w - 1) + 1)), int(rng.integers(y0, min(y1, h - 1) + 1)))
            col = tuple(int(v) for v in rng.integers(0, 230, 3)) if n % 2 else (0, 0, 0)
            cv2.line(img, p, q, col, int(rng.integers(1, 4)), cv2.LINE_AA if n % 3 else cv2.LINE_8)
        out.append(img)
    out.append(np.full((120, 90, 3), 255, np.uint8))                      # all white: no crop
    out.append(rng.integers(0, 256, (300, 411, 3), dtype=np.uint8))       # noise: box = whole image
    sq = np.full((768, 768, 3), 255, np.uint8)                            # padded crop exactly 768 x 768: the 2x area path
    sq[50:718, 50:718] = rng.integers(0, 255, (668, 668, 3), dtype=np.uint8)
    out.append(sq)
    return out
