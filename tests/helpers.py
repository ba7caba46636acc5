"""Shared helpers for parity tests: regenerate the seeded inputs the golden fixtures used."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    for k in ("cfg", "meta", "tok", "edges"):
        if k in d and d[k].dtype.kind == "U":
            d[k] = json.loads(str(d[k]))
    return d


def seeded_images(seed, b, h, w):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn((b, 3, h, w), generator=g, dtype=torch.float32)


def seeded_features(seed, b, s, c=1024):
    g = torch.Generator(device="cpu").manual_seed(seed)
    tok = torch.randn((b, s, c), generator=g, dtype=torch.float32)
    img = torch.randn((b, 1, c), generator=g, dtype=torch.float32)
    return 0.4 * tok + 0.9 * img
