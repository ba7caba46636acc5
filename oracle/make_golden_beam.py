"""Write tests/golden/beam_b4_s64_k5.npz from the beam-search ORACLE (oracle/restate.py
beam_decode).  Unlike the other fixtures this one is NOT produced by the reference: its beam
branch cannot execute, so the fixture only guards the repaired algorithm against
drift ("parity unpinned").  Run:  python -m oracle.make_golden_beam"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from molnextr_b200 import synth  # noqa: E402
from oracle import restate  # noqa: E402
from oracle.make_golden import GOLDEN, seeded_features  # noqa: E402


def main():
    cfg = dict(ckpt_seed=0, feat_seed=16, b=4, s=64, beam=5, n_best=3, max_len=480)
    dec = synth.decoder_state(cfg["ckpt_seed"], "sensitised")
    feats = seeded_features(cfg["feat_seed"], cfg["b"], cfg["s"])
    gaps = [float("inf")] * cfg["b"]
    res = restate.beam_decode(dec, feats, beam_size=cfg["beam"], n_best=cfg["n_best"], max_len=cfg["max_len"], min_gap=gaps)
    B, NB, T = cfg["b"], cfg["n_best"], cfg["max_len"]
    ids = np.zeros((B, NB, T), np.int32)
    lens = np.zeros((B, NB), np.int32)
    scores = np.zeros((B, NB), np.float64)
    for i, r in enumerate(res):
        for n, h in enumerate(r):
            L = len(h["ids"])
            ids[i, n, :L] = h["ids"].numpy()
            lens[i, n] = L
            scores[i, n] = h["score"]
    np.savez_compressed(os.path.join(GOLDEN, "beam_b4_s64_k5.npz"), cfg=json.dumps(cfg), ids=ids, lens=lens,
                        scores=scores, min_gap=np.asarray(gaps))
    print("lens", lens.tolist(), "min gaps", gaps)


if __name__ == "__main__":
    main()
