"""CPU: the beam-search oracle (oracle/restate.py beam_decode).  The reference's beam branch cannot
execute, so this oracle is "parity unpinned"; what CAN be pinned is pinned here:
  * with beam_size = 1 it must reproduce the greedy oracle (itself bit-exact against the
    reference's own GreedySearch fixtures) token for token -- same decoder step, same grammar mask,
    same row-rank positional rule, same end conditions;
  * following its own recorded selections must reproduce itself (the follow mode used by the GPU
    parity test is sound);
  * a committed fixture guards the algorithm against drift (tests/golden/beam_b4_s64_k5.npz,
    written by oracle/make_golden_beam.py)."""
import numpy as np
import torch

from molnextr_b200 import synth
from oracle import restate
from tests.helpers import load_golden, seeded_features

MAXLEN = 96   # keeps the CPU suite short; the state machine does not depend on the cap


def test_beam1_equals_greedy_oracle():
    dec = synth.decoder_state(0, "sensitised")
    feats = seeded_features(16, 4, 64)          # rows finish at different steps (12 .. cap)
    g = restate.greedy_decode(dec, feats, max_len=MAXLEN)
    b = restate.beam_decode(dec, feats, beam_size=1, n_best=1, max_len=MAXLEN)
    assert len({len(r["ids"]) for r in g}) > 1
    for rg, rb in zip(g, b):
        assert rg["ids"].tolist() == rb[0]["ids"].tolist()
        # re-indexing the caches every step changes tensor strides, hence BLAS summation order
        torch.testing.assert_close(rg["logp"], rb[0]["logp"], rtol=0, atol=2e-5)
        torch.testing.assert_close(rg["hidden"], rb[0]["hidden"], rtol=0, atol=2e-5)
        # BeamSearch files cum / (L + 1) (curr_length = len(self) + 1 counts <sos>, beam_search.py:99);
        # GreedySearch files exp(mean(logp))
        L = len(rg["ids"])
        np.testing.assert_allclose(np.exp(rb[0]["score"] * (L + 1) / L), rg["score"], rtol=1e-4)


def test_follow_own_selections_reproduces_and_flags_bad_ones():
    dec = synth.decoder_state(0, "sensitised")
    feats = seeded_features(16, 3, 64)
    K = 4
    rec = np.full((MAXLEN, 3, K), -1, np.int64)
    a = restate.beam_decode(dec, feats, beam_size=K, n_best=2, max_len=MAXLEN, record=rec)
    st = {}
    b = restate.beam_decode(dec, feats, beam_size=K, n_best=2, max_len=MAXLEN, follow=rec, follow_stats=st)
    assert st == {}
    for ra, rb in zip(a, b):
        assert [h["ids"].tolist() for h in ra] == [h["ids"].tolist() for h in rb]
        assert [h["score"] for h in ra] == [h["score"] for h in rb]
    bad = rec.copy()
    bad[3, 1, 0] = bad[3, 1, 1]                   # the same candidate twice
    st = {}
    restate.beam_decode(dec, feats, beam_size=K, n_best=2, max_len=8, follow=bad, follow_stats=st)
    assert st["max_violation"] == float("inf")
    worse = rec.copy()
    worse[2, 0, K - 1] = 3 * 229 + 5              # an arbitrary poor candidate in the last place
    st = {}
    restate.beam_decode(dec, feats, beam_size=K, n_best=2, max_len=8, follow=worse, follow_stats=st)
    assert st["max_violation"] > 1e-2


def test_beam_oracle_matches_committed_fixture():
    g = load_golden("beam_b4_s64_k5.npz")
    cfg = g["cfg"]
    dec = synth.decoder_state(cfg["ckpt_seed"], "sensitised")
    feats = seeded_features(cfg["feat_seed"], cfg["b"], cfg["s"])
    res = restate.beam_decode(dec, feats, beam_size=cfg["beam"], n_best=cfg["n_best"], max_len=cfg["max_len"])
    for i, r in enumerate(res):
        assert len(r) == cfg["n_best"]
        for n, h in enumerate(r):
            L = int(g["lens"][i, n])
            assert len(h["ids"]) == L
            assert h["ids"].tolist() == g["ids"][i, n, :L].tolist()
            np.testing.assert_allclose(h["score"], g["scores"][i, n], rtol=1e-6)
    # properties of the algorithm: hypotheses come out best first, end with <eos> unless capped,
    # and obey the coordinate grammar (X-bin is followed by a Y-bin, Y-bin by a symbol)
    for r in res:
        assert all(r[n]["score"] >= r[n + 1]["score"] for n in range(len(r) - 1))
        for h in r:
            ids = h["ids"].tolist()
            assert ids[-1] == restate.EOS_ID or len(ids) == cfg["max_len"]
            for a, b in zip(ids[:-1], ids[1:]):
                if 101 <= a < 165:
                    assert 165 <= b < 229
                elif a >= 165:
                    assert b < 101
