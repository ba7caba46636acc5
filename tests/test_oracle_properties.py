"""CPU: structural properties of the reference algorithm that the GPU full-size tests lean on, checked on the
(pinned) oracle itself:

  * prefix invariance -- a row's positional-encoding rank is its rank among the ALIVE rows before it
    (SURVEY.md F3), so the first k rows of a batch decode exactly as a batch made of those k rows alone,
    for greedy and for beam search (rows = image-major, beam-minor);
  * conversely a row's output DOES depend on the rows before it: moving it changes its rank and its ids --
    which is why sharded decoding is defined per shard and the facade keeps the
    reference's mini-batch chunking;
  * batch compaction preserves order: dropping finished rows never reorders the survivors."""
import pytest
import torch

from molnextr_b200 import synth
from oracle import restate
from tests.helpers import seeded_features

MAXLEN = 64


@pytest.fixture(scope="module")
def dec():
    return synth.decoder_state(0, "sensitised")


@pytest.mark.parametrize("seed,k", [(16, 2), (16, 4)])
def test_greedy_prefix_invariance(dec, seed, k):
    feats = seeded_features(seed, 5, 64)
    full = restate.greedy_decode(dec, feats, max_len=MAXLEN)
    part = restate.greedy_decode(dec, feats[:k].contiguous(), max_len=MAXLEN)
    assert len({len(r["ids"]) for r in full}) > 1           # rows do finish at different steps
    for a, b in zip(full[:k], part):
        assert a["ids"].tolist() == b["ids"].tolist()
        torch.testing.assert_close(a["logp"], b["logp"], rtol=0, atol=2e-5)


def test_beam_prefix_invariance(dec):
    feats = seeded_features(16, 4, 64)
    full = restate.beam_decode(dec, feats, beam_size=3, n_best=2, max_len=MAXLEN)
    part = restate.beam_decode(dec, feats[:2].contiguous(), beam_size=3, n_best=2, max_len=MAXLEN)
    for a, b in zip(full[:2], part):
        assert [h["ids"].tolist() for h in a] == [h["ids"].tolist() for h in b]


def test_row_position_in_the_batch_matters(dec):
    feats = seeded_features(16, 4, 64)
    a = restate.greedy_decode(dec, feats, max_len=MAXLEN)
    swapped = feats[[1, 0, 2, 3]].contiguous()
    b = restate.greedy_decode(dec, swapped, max_len=MAXLEN)
    # image 0 decoded at rank 1 instead of rank 0 receives another positional encoding at every step
    h0, h1 = a[0]["hidden"], b[1]["hidden"]
    n = min(len(h0), len(h1))
    assert (h0[:n] - h1[:n]).abs().max() > 1e-3
    # rows 2 and 3 keep their ranks as long as rows 0 / 1 are alive in both runs: identical first steps
    alive = min(len(a[0]["ids"]), len(a[1]["ids"]))
    for i in (2, 3):
        m = min(alive, len(a[i]["ids"]), len(b[i]["ids"]))
        assert a[i]["ids"][:m].tolist() == b[i]["ids"][:m].tolist()
