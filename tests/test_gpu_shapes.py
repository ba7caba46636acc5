"""GPU: the shapes beyond the fixtures -- many rows (several clusters, partially filled last cluster,
rows finishing at different steps => the row-rank positional-encoding rule across clusters; B = 33 puts
five rows on a 16-CTA cluster, i.e. all three attention groups of a CTA are in use), a
non-square feature map, and the 1024x1024 high-resolution configuration (BASELINE.json configs[4])."""
import numpy as np
import pytest
import torch

from molnextr_b200 import synth
from molnextr_b200.tokenization import CharTokenizer
from tests.helpers import seeded_features, seeded_images

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,path", [(37, "cluster"), (23, "cluster16"), (33, "cluster16"), (37, "graph"),
                                    (37, "wide"), (32, "wide"), (17, "wide")])
def test_many_rows_match_oracle(B, path, monkeypatch):
    from molnextr_b200.engine import Engine
    from oracle import restate
    monkeypatch.setenv("MNX_DECODE_PATH", path)
    dec = synth.decoder_state(0, "sensitised")
    feats = seeded_features(999, B, 144)
    raw = restate.greedy_decode(dec, feats)
    eng = Engine({"decoder": dec, "encoder": None}, max_batch=B)
    out = eng.decode_greedy(feats.cuda())
    torch.cuda.synchronize()
    lens = out["lens"].cpu().numpy()
    ids = out["ids"].cpu().numpy()
    ref_lens = [len(r["ids"]) for r in raw]
    assert len(set(ref_lens)) > 3, "fixture should finish rows at different steps"
    assert lens.tolist() == ref_lens
    for i, r in enumerate(raw):
        assert ids[i, :ref_lens[i]].tolist() == r["ids"].tolist(), f"row {i}"
    eng.close()


def test_decode_with_non_square_memory():
    """S = 13 * 14 = 182 memory positions (the 408x424 image of the padded-window fixture)."""
    from molnextr_b200.engine import Engine
    from oracle import restate
    dec = synth.decoder_state(2, "sensitised")
    feats = seeded_features(31337, 3, 182)
    preds, raw = restate.decode(dec, feats, CharTokenizer(64), return_raw=True)
    eng = Engine({"decoder": dec, "encoder": None}, max_batch=3, max_height=416, max_width=448)
    out = eng.decode_greedy(feats.cuda())
    for i, r in enumerate(raw):
        L = len(r["ids"])
        assert int(out["lens"][i]) == L and out["ids"][i, :L].cpu().tolist() == r["ids"].tolist()
    eng.close()


def test_highres_1024_features_and_decode():
    from molnextr_b200.engine import Engine
    from oracle import restate
    ck = synth.synthetic_checkpoint(0, "sensitised")
    eng = Engine(ck, max_batch=1, max_height=1024, max_width=1024)
    x = seeded_images(5, 1, 1024, 1024)
    with torch.no_grad():
        ref = restate.swin_b_features(ck["encoder"], x)
    feats = eng.encode(x.cuda())
    assert feats.shape == (1, 1024, 1024)
    err = (feats.cpu() - ref).abs()
    print(f"1024^2 feature |err| max {err.max():.4f} mean {err.mean():.5f}")
    assert float(err.max()) <= 0.15 and float(err.mean()) <= 0.015
    # decoder on the reference features: S = 1024 keys per cross-attention
    raw = restate.greedy_decode(ck["decoder"], ref)
    out = eng.decode_greedy(ref.cuda())
    L = len(raw[0]["ids"])
    assert int(out["lens"][0]) == L and out["ids"][0, :L].cpu().tolist() == raw[0]["ids"].tolist()
    eng.close()
