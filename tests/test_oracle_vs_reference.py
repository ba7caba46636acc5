"""Build-container only: oracle/restate.py against the reference's OWN modules imported from
/root/reference (skipped where the reference is absent, e.g. on the GPU box)."""
import pytest
import torch

from molnextr_b200 import synth
from molnextr_b200.tokenization import CharTokenizer
from oracle import ref_loader, restate

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")


def test_reference_state_dict_schema_matches_synthetic():
    ck = synth.synthetic_checkpoint(3, "sensitised")
    enc, dec, tok = ref_loader.build_reference(ck)      # strict=True load
    assert set(enc.state_dict()) == set(ck["encoder"]) and set(dec.state_dict()) == set(ck["decoder"])
    pe = dec.state_dict()["decoder.chartok_coords.embeddings.make_embedding.pe.pe"]
    assert torch.equal(pe, synth.positional_encoding_table())
    idx = enc.state_dict()["transformer.layers.0.blocks.0.attn.relative_position_index"]
    assert torch.equal(idx, synth.relative_position_index(12))


def test_oracle_equals_reference_swin_decode():
    ck = synth.synthetic_checkpoint(5, "sensitised")
    enc, dec, tok = ref_loader.build_reference(ck)
    g = torch.Generator().manual_seed(123)
    x = torch.randn((2, 3, 384, 384), generator=g)
    with torch.no_grad():
        f_ref, _ = enc(x)
        f_or = restate.swin_b_features(ck["encoder"], x)
        assert torch.allclose(f_ref, f_or, rtol=0, atol=1e-5)
        p_ref = dec.decode(f_ref, None)
    p_or = restate.decode(ck["decoder"], f_or, CharTokenizer(64))
    assert p_ref == p_or


def test_oracle_equals_reference_partial_label_decode():
    """Fresh seed, labels built by the fixture script's recipes, the reference's own TransformerDecoderAR.decode(labels=...)."""
    from oracle.make_golden import partial_labels, seeded_features
    ck = synth.synthetic_checkpoint(6, "sensitised")
    _, dec, _ = ref_loader.build_reference(ck)
    ar = dec.decoder["chartok_coords"]
    feats = seeded_features(77, 4, 36)
    with torch.no_grad():
        free, *_ = ar.decode(feats, 1, 1, max_length=480)
        labels = torch.from_numpy(partial_labels([o[0].numpy() for o in free]))
        outputs, scores, token_scores, hiddens = ar.decode(feats, 1, 1, max_length=480, labels=labels)
    raw = restate.greedy_decode(ck["decoder"], feats, labels=labels)
    for i, r in enumerate(raw):
        assert r["ids"].tolist() == outputs[i][0].tolist()
        assert torch.allclose(torch.exp(r["logp"]).double(), torch.tensor(token_scores[i][0]).double(), rtol=1e-5, atol=1e-9)
        assert torch.allclose(r["hidden"], hiddens[i][0], rtol=0, atol=1e-5)


def test_oracle_equals_reference_partial_label_edge_cases():
    """Labels that give every token, end a row at step 0, or are too short: same results / same IndexError as the reference."""
    from oracle.make_golden import seeded_features
    ck = synth.synthetic_checkpoint(6, "sensitised")
    _, dec, _ = ref_loader.build_reference(ck)
    ar = dec.decoder["chartok_coords"]
    feats = seeded_features(78, 3, 36)
    with torch.no_grad():
        free, *_ = ar.decode(feats, 1, 1, max_length=480)
    L = [len(o[0]) for o in free]
    labels = torch.full((3, 481), 0, dtype=torch.long)
    labels[:, 0] = 1
    n0 = min(L[0], 40)
    labels[0, 1:1 + n0] = free[0][0][:n0]          # row 0: every token given (teacher forcing), then <eos>
    labels[0, n0] = 2
    labels[1, 1] = 2                               # row 1: the first label is <eos>: finished after step 0
    labels[2, 1:] = 4                              # row 2: everything <mask>, no <eos> label: runs to max_len
    with torch.no_grad():
        outputs, scores, token_scores, hiddens = ar.decode(feats, 1, 1, max_length=480, labels=labels)
    raw = restate.greedy_decode(ck["decoder"], feats, labels=labels)
    for i, r in enumerate(raw):
        assert r["ids"].tolist() == outputs[i][0].tolist()
        assert torch.allclose(torch.exp(r["logp"]).double(), torch.tensor(token_scores[i][0]).double(), rtol=1e-5, atol=1e-9)
    assert len(outputs[1][0]) == 1 and len(outputs[2][0]) == 480
    short = torch.full((3, 6), 4, dtype=torch.long)
    with pytest.raises(IndexError):
        with torch.no_grad():
            ar.decode(feats, 1, 1, max_length=480, labels=short)
    with pytest.raises(IndexError):
        restate.greedy_decode(ck["decoder"], feats, labels=short)
