"""CPU: the oracle restatement and the host-side tokenizer against the fixtures the REFERENCE
wrote (oracle/make_golden.py).  This is what pins oracle/restate.py."""
import numpy as np
import pytest
import torch

from molnextr_b200 import synth
from molnextr_b200.tokenization import CharTokenizer
from oracle import restate
from tests.helpers import load_golden, seeded_features, seeded_images


def _check_decode(g, preds, raw, ids_exact=True):
    for i, (p, r) in enumerate(zip(preds, raw)):
        L = int(g["lens"][i])
        assert len(r["ids"]) == L
        assert r["ids"].tolist() == g["ids"][i, :L].tolist()
        np.testing.assert_allclose(np.exp(r["logp"].double().numpy()), g["token_scores"][i, :L], rtol=2e-4, atol=1e-7)
        np.testing.assert_allclose(r["hidden"][:, ::16].numpy(), g["hidden_sub"][i, :L], rtol=0, atol=2e-4)
        m = g["meta"][i]
        assert p["chartok_coords"]["smiles"] == m["smiles"]
        assert p["chartok_coords"]["symbols"] == m["symbols"]
        assert p["chartok_coords"]["indices"] == m["indices"]
        assert p["chartok_coords"]["coords"] == m["coords"]
        k = int(g["natoms"][i])
        assert len(p["edges"]) == k
        if k:
            assert np.array_equal(np.asarray(p["edges"], np.int8), g["edges"][i, :k, :k])


@pytest.mark.parametrize("name", ["decoder_b3_s64.npz", "decoder_b6_s144.npz"])
def test_oracle_decoder_matches_reference_fixture(name):
    g = load_golden(name)
    cfg = g["cfg"]
    dec = synth.decoder_state(cfg["ckpt_seed"], cfg["variant"])
    feats = seeded_features(cfg["feat_seed"], cfg["b"], cfg["s"])
    tok = CharTokenizer(64)
    rawlp = restate.greedy_decode(dec, feats, record_logprobs=True)
    preds, raw = restate.decode(dec, feats, tok, return_raw=True)
    _check_decode(g, preds, raw)
    # per-step masked log-probs as GreedySearch.advance saw them (alive rows only, in order)
    lens = g["lens"]
    for st in (0, 1, 2, 3, 7, 20, 60, 140):
        key = f"logprobs_step{st}"
        if key not in g:
            continue
        rows = [i for i in range(cfg["b"]) if lens[i] > st]
        mine = torch.stack([rawlp[i]["logprobs"][st] for i in rows]).numpy()
        ref = g[key]
        assert mine.shape == ref.shape
        np.testing.assert_allclose(mine, ref, rtol=1e-5, atol=2e-4)


def test_oracle_swin_end_to_end_matches_reference_fixture():
    g = load_golden("swin_b1_408x424.npz")   # exercises window padding and odd patch-merging
    cfg = g["cfg"]
    ck = synth.synthetic_checkpoint(cfg["ckpt_seed"], cfg["variant"])
    x = seeded_images(cfg["img_seed"], cfg["b"], cfg["h"], cfg["w"])
    with torch.no_grad():
        feats = restate.swin_b_features(ck["encoder"], x)
    f = feats.numpy()
    np.testing.assert_allclose(f[:, ::4, ::32], g["feat_sub"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(np.abs(f).sum((1, 2)), g["feat_abs"], rtol=1e-5)
    preds, raw = restate.decode(ck["decoder"], feats, CharTokenizer(64), return_raw=True)
    _check_decode(g, preds, raw)


def test_tokenizer_and_edge_rule_match_reference_fixture():
    g = load_golden("tokenizer_edges.npz")
    tok = CharTokenizer(64)
    assert len(tok) == int(g["vocab_len"]) == 229 and tok.offset == int(g["offset"]) == 101
    masks = np.array([tok.get_output_mask(i) for i in range(len(tok))], bool)
    assert np.array_equal(masks, g["masks"])
    dev = restate.grammar_mask(torch.arange(229), *tok.grammar_rule()).numpy()
    assert np.array_equal(dev, g["masks"])
    for seq, out in zip(g["tok"]["seqs"], g["tok"]["outs"]):
        assert tok.sequence_to_smiles(seq) == out
    for case in g["edges"]:
        prob = np.asarray(case["prob"], np.float32).reshape(len(case["prob"]), len(case["prob"]), 7)
        pred, score = restate.get_edge_prediction(prob)
        assert pred == case["pred"]
        np.testing.assert_allclose(np.asarray(score, float).reshape(-1), np.asarray(case["score"], float).reshape(-1), rtol=1e-12)


def test_synthetic_checkpoint_schema():
    enc = synth.swin_b_state(0)
    dec = synth.decoder_state(0)
    assert len(enc) == 351 and sum(v.numel() for k, v in enc.items() if "relative_position_index" not in k) == 86878584
    n_dec = sum(v.numel() for k, v in dec.items() if not k.endswith("pe.pe"))
    assert n_dec == 6834156
    assert dec["decoder.chartok_coords.embeddings.make_embedding.emb_luts.0.weight"][0].abs().sum() == 0


def test_oracle_partial_label_decode_matches_reference_fixture():
    """TransformerDecoderAR.decode(labels=...) (components.py:286-289,305,317-318,326-332): fixture written by the reference."""
    g = load_golden("partial_b8_s64.npz")
    cfg = g["cfg"]
    dec = synth.decoder_state(cfg["ckpt_seed"], cfg["variant"])
    feats = seeded_features(cfg["feat_seed"], cfg["b"], cfg["s"])
    labels = torch.from_numpy(g["labels"]).long()
    raw = restate.greedy_decode(dec, feats, labels=labels)
    for i, r in enumerate(raw):
        L, D = int(g["lens"][i]), int(g["dec_len"][i])
        assert r["ids"].tolist() == g["ids"][i, :L].tolist()
        assert len(r["logp"]) == D
        np.testing.assert_allclose(np.exp(r["logp"].double().numpy()), g["token_scores"][i, :D], rtol=2e-4, atol=1e-7)
        np.testing.assert_allclose(r["hidden"][:, ::16].numpy(), g["hidden_sub"][i, :D], rtol=0, atol=2e-4)
    # the recipes did what they were meant to: a row cut short by its label, one that outlives its own <eos>
    assert (g["lens"] < g["free_lens"]).any() and (g["lens"] > g["free_lens"]).any()
    # labels narrower than the decode: the reference's IndexError (components.py:287)
    with pytest.raises(IndexError):
        restate.greedy_decode(dec, feats[:2], labels=torch.full((2, 5), 4, dtype=torch.long))
