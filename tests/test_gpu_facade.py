"""GPU: the reference-facing Python surface (components.Decoder.decode, model.molnextr.predict_*)."""
import numpy as np
import pytest
import torch

from molnextr_b200 import synth
from tests.helpers import load_golden, seeded_features

pytestmark = pytest.mark.gpu


def test_decoder_decode_reproduces_reference_prediction_dicts():
    from molnextr_b200.components import Decoder
    from molnextr_b200.engine import Engine
    g = load_golden("decoder_b6_s144.npz")
    cfg = g["cfg"]
    eng = Engine({"decoder": synth.decoder_state(cfg["ckpt_seed"], cfg["variant"]), "encoder": None}, max_batch=cfg["b"])
    feats = seeded_features(cfg["feat_seed"], cfg["b"], cfg["s"]).cuda()
    preds = Decoder(eng).decode(feats, None)
    assert len(preds) == cfg["b"]
    for i, (p, m) in enumerate(zip(preds, g["meta"])):
        ct = p["chartok_coords"]
        assert ct["smiles"] == m["smiles"] and ct["symbols"] == m["symbols"]
        assert ct["indices"] == m["indices"] and ct["coords"] == m["coords"]
        k = int(g["natoms"][i])
        assert np.array_equal(np.asarray(p["edges"], np.int8).reshape(k, k), g["edges"][i, :k, :k])
    conf = Decoder(eng, compute_confidence=True).decode(feats, None)
    assert all("overall_score" in p and "atom_scores" in p["chartok_coords"] and "edge_scores" in p for p in conf)
    eng.close()


def test_confidence_values_match_reference_fixture():
    """compute_confidence=True (components.py:456-469,485-491): per-atom geometric-mean token scores, symmetrised edge
    scores and the overall score against values the reference itself produced (oracle/make_golden.make_confidence)."""
    from molnextr_b200.components import Decoder
    from molnextr_b200.engine import Engine
    g = load_golden("confidence_b5_s144.npz")
    cfg = g["cfg"]
    eng = Engine({"decoder": synth.decoder_state(cfg["ckpt_seed"], cfg["variant"]), "encoder": None}, max_batch=cfg["b"])
    feats = seeded_features(cfg["feat_seed"], cfg["b"], cfg["s"]).cuda()
    preds = Decoder(eng, compute_confidence=True).decode(feats, None)
    assert [p["chartok_coords"]["smiles"] for p in preds] == g["meta"]["smiles"]
    for i, p in enumerate(preds):
        k = int(g["natoms"][i])
        assert len(p["edges"]) == k and "average_token_score" not in p["chartok_coords"]
        np.testing.assert_allclose(p["chartok_coords"]["atom_scores"], g["atom_scores"][i, :k], rtol=2e-4, atol=1e-7)
        np.testing.assert_allclose(np.asarray(p["edge_scores"]), g["edge_scores"][i, :k, :k], rtol=0, atol=2e-5)
        want = float(g["overall_score"][i])
        # a product of k*k edge scores (up to 25 000 factors, often underflowing): compare in the log domain
        if want == 0.0:
            assert p["overall_score"] < 1e-250
        else:
            assert abs(np.log(p["overall_score"]) - np.log(want)) <= 1e-6 * k * k + 1e-3
    eng.close()


def test_molnextr_predict_images_schema():
    from molnextr_b200.model import molnextr
    ck = synth.synthetic_checkpoint(0, "sensitised")
    m = molnextr(None, device="cuda", max_batch=16, checkpoint=ck)
    rng = np.random.default_rng(0)
    images = []
    for k in range(3):
        img = np.full((180 + 40 * k, 260, 3), 255, np.uint8)
        for _ in range(12):
            y, x = rng.integers(20, img.shape[0] - 20), rng.integers(20, 240)
            img[y:y + 3, x:x + 25] = 0
        images.append(img)
    out = m.predict_images(images, return_atoms_bonds=True, return_confidence=True)
    assert len(out) == 3
    for d in out:
        assert set(d) >= {"predicted_smiles", "predicted_molfile", "atom_sets", "bond_sets"}
        for a in d["atom_sets"]:
            assert set(a) == {"atom_number", "atom_symbol", "coords", "confidence"}
        for b in d["bond_sets"]:
            assert b["bond_type"] in ("single", "double", "triple", "aromatic", "solid wedge", "dashed wedge")
    one = m.predict_image(images[0])
    assert one["predicted_smiles"] == out[0]["predicted_smiles"]   # row 0 of a batch == the same image alone
    m.engine.close()
