"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (its own Encoder/Decoder/
GreedySearch/tokenizer modules from /root/reference, third-party classes restated in
oracle/ref_shims) on seeded synthetic checkpoints and seeded inputs.

Run in the build container only (needs /root/reference):
    python -m oracle.make_golden
The fixtures are small, committed, and are what pins oracle/restate.py and the CUDA path
on the GPU box, where the reference does not exist.  Inputs/weights are never stored: both
sides regenerate them from seeds (molnextr_b200/synth.py, torch CPU generator).
"""
from __future__ import annotations

import json
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from molnextr_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
LOGPROB_STEPS = (0, 1, 2, 3, 7, 20, 60, 140)


def seeded_images(seed: int, b: int, h: int, w: int) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn((b, 3, h, w), generator=g, dtype=torch.float32)


def seeded_features(seed: int, b: int, s: int, c: int = 1024) -> torch.Tensor:
    """Synthetic encoder maps: per-token noise plus a per-image common component (rows then
    finish at different steps, which exercises compaction and the row-rank PE rule)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    tok = torch.randn((b, s, c), generator=g, dtype=torch.float32)
    img = torch.randn((b, 1, c), generator=g, dtype=torch.float32)
    return 0.4 * tok + 0.9 * img


def _pack_predictions(preds, raw_ids, raw_logp, raw_hidden, max_len):
    B = len(preds)
    ids = np.zeros((B, max_len), np.int32)
    logp = np.zeros((B, max_len), np.float64)
    lens = np.zeros((B,), np.int32)
    kmax = max([len(p["edges"]) for p in preds] + [1])
    edges = np.full((B, kmax, kmax), -1, np.int8)
    natoms = np.zeros((B,), np.int32)
    atom_idx = np.full((B, kmax), -1, np.int32)
    hidden_sub = np.zeros((B, max_len, 16), np.float32)
    hidden_sum = np.zeros((B, max_len), np.float32)
    for i, p in enumerate(preds):
        L = len(raw_ids[i])
        lens[i] = L
        ids[i, :L] = raw_ids[i]
        logp[i, :L] = raw_logp[i]
        h = raw_hidden[i]
        hidden_sub[i, :L] = h[:, ::16]
        hidden_sum[i, :L] = h.sum(1)
        k = len(p["edges"])
        natoms[i] = k
        if k:
            edges[i, :k, :k] = np.asarray(p["edges"], np.int8)
            atom_idx[i, :k] = p["chartok_coords"]["indices"]
    meta = [dict(smiles=p["chartok_coords"]["smiles"], symbols=p["chartok_coords"]["symbols"],
                 coords=p["chartok_coords"]["coords"], indices=p["chartok_coords"]["indices"]) for p in preds]
    return dict(ids=ids, token_scores=logp, lens=lens, edges=edges, natoms=natoms, atom_idx=atom_idx,
                hidden_sub=hidden_sub, hidden_sum=hidden_sum, meta=np.array(json.dumps(meta)))


def _run_reference_decode(dec, features):
    """decoder.decode through the reference, recording what GreedySearch.advance was fed."""
    import MolNexTR.components as comp
    rec = {"lp": []}
    orig_advance = comp.GreedySearch.advance

    def advance(self, log_probs, attn=None, hidden=None, label=None):
        orig_advance(self, log_probs, attn, hidden, label)   # ensure_min_length edits in place first
        rec["lp"].append(log_probs.detach().clone())

    comp.GreedySearch.advance = advance
    try:
        with torch.no_grad():
            ar = dec.decoder["chartok_coords"]
            from MolNexTR.utils import FORMAT_INFO
            max_len = FORMAT_INFO["chartok_coords"]["max_len"]
            outputs, scores, token_scores, hiddens = ar.decode(features, 1, 1, max_length=max_len)
            preds = dec.decode(features, None)
    finally:
        comp.GreedySearch.advance = orig_advance
    # the second call re-ran the same deterministic decode; keep the first call's log-probs
    nsteps = len(rec["lp"]) // 2
    lps = rec["lp"][:nsteps]
    raw_ids = [o[0].numpy() for o in outputs]
    raw_logp = [np.asarray(t[0], np.float64) for t in token_scores]   # exp(log-prob), as the reference reports it
    raw_hidden = [h[0].numpy() for h in hiddens]
    return preds, raw_ids, raw_logp, raw_hidden, lps, scores


def make_swin_e2e(name, seed, b, h, w):
    ck = synth.synthetic_checkpoint(seed, "sensitised")
    enc, dec, tok = ref_loader.build_reference(ck)
    x = seeded_images(1000 + seed, b, h, w)
    with torch.no_grad():
        feats, hiddens = enc(x)
    preds, raw_ids, raw_logp, raw_hidden, lps, scores = _run_reference_decode(dec, feats)
    out = _pack_predictions(preds, raw_ids, raw_logp, raw_hidden, 480)
    f = feats.numpy()
    out.update(feat_sub=f[:, ::4, ::32].copy(), feat_sum=f.sum((1, 2)), feat_abs=np.abs(f).sum((1, 2)),
               stage_mean=np.array([float(hd.mean()) for hd in hiddens], np.float32),
               stage_std=np.array([float(hd.std()) for hd in hiddens], np.float32),
               seq_score=np.array([s[0] for s in scores], np.float32),
               cfg=np.array(json.dumps(dict(ckpt_seed=seed, variant="sensitised", img_seed=1000 + seed, b=b, h=h, w=w))))
    for st in LOGPROB_STEPS:
        if st < len(lps):
            out[f"logprobs_step{st}"] = lps[st].numpy()
    np.savez_compressed(os.path.join(GOLDEN, name), **out)
    print(name, "lens", out["lens"].tolist(), "atoms", out["natoms"].tolist())


def make_decoder_only(name, seed, b, s):
    ck = synth.synthetic_checkpoint(seed, "sensitised")
    _, dec, tok = ref_loader.build_reference(ck)
    feats = seeded_features(2000 + seed, b, s)
    preds, raw_ids, raw_logp, raw_hidden, lps, scores = _run_reference_decode(dec, feats)
    out = _pack_predictions(preds, raw_ids, raw_logp, raw_hidden, 480)
    out.update(seq_score=np.array([sc[0] for sc in scores], np.float32),
               cfg=np.array(json.dumps(dict(ckpt_seed=seed, variant="sensitised", feat_seed=2000 + seed, b=b, s=s))))
    for st in LOGPROB_STEPS:
        if st < len(lps):
            out[f"logprobs_step{st}"] = lps[st].numpy()
    np.savez_compressed(os.path.join(GOLDEN, name), **out)
    print(name, "lens", out["lens"].tolist(), "atoms", out["natoms"].tolist())


def make_confidence(name, seed, b, s):
    """Decoder.decode with compute_confidence=True (components.py:456-469,485-491): per-atom geometric-mean token
    score, symmetrised edge scores, overall score = average token score * sqrt(prod(edge scores))."""
    ck = synth.synthetic_checkpoint(seed, "sensitised")
    _, dec, tok = ref_loader.build_reference(ck, compute_confidence=True)
    feats = seeded_features(3000 + seed, b, s)
    with torch.no_grad():
        preds = dec.decode(feats, None)
    kmax = max([len(p["edges"]) for p in preds] + [1])
    natoms = np.array([len(p["edges"]) for p in preds], np.int32)
    atom_scores = np.zeros((b, kmax), np.float64)
    edge_scores = np.zeros((b, kmax, kmax), np.float64)
    overall = np.zeros((b,), np.float64)
    for i, p in enumerate(preds):
        k = natoms[i]
        assert "average_token_score" not in p["chartok_coords"]
        atom_scores[i, :k] = p["chartok_coords"]["atom_scores"]
        if k:
            edge_scores[i, :k, :k] = np.asarray(p["edge_scores"], np.float64)
        overall[i] = p["overall_score"]
    smiles = [p["chartok_coords"]["smiles"] for p in preds]
    np.savez_compressed(os.path.join(GOLDEN, name), natoms=natoms, atom_scores=atom_scores, edge_scores=edge_scores,
                        overall_score=overall, meta=np.array(json.dumps(dict(smiles=smiles))),
                        cfg=np.array(json.dumps(dict(ckpt_seed=seed, variant="sensitised", feat_seed=3000 + seed, b=b, s=s))))
    print(name, "atoms", natoms.tolist(), "overall", overall.tolist())


def partial_labels(free_ids, max_len=480, offset=101):
    """Label rows for a partial-label decode from the rows' FREE greedy ids; recipes 2, 0, 1, 3 are dealt in that order
    over the rows sorted by free length (shortest first, so the rows that stop by themselves get recipes 2 and 0):
      0  the reference's own use (tokenization.py:448-449): symbols given, both coordinate tokens <mask>, <eos> given
      1  everything <mask>, <eos> label half way through the free sequence (the label ends the row early)
      2  seven given tokens borrowed from the next row, then <mask>, <eos> label five positions AFTER the free
         sequence's end (the model's own <eos> must not end the row)
      3  everything <mask>, no <eos> label at all (runs to max_len)
    Width max_len + 1 so that labels[:, step] and labels[:, step + 1] exist for every step (components.py:287,305)."""
    PAD, SOS, EOS, MASK = 0, 1, 2, 4
    B = len(free_ids)
    lab = np.full((B, max_len + 1), PAD, np.int64)
    lab[:, 0] = SOS
    order = sorted(range(B), key=lambda r: (len(free_ids[r]), r))
    kinds = {r: (2, 0, 1, 3)[n % 4] for n, r in enumerate(order)}
    for i, ids in enumerate(free_ids):
        ids = [int(v) for v in ids]
        L = len(ids)
        kind = kinds[i]
        if kind == 0:
            body = [MASK if v >= offset else v for v in ids]
            if body[-1] != EOS and L < max_len:
                body.append(EOS)
            lab[i, 1:1 + len(body)] = body
        elif kind == 1:
            p = max(2, L // 2)
            lab[i, 1:p] = MASK
            lab[i, p] = EOS
        elif kind == 2:
            other = [int(v) for v in free_ids[(i + 1) % B]][:7]
            end = min(max_len, L + 5)
            lab[i, 1:end] = MASK
            lab[i, 1:1 + len(other)] = [v if v != EOS else MASK for v in other]
            lab[i, end] = EOS
        else:
            lab[i, 1:] = MASK
    return lab


def make_partial(name, seed, b, s):
    """TransformerDecoderAR.decode(labels=...) run by the reference itself (components.py:286-289,305,317-318,326-332)."""
    ck = synth.synthetic_checkpoint(seed, "sensitised")
    _, dec, tok = ref_loader.build_reference(ck)
    feats = seeded_features(4000 + seed, b, s)
    ar = dec.decoder["chartok_coords"]
    with torch.no_grad():
        free, *_ = ar.decode(feats, 1, 1, max_length=480)
        labels = torch.from_numpy(partial_labels([o[0].numpy() for o in free]))
        outputs, scores, token_scores, hiddens = ar.decode(feats, 1, 1, max_length=480, labels=labels)
    ids = np.zeros((b, 480), np.int32)
    lens = np.zeros((b,), np.int32)
    dec_len = np.zeros((b,), np.int32)         # steps the row was alive (token_scores / hidden keep this length)
    logp = np.zeros((b, 480), np.float64)
    hidden_sub = np.zeros((b, 480, 16), np.float32)
    hidden_sum = np.zeros((b, 480), np.float32)
    for i in range(b):
        o = outputs[i][0].numpy()
        lens[i] = len(o)
        ids[i, :len(o)] = o
        ts = np.asarray(token_scores[i][0], np.float64)
        dec_len[i] = len(ts)
        logp[i, :len(ts)] = ts
        h = hiddens[i][0].numpy()
        hidden_sub[i, :len(ts)] = h[:, ::16]
        hidden_sum[i, :len(ts)] = h.sum(1)
    np.savez_compressed(os.path.join(GOLDEN, name), labels=labels.numpy().astype(np.int32), ids=ids, lens=lens, dec_len=dec_len,
                        token_scores=logp, hidden_sub=hidden_sub, hidden_sum=hidden_sum,
                        free_lens=np.array([len(o[0]) for o in free], np.int32),
                        seq_score=np.array([sc[0] for sc in scores], np.float32),
                        cfg=np.array(json.dumps(dict(ckpt_seed=seed, variant="sensitised", feat_seed=4000 + seed, b=b, s=s))))
    print(name, "free lens", [len(o[0]) for o in free], "lens", lens.tolist(), "alive steps", dec_len.tolist())


def make_tokenizer_and_edges(name):
    ck = synth.synthetic_checkpoint(0, "sensitised")
    _, dec, tok = ref_loader.build_reference(ck)
    import MolNexTR.components as comp
    rt = tok["chartok_coords"]
    rnd = random.Random(7)
    seqs, outs = [], []
    for _ in range(400):
        L = rnd.randint(0, 70)
        seq = [rnd.choice([rnd.randint(0, 228), rnd.randint(5, 100), rnd.randint(101, 164), rnd.randint(165, 228),
                           57, 14, 15, 40, 56, 46, 68, 69]) for _ in range(L)]
        seqs.append(seq)
        outs.append(rt.sequence_to_smiles(seq))
    masks = np.array([rt.get_output_mask(i) for i in range(len(rt))], bool)
    g = np.random.default_rng(11)
    edge_cases = []
    for n in (0, 1, 2, 5, 17):
        p = g.random((n, n, 7)).astype(np.float32)
        p = p / p.sum(2, keepdims=True) if n else p
        pred, score = comp.get_edge_prediction(p.astype(np.float64).tolist())
        edge_cases.append(dict(prob=p.tolist(), pred=pred, score=score))
    np.savez_compressed(os.path.join(GOLDEN, name), masks=masks, vocab_len=len(rt), offset=rt.offset,
                        tok=np.array(json.dumps(dict(seqs=seqs, outs=outs))),
                        edges=np.array(json.dumps(edge_cases)))
    print(name, "ok")


def main():
    assert ref_loader.available(), "needs /root/reference (build container)"
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    make_tokenizer_and_edges("tokenizer_edges.npz")
    make_decoder_only("decoder_b6_s144.npz", seed=0, b=6, s=144)
    make_decoder_only("decoder_b3_s64.npz", seed=1, b=3, s=64)
    make_swin_e2e("swin_b4_384.npz", seed=0, b=4, h=384, w=384)
    make_swin_e2e("swin_b1_408x424.npz", seed=2, b=1, h=408, w=424)
    make_confidence("confidence_b5_s144.npz", seed=0, b=5, s=144)
    make_partial("partial_b8_s64.npz", seed=0, b=8, s=64)


if __name__ == "__main__":
    main()
