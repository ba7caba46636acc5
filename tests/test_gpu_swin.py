"""GPU parity of the Swin-B encoder (C ABI: mnx_encode) and of the whole image -> ids/bonds path
(mnx_predict / mnx_predict_host).

Encoder arithmetic is bf16 operands with fp32 accumulation on tcgen05 (the reference's own
multi-GPU eval runs the encoder under fp16 autocast, main.py:277), so features are compared to
the fp32 reference within a stated tolerance: max |err| <= 0.12 and mean |err| <= 0.012 on
LayerNorm-ed features of unit scale.  The decoder is fp32; ids are compared exactly and, where
the bf16 encoder flips a near-tie, the test requires the flip to be explained by the tolerance:
at the first diverging step the reference's own top-2 log-prob gap must be below LOGP_TOL."""
import numpy as np
import pytest
import torch

from molnextr_b200 import synth
from molnextr_b200.tokenization import CharTokenizer
from tests.helpers import load_golden, seeded_images

pytestmark = pytest.mark.gpu
FEAT_MAX_TOL, FEAT_MEAN_TOL, LOGP_TOL = 0.12, 0.012, 0.15
# Bond classes given the SAME features are bit-exact (tests/test_gpu_decoder.py, and checked again below on the engine's own
# features).  Against the fp32-encoder fixture a class may flip only where the reference's own symmetrised top-2 probability
# margin is small: features that differ by <= 0.035 (bf16 GEMM operands) move the bond head's probabilities by several
# hundredths; observed margins at flipped entries over rounds 1-2: <= 0.14.
EDGE_GAP_TOL = 0.2


@pytest.fixture(scope="module")
def engine_cache():
    from molnextr_b200.engine import Engine
    cache = {}

    def get(seed, max_batch=4, hw=(384, 384)):
        key = (seed, max_batch, hw)
        if key not in cache:
            ck = synth.synthetic_checkpoint(seed, "sensitised")
            cache[key] = Engine(ck, max_batch=max_batch, max_height=hw[0], max_width=hw[1])
        return cache[key]

    yield get
    for e in cache.values():
        e.close()


def _feature_check(feats, ref_sub, ref_abs):
    f = feats.float().cpu().numpy()
    err = np.abs(f[:, ::4, ::32] - ref_sub)
    print(f"feature |err| max {err.max():.4f} mean {err.mean():.5f}")
    assert err.max() <= FEAT_MAX_TOL and err.mean() <= FEAT_MEAN_TOL
    np.testing.assert_allclose(np.abs(f).sum((1, 2)), ref_abs, rtol=5e-3)


@pytest.mark.parametrize("name,hw", [("swin_b4_384.npz", (384, 384)), ("swin_b1_408x424.npz", (408, 424))])
def test_swin_features_match_reference_fixture(engine_cache, name, hw):
    g = load_golden(name)
    cfg = g["cfg"]
    eng = engine_cache(cfg["ckpt_seed"], max_batch=cfg["b"], hw=hw)
    x = seeded_images(cfg["img_seed"], cfg["b"], cfg["h"], cfg["w"]).cuda()
    feats = eng.encode(x)
    torch.cuda.synchronize()
    _feature_check(feats, g["feat_sub"], g["feat_abs"])


def test_swin_features_match_oracle_fresh_inputs(engine_cache):
    from oracle import restate
    eng = engine_cache(0, max_batch=4)
    x = seeded_images(77, 2, 384, 384)
    with torch.no_grad():
        ref = restate.swin_b_features(synth.swin_b_state(0), x)
    feats = eng.encode(x.cuda()).cpu()
    err = (feats - ref).abs()
    print(f"feature |err| max {err.max():.4f} mean {err.mean():.5f}")
    assert err.max() <= FEAT_MAX_TOL and err.mean() <= FEAT_MEAN_TOL


def _explainable(ids, ref_ids, ref_len, lp_by_step, row_rank_fn):
    """ids must equal the reference's; a divergence is accepted only at a reference near-tie."""
    L = int(ref_len)
    mism = np.nonzero(ids[:L] != ref_ids[:L])[0]
    return (len(mism) == 0), (int(mism[0]) if len(mism) else -1)


def test_predict_end_to_end_vs_reference_fixture(engine_cache):
    g = load_golden("swin_b4_384.npz")
    cfg = g["cfg"]
    eng = engine_cache(cfg["ckpt_seed"], max_batch=cfg["b"])
    x = seeded_images(cfg["img_seed"], cfg["b"], cfg["h"], cfg["w"])
    out = eng.predict(x.cuda())
    host = eng.predict_host(x)
    torch.cuda.synchronize()
    for k in out:
        assert torch.equal(out[k].cpu(), host[k]), f"predict and predict_host disagree on {k}"
    ids, lens = out["ids"].cpu().numpy(), out["lens"].cpu().numpy()
    exact_rows = 0
    from oracle import restate
    # the decoder + bond head are fp32: the oracle fed the ENGINE's features must agree (ids and atoms exactly, bond classes as in smoke(): > 99.9 %)
    feats = eng.encode(x.cuda()).cpu()
    preds_own, raw_own = restate.decode(synth.synthetic_checkpoint(cfg["ckpt_seed"], cfg["variant"])["decoder"], feats,
                                        CharTokenizer(64), return_raw=True)
    for i, r in enumerate(raw_own):
        L = len(r["ids"])
        assert int(lens[i]) == L and ids[i, :L].tolist() == r["ids"].tolist(), f"row {i}: ids differ from the oracle on the same features"
        k = len(preds_own[i]["edges"])
        assert int(out["n_atoms"][i]) == k
        if k:
            same_e = (out["edges"][i, :k, :k].cpu().numpy().astype(np.int8) == np.asarray(preds_own[i]["edges"], np.int8)).mean()
            assert same_e > 0.999, f"row {i}: only {same_e:.4f} of bond classes match the oracle on the same features"
    # teacher-forced reference log-probs to judge divergences (oracle on the host CPU)
    ck = synth.synthetic_checkpoint(cfg["ckpt_seed"], cfg["variant"])
    with torch.no_grad():
        ref_feats = restate.swin_b_features(ck["encoder"], x)
    raw = restate.greedy_decode(ck["decoder"], ref_feats, record_logprobs=True)
    for i in range(cfg["b"]):
        L = int(g["lens"][i])
        same = lens[i] == L and (ids[i, :L] == g["ids"][i, :L]).all()
        if same:
            exact_rows += 1
            k = int(g["natoms"][i])
            assert int(out["n_atoms"][i]) == k
            assert out["atom_idx"][i, :k].cpu().tolist() == g["atom_idx"][i, :k].tolist()
            # the bond head reads decoder hidden states computed from bf16-GEMM encoder features: a bond class may flip
            # only where the reference's own symmetrised probabilities are a near-tie (top-2 gap <= EDGE_GAP_TOL)
            got_e = out["edges"][i, :k, :k].cpu().numpy().astype(np.int8)
            mism = np.argwhere(got_e != g["edges"][i, :k, :k])
            frac = 1.0 - len(mism) / max(1, k * k)
            if len(mism):
                hid = raw[i]["hidden"]
                prob = restate.edge_probabilities(ck["decoder"], hid, g["atom_idx"][i, :k].tolist()).numpy()
                top2 = np.sort(restate.symmetrised_edge_probabilities(prob), axis=2)[:, :, -2:]
                gaps = np.array([top2[a, b, 1] - top2[a, b, 0] for a, b in mism])
                print(f"row {i}: ids exact, {len(mism)} of {k * k} bond classes differ, reference top-2 gaps max {gaps.max():.4f}")
                assert gaps.max() <= EDGE_GAP_TOL, f"row {i}: a bond class differs where the reference margin is {gaps.max():.3f}"
            # (the random-weight bond head is close to uniform over its 7 classes, so many entries sit inside the margin:
            #  observed 0.95 - 1.0 over rounds 1-2; the margin test above is the parity statement, this one a sanity floor)
            assert frac >= 0.90, f"row {i}: only {frac:.3f} of bond classes match"
        else:
            m = np.nonzero(ids[i, :L] != g["ids"][i, :L])[0]
            t = int(m[0]) if len(m) else min(L, int(lens[i]))
            top2 = raw[i]["logprobs"][t].topk(2).values
            gap = float(top2[0] - top2[1])
            print(f"row {i}: diverges at step {t}, reference top-2 gap {gap:.4f}")
            assert gap <= LOGP_TOL, f"row {i} diverged at step {t} where the reference margin is {gap:.3f}"
    print(f"{exact_rows}/{cfg['b']} rows bit-exact end to end")


def test_pipelined_batches_equal_batch_by_batch():
    """Engine.predict_pipelined (encoder of batch i+1 overlapped with the persistent decode kernel of
    batch i on two streams) must return exactly what predict returns batch by batch, device and host mode."""
    from molnextr_b200.engine import Engine
    ck = synth.synthetic_checkpoint(0, "sensitised")
    eng = Engine(ck, max_batch=4)
    xs = [seeded_images(100 + i, 4 if i != 2 else 3, 384, 384) for i in range(4)]
    ref = [{k: v.cpu() for k, v in eng.predict(x.cuda()).items()} for x in xs]
    got = eng.predict_pipelined([x.cuda() for x in xs])
    torch.cuda.synchronize()
    got_host = eng.predict_pipelined([x.pin_memory() for x in xs], host=True)
    for r, g, h in zip(ref, got, got_host):
        for k in ("ids", "lens", "n_atoms", "atom_idx", "edges"):
            if k in ("atom_idx", "edges"):      # entries beyond n_atoms are unspecified scratch
                for i, n in enumerate(r["n_atoms"].tolist()):
                    a, b, c = (t[k][i][:n] if k == "atom_idx" else t[k][i][:n, :n] for t in (r, g, h))
                    assert torch.equal(a, b.cpu()) and torch.equal(a, c), k
            else:
                assert torch.equal(r[k], g[k].cpu()) and torch.equal(r[k], h[k]), k
    eng.close()


def test_pipelined_contexts_and_tail_spreading():
    """Several decode contexts in flight (throughput kernel, two clusters per batch) and a partial last wave whose
    batches are spread over more clusters of fewer rows: still exactly the batch-by-batch results."""
    from molnextr_b200.engine import Engine
    ck = synth.synthetic_checkpoint(0, "sensitised")
    eng = Engine(ck, max_batch=20)
    xs = [seeded_images(200 + i, 20 if i % 2 == 0 else 17, 384, 384) for i in range(5)]
    ref = [{k: v.cpu() for k, v in eng.predict(x.cuda()).items()} for x in xs]
    for depth in (2, 4):        # depth 4: the fifth batch is alone in its wave -> 4 rows per cluster
        got = eng.predict_pipelined([x.cuda() for x in xs], depth=depth)
        torch.cuda.synchronize()
        for r, g in zip(ref, got):
            assert torch.equal(r["lens"], g["lens"].cpu()) and torch.equal(r["ids"], g["ids"].cpu())
            assert torch.equal(r["n_atoms"], g["n_atoms"].cpu())
            for i, n in enumerate(r["n_atoms"].tolist()):
                assert torch.equal(r["edges"][i][:n, :n], g["edges"][i][:n, :n].cpu())
    eng.close()
