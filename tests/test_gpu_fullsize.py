"""GPU: BASELINE.json's large configurations, checked through size-independent properties (the
CPU oracle would need minutes at these sizes):

  * configs[2]  bs = 256, beam_size = 5 (1280 decoder rows): KV-cache stress;
  * configs[3]  bs = 2048 sharded over 8 GPUs = 256 greedy rows per GPU (the per-GPU shard).

Properties:
  1. prefix invariance -- the positional-encoding rank of a row depends only on the rows BEFORE it
     in the batch, and every kernel treats rows independently, so the first images
     of the big batch must decode to exactly what they decode to in a small batch, which
     test_gpu_beam.py / test_gpu_decoder.py pin against the oracle;
  2. every sequence obeys the coordinate grammar and ends with <eos> or at the 480 cap;
  3. beam hypotheses come out best-first with finite scores.
Throughput is printed (-s) for the record; bench.py's line stays on configs[1]."""
import time

import numpy as np
import pytest
import torch

from molnextr_b200 import synth
from tests.helpers import seeded_features

pytestmark = pytest.mark.gpu
EOS = 2


def _check_grammar(ids, L):
    seq = ids[:L].tolist()
    assert L > 0 and (seq[-1] == EOS or L == 480)
    for a, b in zip(seq[:-1], seq[1:]):
        if 101 <= a < 165:
            assert 165 <= b < 229
        elif a >= 165:
            assert b < 101
    assert (ids[L:] == 0).all()


def _timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0


def test_beam5_bs256_prefix_invariance_and_grammar():
    from molnextr_b200.engine import Engine
    dec = synth.decoder_state(0, "sensitised")
    B, K = 256, 5
    feats = seeded_features(77, B, 144).cuda()
    eng = Engine({"decoder": dec, "encoder": None}, max_batch=B, max_beam=K)
    eng.decode_beam(feats[:8], K, 2)                       # warm-up (graph capture for another shape)
    big, dt = _timed(lambda: eng.decode_beam(feats, K, 2))
    print(f"beam=5 bs=256 (1280 rows): {dt * 1e3:.1f} ms, {eng.last_decode_steps()} steps, {B / dt:.0f} img/s")
    small = eng.decode_beam(feats[:6], K, 2)
    torch.cuda.synchronize()
    for k in ("ids", "lens"):
        assert torch.equal(big[k][:6], small[k]), k
    # 1280 rows run the register-tiled GEMM, 30 rows the skinny one: the k-sums are ordered differently, so scores agree to
    # fp32 rounding (ids and lengths exactly)
    np.testing.assert_allclose(big["scores"][:6].cpu().numpy(), small["scores"].cpu().numpy(), rtol=0, atol=2e-5)
    ids, lens, scores = big["ids"].cpu().numpy(), big["lens"].cpu().numpy(), big["scores"].cpu().numpy()
    assert np.isfinite(scores).all() and (scores[:, 0] >= scores[:, 1]).all()
    for i in range(B):
        for n in range(2):
            _check_grammar(ids[i, n], int(lens[i, n]))
    assert len(set(lens[:, 0].tolist())) > 10              # images end at many different steps
    eng.close()


def test_greedy_bs256_shard_prefix_invariance_and_grammar():
    from molnextr_b200.engine import Engine
    dec = synth.decoder_state(0, "sensitised")
    B = 256
    feats = seeded_features(78, B, 144).cuda()
    eng = Engine({"decoder": dec, "encoder": None}, max_batch=B)
    eng.decode_greedy(feats)
    big, dt = _timed(lambda: eng.decode_greedy(feats))
    print(f"greedy bs=256 (one GPU's shard of configs[3]): {dt * 1e3:.1f} ms, {eng.last_decode_steps()} steps, {B / dt:.0f} img/s")
    # auto path at 256 rows = the throughput kernel as two launches (15 resident clusters = 240 rows, then the last 16 rows,
    # whose row ranks come from the final row_state words of the first launch)
    assert int(eng.time_kernel(1003, 1)) == 6
    # the same rows through the persistent cluster kernel in a small batch (a different code path)
    small = eng.decode_greedy(feats[:7])
    torch.cuda.synchronize()
    assert torch.equal(big["lens"][:7], small["lens"])
    assert torch.equal(big["ids"][:7], small["ids"])
    # ... and ALL 256 rows through the multi-kernel graph path
    eng.set_decode_path("graph")
    ref, dtg = _timed(lambda: eng.decode_greedy(feats))
    eng.set_decode_path("auto")
    print(f"greedy bs=256 on the multi-kernel graph path: {dtg * 1e3:.1f} ms")
    assert torch.equal(big["lens"], ref["lens"]) and torch.equal(big["ids"], ref["ids"])
    ids, lens = big["ids"].cpu().numpy(), big["lens"].cpu().numpy()
    for i in range(B):
        _check_grammar(ids[i], int(lens[i]))
    atom_idx, n_atoms = eng.atom_indices(big["ids"], big["lens"])
    edges = eng.edges(atom_idx, n_atoms)
    torch.cuda.synchronize()
    na = n_atoms.cpu().numpy()
    ed = edges.cpu().numpy()
    assert (na >= 0).all() and (na <= 160).all()
    for i in range(0, B, 17):                                # bond classes are symmetric apart from the 5 <-> 6 swap
        k = int(na[i])
        e = ed[i, :k, :k].astype(int)
        sw = e.T.copy()
        sw[e.T == 5], sw[e.T == 6] = 6, 5
        off = ~np.eye(k, dtype=bool)
        assert (e[off] == sw[off]).all()
    eng.close()


def test_highres_bs8_1024_prefix_invariance():
    """configs[4]: bs = 8 at 1024 x 1024 (S = 1024 memory positions, Swin windows padded at every stage)."""
    from molnextr_b200.engine import Engine
    from tests.helpers import seeded_images
    ck = synth.synthetic_checkpoint(0, "sensitised")
    eng = Engine(ck, max_batch=8, max_height=1024, max_width=1024)
    x = seeded_images(9, 8, 1024, 1024).cuda()
    eng.predict(x)
    big, dt = _timed(lambda: eng.predict(x))
    print(f"bs=8 1024x1024: {dt * 1e3:.1f} ms, {eng.last_decode_steps()} steps, {8 / dt:.1f} img/s")
    small = eng.predict(x[:2])
    torch.cuda.synchronize()
    for k in ("ids", "lens", "n_atoms", "atom_idx"):
        assert torch.equal(big[k][:2], small[k]), k
    ids, lens = big["ids"].cpu().numpy(), big["lens"].cpu().numpy()
    for i in range(8):
        _check_grammar(ids[i], int(lens[i]))
    eng.close()
