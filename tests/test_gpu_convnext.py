"""GPU parity of the ConvNeXt-B encoder (fused dwconv7x7+LayerNorm kernel, tcgen05 pointwise GEMMs)
against the CPU oracle's restatement of timm ConvNeXt-B `forward_features`.

PARITY UNPINNED by the reference: its ConvNeXt branch is dead code, so the oracle here is
anchored only on timm's published algorithm.  Tolerance as for Swin-B (bf16 GEMM operands)."""
import pytest
import torch

from molnextr_b200 import synth
from tests.helpers import seeded_images

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hw", [(384, 384), (256, 320)])
def test_convnext_features_match_oracle(hw):
    from molnextr_b200.engine import Engine
    from oracle import restate
    ck = synth.synthetic_checkpoint(0, "sensitised", encoder="convnext_base")
    eng = Engine(ck, max_batch=2, max_height=hw[0], max_width=hw[1])
    x = seeded_images(31, 2, hw[0], hw[1])
    with torch.no_grad():
        ref = restate.convnext_b_features(ck["encoder"], x)
    feats = eng.encode(x.cuda()).cpu()
    assert feats.shape == ref.shape == (2, (hw[0] // 32) * (hw[1] // 32), 1024)
    err = (feats - ref).abs()
    scale = float(ref.abs().mean())
    print(f"convnext feature |err| max {err.max():.4f} mean {err.mean():.5f} (mean |ref| {scale:.3f})")
    assert float(err.max()) <= 0.15 * max(1.0, scale) and float(err.mean()) <= 0.015 * max(1.0, scale)
    out = eng.predict(x.cuda())
    assert int(out["lens"].min()) >= 1
    eng.close()
