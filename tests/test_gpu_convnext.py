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


def test_convnext_1024_highres_parity_and_timing():
    """BASELINE.json configs[4] ("bs=8 1024x1024 ... encoder dwconv/TMA stress"): parity of one 1024x1024 image against
    the oracle (pinned to torchvision's ConvNeXt-B by tests/test_oracle_convnext_torchvision.py), rows of a bs=8 batch
    equal to the same image encoded alone, and the bs=8 encoder time with the dwconv share (printed with -s)."""
    from molnextr_b200.engine import Engine
    from oracle import restate
    ck = synth.synthetic_checkpoint(0, "sensitised", encoder="convnext_base")
    eng = Engine(ck, max_batch=8, max_height=1024, max_width=1024)
    x = seeded_images(41, 8, 1024, 1024)
    with torch.no_grad():
        ref = restate.convnext_b_features(ck["encoder"], x[:1])
    xd = x.cuda()
    feats = eng.encode(xd)
    assert feats.shape == (8, 1024, 1024)
    one = eng.encode(xd[:1].contiguous())
    err = (one.cpu() - ref).abs()
    scale = float(ref.abs().mean())
    print(f"convnext 1024^2 feature |err| max {err.max():.4f} mean {err.mean():.5f} (mean |ref| {scale:.3f})")
    assert float(err.max()) <= 0.15 * max(1.0, scale) and float(err.mean()) <= 0.015 * max(1.0, scale)
    assert torch.equal(feats[:1], one), "row 0 of the batch differs from the same image alone"
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        eng.encode(xd)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    dw = [eng.time_kernel(101 + s, 10) for s in range(4)]
    dw_ms = 3 * dw[0] + 3 * dw[1] + 27 * dw[2] + 3 * dw[3]
    print(f"convnext-B bs=8 1024^2: encoder {ms:.2f} ms = {8 / ms * 1e3:.0f} img/s; dwconv+LN {dw_ms:.2f} ms "
          f"(per stage {', '.join(f'{1e3 * d:.0f} us' for d in dw)})")
    eng.close()
