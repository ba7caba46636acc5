"""CPU: external anchor for the ConvNeXt-B restatement.

Nothing in the reference can execute its ConvNeXt branch (timm 0.4.12 has no ConvNeXt, F2), so
`oracle/restate.convnext_b_features` restates timm's published algorithm.  torchvision ships an independent
implementation of the SAME published architecture (`torchvision.models.convnext_base`: stem conv4x4/4 + LayerNorm2d,
per stage [LayerNorm2d + conv2x2/2] then CNBlocks = dwconv7x7 -> LayerNorm -> Linear -> GELU -> Linear -> layer_scale
-> residual).  Loading the same seeded weights into it (timm key names -> torchvision key names) and comparing
`features(x)` pins the restatement to an implementation this repo did not write."""
import pytest
import torch

from molnextr_b200 import synth
from oracle import restate

tv = pytest.importorskip("torchvision")


def _to_torchvision(sd):
    out = {}
    for k, v in sd.items():
        k = k.replace("cnn.", "", 1)
        p = k.split(".")
        if p[0] == "stem":
            out[f"features.0.{p[1]}.{p[2]}"] = v
        elif p[0] == "stages" and p[2] == "downsample":
            out[f"features.{2 * int(p[1])}.{p[3]}.{p[4]}"] = v
        elif p[0] == "stages" and p[2] == "blocks":
            base = f"features.{2 * int(p[1]) + 1}.{p[3]}."
            if p[4] == "gamma":
                out[base + "layer_scale"] = v.reshape(-1, 1, 1)
            else:
                idx = {"conv_dw": 0, "norm": 2}.get(p[4]) if p[4] != "mlp" else {"fc1": 3, "fc2": 5}[p[5]]
                out[base + f"block.{idx}.{p[-1]}"] = v
    return out


@pytest.mark.parametrize("hw", [(64, 64), (96, 160)])
def test_restated_convnext_b_equals_torchvision(hw):
    sd = synth.convnext_b_state(3)
    model = tv.models.convnext_base(weights=None).eval()
    mapped = _to_torchvision({k: v for k, v in sd.items() if not k.startswith("cnn.head")})
    feat_keys = {k for k in model.state_dict() if k.startswith("features.")}
    assert set(mapped) == feat_keys, (sorted(set(mapped) ^ feat_keys)[:6])
    model.load_state_dict(mapped, strict=False)
    x = torch.randn((2, 3) + hw, generator=torch.Generator().manual_seed(11))
    with torch.no_grad():
        ref = model.features(x).permute(0, 2, 3, 1).reshape(2, -1, 1024)
        got = restate.convnext_b_features(sd, x)
    assert got.shape == ref.shape
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-5 * max(1.0, scale), (err, scale)
