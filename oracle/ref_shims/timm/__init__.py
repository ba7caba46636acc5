"""Test-infrastructure shim: the handful of timm==0.4.12 helpers the reference Swin file
imports (MolNexTR/models/transformers.py:12-16) plus `create_model` (MolNexTR/components.py:129),
restated so the reference's own Swin-B definition runs here.  Never imported by the product."""
from .models.registry import _REGISTRY


def create_model(model_name, pretrained=False, **kwargs):
    if model_name not in _REGISTRY:
        raise RuntimeError("Unknown model (%s)" % model_name)
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    return _REGISTRY[model_name](pretrained=pretrained, **kwargs)
