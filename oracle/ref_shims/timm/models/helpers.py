"""build_model_with_cfg / overlay_external_default_cfg restated to the extent the reference
uses them (MolNexTR/models/transformers.py:519-541): construct the class, attach the cfg."""


def overlay_external_default_cfg(default_cfg, kwargs):
    external = kwargs.pop("external_default_cfg", None)
    if external:
        default_cfg.pop("url", None)
        default_cfg.update(external)


def build_model_with_cfg(model_cls, variant, pretrained, default_cfg, pretrained_filter_fn=None,
                         pretrained_strict=True, **kwargs):
    kwargs.pop("pretrained_custom_load", None)
    kwargs.pop("features_only", None)
    assert not pretrained, "no network: pretrained weights are unavailable"
    model = model_cls(**kwargs)
    model.default_cfg = dict(default_cfg)
    return model
