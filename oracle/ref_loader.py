"""Import the reference's OWN modules from /root/reference (read-only, build container) or, on the GPU box,
from the git-ignored copy staged by oracle/stage_ref.py under baseline/_ref.

Used by oracle/make_golden.py and tests/test_oracle_vs_reference.py to pin oracle/restate.py, and by
`bench.py --impl reference` (the reference arm).  The GPU tests and smoke() never call into this file.

Recipe: register a stub package object `MolNexTR` whose __path__ is the
reference package directory (skipping its __init__, which needs pystow/cv2/rdkit), put the
restated third-party classes in oracle/ref_shims first on sys.path, then import the
reference's components/tokenization unmodified."""
from __future__ import annotations

import argparse
import os
import sys
import types
import warnings

# /root/reference in the build container; on the GPU box the copy staged by oracle/stage_ref.py (git-ignored)
_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
REFERENCE_ROOT = os.environ.get("MOLNEXTR_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isdir("/root/reference/MolNexTR") else _STAGED)
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "MolNexTR"))


def _install():
    if "MolNexTR" in sys.modules and getattr(sys.modules["MolNexTR"], "_oracle_stub", False):
        return
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    pkg = types.ModuleType("MolNexTR")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "MolNexTR")]
    pkg._oracle_stub = True
    sys.modules["MolNexTR"] = pkg


def reference_args(encoder: str = "swin_base", **over):
    """Namespace mirroring molnextr._get_args (MolNexTR/model.py:50-81) + checkpoint args."""
    ns = argparse.Namespace(
        encoder=encoder, decoder="transformer", trunc_encoder=False, no_pretrained=True,
        use_checkpoint=True, dropout=0.5, embed_dim=256, enc_pos_emb=False, dec_num_layers=6,
        dec_hidden_size=256, dec_attn_heads=8, dec_num_queries=128, hidden_dropout=0.1,
        attn_dropout=0.1, max_relative_positions=0, continuous_coords=False,
        compute_confidence=False, input_size=384, vocab_file=None, coord_bins=64, sep_xy=True,
        formats=["chartok_coords", "edges"])
    for k, v in over.items():
        setattr(ns, k, v)
    return ns


def build_reference(ckpt: dict, **over):
    """(encoder, decoder, tokenizer) built by the reference's constructors and loaded the way
    molnextr._get_model does (MolNexTR/model.py:83-95), but with strict=True.  `over` overrides
    fields of the argument namespace (e.g. compute_confidence=True)."""
    _install()
    warnings.filterwarnings("ignore")
    from MolNexTR.components import Encoder, Decoder  # noqa
    from MolNexTR.tokenization import get_tokenizer  # noqa
    args = reference_args(**over)
    for k, v in ckpt["args"].items():
        setattr(args, k, v)
    tokenizer = get_tokenizer(args)
    encoder = Encoder(args, pretrained=False)
    args.encoder_dim = encoder.n_features
    decoder = Decoder(args, tokenizer)
    strip = lambda sd: {k.replace("module.", ""): v for k, v in sd.items()}
    encoder.load_state_dict(strip(ckpt["encoder"]), strict=True)
    decoder.load_state_dict(strip(ckpt["decoder"]), strict=True)
    return encoder.eval(), decoder.eval(), tokenizer
