"""Deterministic synthetic checkpoints in the reference `.pth` schema.

No trained checkpoint is reachable offline (the reference downloads `molnextr_best.pth`,
MolNexTR/molnextr.py:129-143), so benchmarks and parity fixtures use a per-tensor seeded
checkpoint with the exact key/shape schema `torch.save` writes at main.py:389-398:
`{'encoder': state_dict, 'decoder': state_dict, 'args': {...}}`.

Variants
--------
sensitised : O(1)-gain weights so features depend on the image, the greedy stream visits
             many tokens and rows terminate at different steps (exercises compaction and
             the row-rank positional-encoding rule).
fixed480   : the same weights with the <eos> logit bias pushed to -30 so no row ever emits
             <eos>: every row runs the full 480 steps (fixed work, used by bench.py).
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import torch

SWIN_DEPTHS = (2, 2, 18, 2)
SWIN_HEADS = (4, 8, 16, 32)
SWIN_DIM = 128
CONVNEXT_DEPTHS = (3, 3, 27, 3)
CONVNEXT_DIMS = (128, 256, 512, 1024)
DEC_LAYERS = 6
DEC_DIM = 256
DEC_FF = 1024
VOCAB = 229
EDGE_CLASSES = 7
EOS_ID = 2

CKPT_ARGS = {"formats": ["chartok_coords", "edges"], "input_size": 384, "coord_bins": 64, "sep_xy": True}


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _normal(name, shape, std, seed):
    return torch.randn(shape, generator=_gen(name, seed), dtype=torch.float32) * std


def _linear(sd, prefix, n_out, n_in, seed, gain=1.0, bias=True, bias_std=0.05):
    sd[prefix + ".weight"] = _normal(prefix + ".weight", (n_out, n_in), gain / math.sqrt(n_in), seed)
    if bias:
        sd[prefix + ".bias"] = _normal(prefix + ".bias", (n_out,), bias_std, seed)


def _layernorm(sd, prefix, n, seed):
    sd[prefix + ".weight"] = 1.0 + _normal(prefix + ".weight", (n,), 0.1, seed)
    sd[prefix + ".bias"] = _normal(prefix + ".bias", (n,), 0.1, seed)


def relative_position_index(ws: int = 12) -> torch.Tensor:
    """The (ws*ws, ws*ws) index buffer the reference registers (transformers.py:127-137)."""
    coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def swin_b_state(seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Encoder state-dict for the Swin-B the reference builds (prefix `transformer.`)."""
    sd = OrderedDict()
    p = "transformer."
    sd[p + "patch_embed.proj.weight"] = _normal(p + "patch_embed.proj.weight", (SWIN_DIM, 3, 4, 4), 1 / math.sqrt(48), seed)
    sd[p + "patch_embed.proj.bias"] = _normal(p + "patch_embed.proj.bias", (SWIN_DIM,), 0.05, seed)
    _layernorm(sd, p + "patch_embed.norm", SWIN_DIM, seed)
    rpi = relative_position_index(12)
    for s, (depth, heads) in enumerate(zip(SWIN_DEPTHS, SWIN_HEADS)):
        c = SWIN_DIM << s
        for j in range(depth):
            b = f"{p}layers.{s}.blocks.{j}."
            _layernorm(sd, b + "norm1", c, seed)
            sd[b + "attn.relative_position_bias_table"] = _normal(b + "attn.relative_position_bias_table", (529, heads), 0.5, seed)
            sd[b + "attn.relative_position_index"] = rpi.clone()
            _linear(sd, b + "attn.qkv", 3 * c, c, seed)
            _linear(sd, b + "attn.proj", c, c, seed, gain=0.7)
            _layernorm(sd, b + "norm2", c, seed)
            _linear(sd, b + "mlp.fc1", 4 * c, c, seed)
            _linear(sd, b + "mlp.fc2", c, 4 * c, seed, gain=0.7)
        if s < 3:
            d = f"{p}layers.{s}.downsample."
            _linear(sd, d + "reduction", 2 * c, 4 * c, seed, bias=False)
            _layernorm(sd, d + "norm", 4 * c, seed)
    _layernorm(sd, p + "norm", 1024, seed)
    return sd


def convnext_b_state(seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Encoder state-dict for timm ConvNeXt-B `forward_features` (prefix `cnn.`), the encoder
    `north_star` names; the reference's branch for it is dead code (components.py:121-126)."""
    sd = OrderedDict()
    p = "cnn."
    sd[p + "stem.0.weight"] = _normal(p + "stem.0.weight", (128, 3, 4, 4), 1 / math.sqrt(48), seed)
    sd[p + "stem.0.bias"] = _normal(p + "stem.0.bias", (128,), 0.05, seed)
    _layernorm(sd, p + "stem.1", 128, seed)
    for s, (depth, c) in enumerate(zip(CONVNEXT_DEPTHS, CONVNEXT_DIMS)):
        if s > 0:
            cin = CONVNEXT_DIMS[s - 1]
            _layernorm(sd, f"{p}stages.{s}.downsample.0", cin, seed)
            sd[f"{p}stages.{s}.downsample.1.weight"] = _normal(f"{p}stages.{s}.downsample.1.weight", (c, cin, 2, 2), 1 / math.sqrt(4 * cin), seed)
            sd[f"{p}stages.{s}.downsample.1.bias"] = _normal(f"{p}stages.{s}.downsample.1.bias", (c,), 0.05, seed)
        for j in range(depth):
            b = f"{p}stages.{s}.blocks.{j}."
            sd[b + "conv_dw.weight"] = _normal(b + "conv_dw.weight", (c, 1, 7, 7), 1 / 7.0, seed)
            sd[b + "conv_dw.bias"] = _normal(b + "conv_dw.bias", (c,), 0.05, seed)
            _layernorm(sd, b + "norm", c, seed)
            _linear(sd, b + "mlp.fc1", 4 * c, c, seed)
            _linear(sd, b + "mlp.fc2", c, 4 * c, seed)
            sd[b + "gamma"] = 0.5 + _normal(b + "gamma", (c,), 0.05, seed)
    return sd


def positional_encoding_table(max_len: int = 5000, dim: int = DEC_DIM) -> torch.Tensor:
    """The `pe` buffer of the reference (models/embedding.py:30-36), shape (max_len,1,dim)."""
    pe = torch.zeros(max_len, dim)
    position = torch.arange(0, max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dim, 2, dtype=torch.float) * -(math.log(10000.0) / dim))
    pe[:, 0::2] = torch.sin(position.float() * div_term)
    pe[:, 1::2] = torch.cos(position.float() * div_term)
    return pe.unsqueeze(1)


# Token classes used to give the random decoder a molecule-like bigram prior.
_ATOM_CHARS = "CNOFSPIBcnos"
_BOND_CHARS = "=#()12"


def _token_class(i: int, itos) -> int:
    """0 sos, 1 atom char, 2 x-bin, 3 y-bin, 4 bond/branch/ring char, 5 everything else."""
    if i == 1:
        return 0
    if i >= 101 + 64:
        return 3
    if i >= 101:
        return 2
    ch = itos.get(i, "")
    if ch in _ATOM_CHARS and len(ch) == 1:
        return 1
    if ch in _BOND_CHARS and len(ch) == 1:
        return 4
    return 5


def _token_grammar_prior(sd, p, seed, kappa=1.0, kappa2=0.3, n_hash=8):
    """Embedding table + output layer with a class-level bigram prior.

    A purely random decoder collapses to a fixed point (the reference has no step-dependent
    positional signal).  To get streams that look like the real model's
    (`C x y C x y = O x y ... <eos>`: atoms with coordinates, rows ending at different
    steps, image-dependent choices) every embedding carries (i) a shared per-class vector
    and (ii) an `n_hash`-bit sign code along fixed hash directions; the output layer reads
    both back: after an atom char favour X bins, after a Y bin favour atoms / bond chars /
    <eos>, after a bond char favour atoms, and within the favoured class the hash code of
    the exact previous token (e.g. which coordinate bin) re-ranks the candidates, so the
    coordinate bins act as a pseudo-random state.  The random part of both matrices and
    every attention/FFN layer supply the image- and history-dependent variation."""
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "vocab", "vocab_chars.json")) as f:
        itos = {i: s for s, i in json.load(f).items()}
    cls = torch.tensor([_token_class(i, itos) for i in range(VOCAB)])
    cvec, hvec = _basis(p, seed, n_hash)
    code = torch.sign(_normal(p + "hash_codes", (VOCAB, n_hash), 1.0, seed))
    indiv = _normal(p + "embeddings.make_embedding.emb_luts.0.weight", (VOCAB, DEC_DIM), 1.0, seed)
    emb = 2.0 * (0.7 * cvec[cls] + 0.5 * (code @ hvec) * (16.0 / math.sqrt(n_hash)) + 0.4 * indiv) / 16.0
    emb[0].zero_()  # padding_idx row (models/embedding.py:141)
    sd[p + "embeddings.make_embedding.emb_luts.0.weight"] = emb
    _linear(sd, p + "output_layer", VOCAB, DEC_DIM, seed, gain=1.0, bias_std=0.3)
    # score[next token | prev class]
    S = torch.zeros(VOCAB, 6)
    atoms = (cls == 1)
    S[atoms, 0] = 0.8
    S[57, 0] = 1.2                      # 'C' first
    S[cls == 2, 1] = 1.0                # atom -> x-bin
    S[atoms, 3] = 1.0
    S[57, 3] = 1.1
    S[cls == 4, 3] = 0.9
    S[EOS_ID, 3] = 1.5
    S[atoms, 4] = 1.0
    S[atoms, 5] = 1.0
    unit = cvec / cvec.norm(dim=1, keepdim=True)
    G = _normal(p + "hash_readout", (VOCAB, n_hash), 1.0, seed)
    sd[p + "output_layer.weight"] = sd[p + "output_layer.weight"] + kappa * (S @ unit) + kappa2 * (G @ hvec)


def _basis(p, seed, n_hash=8):
    """Orthogonal class vectors (norm 16, i.e. unit per-dim scale) and unit hash directions."""
    raw = _normal(p + "class_and_hash_vectors", (6 + n_hash, DEC_DIM), 1.0, seed)
    q, _ = torch.linalg.qr(raw.t())          # (256, 6+n_hash), orthonormal columns
    q = q.t().contiguous()
    return q[:6] * 16.0, q[6:]


def _eos_clock(sd, p, seed, sink=300.0, beta=40.0, kappa3=1.1):
    """A crafted slow clock so rows terminate at different, finite steps.

    Head 7 of the last layer's self-attention gets a constant query (zero projection, bias
    only) and one key/value channel that reads the <sos> class vector, which only position 0
    carries: position 0 becomes an attention sink with weight ~ sink/(sink+t).
    `final_linear` writes that weight along a fixed direction and the <eos> logit is
    penalised along it, so <eos> becomes competitive only as t grows -- when exactly depends
    on the row's own stream, hence on the image."""
    l = f"{p}decoder.transformer_layers.{DEC_LAYERS - 1}.self_attn."
    cvec, _ = _basis(p, seed)
    c0 = cvec[0] / cvec[0].norm()
    ue = _normal(p + "eos_clock_dir", (DEC_DIM,), 1.0, seed)
    ue = ue / ue.norm()
    ch = 7 * 32
    proj0 = 7.0   # ~ <LN1(x_pos0), c0> for the synthetic weights
    sd[l + "linear_query.weight"][ch:ch + 32].zero_()
    sd[l + "linear_query.bias"][ch:ch + 32].zero_()
    sd[l + "linear_query.bias"][ch] = math.sqrt(32.0)
    sd[l + "linear_keys.weight"][ch] = c0 * (math.log(sink) / proj0)
    sd[l + "linear_keys.bias"][ch] = 0.0
    sd[l + "linear_values.weight"][ch] = c0 / proj0
    sd[l + "linear_values.bias"][ch] = 0.0
    sd[l + "final_linear.weight"][:, ch] = beta * ue
    sd[p + "output_layer.weight"][EOS_ID] -= kappa3 * ue


def decoder_state(seed: int = 0, variant: str = "sensitised", encoder_dim: int = 1024) -> "OrderedDict[str, torch.Tensor]":
    sd = OrderedDict()
    p = "decoder.chartok_coords."
    _linear(sd, p + "enc_trans_layer.0", DEC_DIM, encoder_dim, seed)
    _layernorm(sd, p + "decoder.layer_norm", DEC_DIM, seed)
    for l in range(DEC_LAYERS):
        b = f"{p}decoder.transformer_layers.{l}."
        for att in ("self_attn", "context_attn"):
            for lin in ("linear_keys", "linear_values", "linear_query"):
                _linear(sd, f"{b}{att}.{lin}", DEC_DIM, DEC_DIM, seed, gain=1.5)
            _linear(sd, f"{b}{att}.final_linear", DEC_DIM, DEC_DIM, seed, gain=(1.0 if att == "context_attn" else 0.4))
        _linear(sd, b + "feed_forward.w_1", DEC_FF, DEC_DIM, seed)
        _linear(sd, b + "feed_forward.w_2", DEC_DIM, DEC_FF, seed, gain=0.5)
        _layernorm(sd, b + "feed_forward.layer_norm", DEC_DIM, seed)
        _layernorm(sd, b + "layer_norm_1", DEC_DIM, seed)
        _layernorm(sd, b + "layer_norm_2", DEC_DIM, seed)
    _token_grammar_prior(sd, p, seed)
    sd[p + "embeddings.make_embedding.pe.pe"] = positional_encoding_table()
    _linear(sd, "decoder.edges.mlp.0", DEC_DIM, 2 * DEC_DIM, seed, gain=4.0)
    _linear(sd, "decoder.edges.mlp.2", EDGE_CLASSES, DEC_DIM, seed, gain=3.0, bias_std=0.5)
    sd["decoder.edges.mlp.2.bias"][0] += 3.0   # "no bond" is the most common class
    # atom hiddens share large common components (class / hash / clock directions); remove
    # them from the first bond-head layer so pairs of atoms actually differ
    cv, hv = _basis(p, seed)
    ue = _normal(p + "eos_clock_dir", (DEC_DIM,), 1.0, seed)
    common = torch.cat([cv / 16.0, hv, (ue / ue.norm()).unsqueeze(0)], 0)
    qc, _ = torch.linalg.qr(common.t())
    w0 = sd["decoder.edges.mlp.0.weight"]
    for half in (slice(0, DEC_DIM), slice(DEC_DIM, 2 * DEC_DIM)):
        w0[:, half] -= (w0[:, half] @ qc) @ qc.t()
    _eos_clock(sd, p, seed)
    bias = sd[p + "output_layer.bias"]
    if variant == "sensitised":
        pass
    elif variant == "fixed480":
        bias[EOS_ID] = -100.0  # never selected: every row runs all 480 steps
    else:
        raise ValueError(f"unknown variant {variant!r}")
    return sd


def synthetic_checkpoint(seed: int = 0, variant: str = "sensitised", encoder: str = "swin_base") -> dict:
    """A full checkpoint dict in the reference schema (main.py:389-398)."""
    if encoder.startswith("swin"):
        enc = swin_b_state(seed)
    elif encoder.startswith("convnext"):
        enc = convnext_b_state(seed)
    else:
        raise ValueError(encoder)
    args = dict(CKPT_ARGS)
    if not encoder.startswith("swin"):
        args["encoder"] = encoder
    return {"encoder": enc, "decoder": decoder_state(seed, variant), "args": args, "global_step": 0}
