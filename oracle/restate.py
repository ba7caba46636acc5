"""CPU oracle: a plain torch-fp32 restatement of the reference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under `molnextr_b200/` may import this module; only
`tests/`, `__graft_entry__.smoke()` and bench.py's CPU-baseline / `--impl reference` legs do.
It restates, function by function, what the reference computes between
`features, hiddens = self.encoder(images)` and the prediction list returned by
`self.decoder.decode(features, hiddens)` (MolNexTR/model.py:106-108), reading weights straight
from a checkpoint in the reference's own `.pth` schema.

Pinning status (see oracle/README.md and tests/test_oracle_vs_reference.py):
  * Swin-B encoder, decoder step, greedy state machine, bond head, tokenizer: PINNED against
    the reference's own modules imported in the build container (third-party onmt/timm classes
    restated under oracle/ref_shims) and against the fixtures that run wrote to tests/golden/.
  * ConvNeXt-B encoder: "parity unpinned" -- the reference's ConvNeXt branch cannot execute in
    any environment; this file restates timm's published ConvNeXt-B
    `forward_features` and nothing in the reference can confirm it.
  * Beam search: "parity unpinned" -- unrunnable in the reference.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

PAD_ID, SOS_ID, EOS_ID, UNK_ID, MASK_ID = 0, 1, 2, 3, 4
SWIN_DEPTHS = (2, 2, 18, 2)
SWIN_HEADS = (4, 8, 16, 32)
WINDOW = 12
CONVNEXT_DEPTHS = (3, 3, 27, 3)
DEC_LAYERS = 6
DEC_HEADS = 8
DEC_DIM = 256

SD = Dict[str, torch.Tensor]


def _strip(sd: SD) -> SD:
    """`module.` prefixes from DDP are dropped on load (MolNexTR/model.py:25-26)."""
    return {k.replace("module.", ""): v for k, v in sd.items()}


# --------------------------------------------------------------------------------------
# Swin-B encoder  (MolNexTR/models/transformers.py)
# --------------------------------------------------------------------------------------
def _window_partition(x: torch.Tensor, ws: int) -> torch.Tensor:
    # transformers.py:68-80
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)


def _window_reverse(win: torch.Tensor, ws: int, H: int, W: int) -> torch.Tensor:
    # transformers.py:83-97
    B = win.shape[0] // ((H // ws) * (W // ws))
    x = win.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def swin_shift_mask(Hp: int, Wp: int, ws: int, shift: int, device=None) -> Optional[torch.Tensor]:
    """(nW, ws*ws, ws*ws) additive mask, 0 / -100  (transformers.py:220-243)."""
    if shift == 0:
        return None
    img = torch.zeros((1, Hp, Wp, 1), device=device)
    cnt = 0
    for h in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for w in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, h, w, :] = cnt
            cnt += 1
    mw = _window_partition(img, ws).view(-1, ws * ws)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)


def _swin_block(sd: SD, p: str, x: torch.Tensor, H: int, W: int, heads: int, shift: int) -> torch.Tensor:
    # SwinTransformerBlock.forward, transformers.py:245-292
    B, L, C = x.shape
    ws = WINDOW
    shortcut = x
    x = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5).view(B, H, W, C)
    pad_r = (ws - W % ws) % ws
    pad_b = (ws - H % ws) % ws
    x = F.pad(x, (0, 0, 0, pad_r, 0, pad_b))
    Hp, Wp = x.shape[1], x.shape[2]
    if shift > 0:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
    xw = _window_partition(x, ws).view(-1, ws * ws, C)
    mask = swin_shift_mask(Hp, Wp, ws, shift, device=x.device)
    # WindowAttention.forward, transformers.py:147-178
    B_, N, _ = xw.shape
    hd = C // heads
    qkv = F.linear(xw, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
    qkv = qkv.reshape(B_, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * (hd ** -0.5)
    attn = q @ k.transpose(-2, -1)
    idx = sd[p + "attn.relative_position_index"].view(-1)
    bias = sd[p + "attn.relative_position_bias_table"][idx].view(N, N, -1).permute(2, 0, 1).contiguous()
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.view(B_ // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, heads, N, N)
    attn = torch.softmax(attn, dim=-1)
    xo = (attn @ v).transpose(1, 2).reshape(B_, N, C)
    xo = F.linear(xo, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    x = _window_reverse(xo.view(-1, ws, ws, C), ws, Hp, Wp)
    if shift > 0:
        x = torch.roll(x, shifts=(shift, shift), dims=(1, 2))
    if pad_r > 0 or pad_b > 0:
        x = x[:, :H, :W, :].contiguous()
    x = shortcut + x.view(B, H * W, C)
    y = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    y = F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    y = F.gelu(y)
    y = F.linear(y, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x + y


def _patch_merging(sd: SD, p: str, x: torch.Tensor, H: int, W: int):
    # PatchMerging.forward, transformers.py:310-336
    B, L, C = x.shape
    x = x.view(B, H, W, C)
    if H % 2 == 1 or W % 2 == 1:
        x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
    x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
    H2, W2 = x.shape[1], x.shape[2]
    x = x.view(B, -1, 4 * C)
    x = F.layer_norm(x, (4 * C,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    return F.linear(x, sd[p + "reduction.weight"]), H2, W2


def swin_b_features(enc_sd: SD, images: torch.Tensor, return_stage_outputs: bool = False):
    """Vision_Transformer.forward (transformers.py:504-515) for `swin_base` (:547-551).
    images: (B,3,H,W) fp32 normalised -> (B, (H/32)*(W/32), 1024)."""
    sd = _strip(enc_sd)
    t = "transformer."
    x = images
    # PatchEmbed.forward, transformers.py:405-419
    _, _, H, W = x.shape
    if W % 4 != 0:
        x = F.pad(x, (0, 4 - W % 4))
    if H % 4 != 0:
        x = F.pad(x, (0, 0, 0, 4 - H % 4))
    x = F.conv2d(x, sd[t + "patch_embed.proj.weight"], sd[t + "patch_embed.proj.bias"], stride=4)
    H, W = x.shape[2], x.shape[3]
    x = x.flatten(2).transpose(1, 2)
    x = F.layer_norm(x, (128,), sd[t + "patch_embed.norm.weight"], sd[t + "patch_embed.norm.bias"], 1e-5)
    stage_out = []
    for s, (depth, heads) in enumerate(zip(SWIN_DEPTHS, SWIN_HEADS)):
        for j in range(depth):
            shift = 0 if j % 2 == 0 else WINDOW // 2
            x = _swin_block(sd, f"{t}layers.{s}.blocks.{j}.", x, H, W, heads, shift)
        stage_out.append(x)
        if s < 3:
            x, H, W = _patch_merging(sd, f"{t}layers.{s}.downsample.", x, H, W)
    x = F.layer_norm(x, (1024,), sd[t + "norm.weight"], sd[t + "norm.bias"], 1e-5)
    if return_stage_outputs:
        return x, stage_out
    return x


# --------------------------------------------------------------------------------------
# ConvNeXt-B encoder (timm ConvNeXt.forward_features; the reference branch at
# MolNexTR/components.py:121-126,163-166 is dead code -> PARITY UNPINNED)
# --------------------------------------------------------------------------------------
def _ln2d(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float) -> torch.Tensor:
    return F.layer_norm(x.permute(0, 2, 3, 1), (x.shape[1],), w, b, eps).permute(0, 3, 1, 2)


def convnext_b_features(enc_sd: SD, images: torch.Tensor) -> torch.Tensor:
    """stem -> 4 stages -> NHWC -> (B, HW, 1024), as `features.permute(0,2,3,1)` then the
    decoder's flatten (components.py:165, :207-209)."""
    sd = _strip(enc_sd)
    c = "cnn."
    x = F.conv2d(images, sd[c + "stem.0.weight"], sd[c + "stem.0.bias"], stride=4)
    x = _ln2d(x, sd[c + "stem.1.weight"], sd[c + "stem.1.bias"], 1e-6)
    for s, depth in enumerate(CONVNEXT_DEPTHS):
        if s > 0:
            p = f"{c}stages.{s}.downsample."
            x = _ln2d(x, sd[p + "0.weight"], sd[p + "0.bias"], 1e-6)
            x = F.conv2d(x, sd[p + "1.weight"], sd[p + "1.bias"], stride=2)
        for j in range(depth):
            p = f"{c}stages.{s}.blocks.{j}."
            C = x.shape[1]
            y = F.conv2d(x, sd[p + "conv_dw.weight"], sd[p + "conv_dw.bias"], padding=3, groups=C)
            y = y.permute(0, 2, 3, 1)
            y = F.layer_norm(y, (C,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
            y = F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
            y = F.gelu(y)
            y = F.linear(y, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
            y = y.permute(0, 3, 1, 2)
            x = x + y * sd[p + "gamma"].reshape(1, -1, 1, 1)
    B, C = x.shape[0], x.shape[1]
    return x.permute(0, 2, 3, 1).reshape(B, -1, C)


def encoder_features(enc_sd: SD, images: torch.Tensor) -> torch.Tensor:
    """Encoder.forward (components.py:162-174): backbone chosen by the state-dict prefix."""
    key = next(iter(_strip(enc_sd)))
    if key.startswith("transformer."):
        return swin_b_features(enc_sd, images)
    if key.startswith("cnn."):
        return convnext_b_features(enc_sd, images)
    raise ValueError(f"unrecognised encoder state-dict (first key {key!r})")


# --------------------------------------------------------------------------------------
# Decoder  (components.py:177-334, models/decoder.py, models/embedding.py, onmt MHA/FFN)
# --------------------------------------------------------------------------------------
_P = "decoder.chartok_coords."


def memory_bank(dec_sd: SD, features: torch.Tensor) -> torch.Tensor:
    # TransformerDecoderBase.enc_transform, components.py:206-216 (enc_pos_emb is off)
    sd = _strip(dec_sd)
    B, C = features.size(0), features.size(-1)
    return F.linear(features.view(B, -1, C), sd[_P + "enc_trans_layer.0.weight"], sd[_P + "enc_trans_layer.0.bias"])


def _shape(x: torch.Tensor) -> torch.Tensor:
    return x.view(x.size(0), -1, DEC_HEADS, DEC_DIM // DEC_HEADS).transpose(1, 2)


def _unshape(x: torch.Tensor) -> torch.Tensor:
    return x.transpose(1, 2).contiguous().view(x.size(0), -1, DEC_DIM)


def _mha(sd: SD, p: str, key, value, query, cache: dict, kind: str, mask=None):
    """onmt.modules.MultiHeadedAttention.forward with a layer cache (OpenNMT-py 2.2.0; call
    sites models/decoder.py:144-151 and :269-276)."""
    lin = lambda name, x: F.linear(x, sd[f"{p}{name}.weight"], sd[f"{p}{name}.bias"])
    if kind == "self":
        q, k, v = lin("linear_query", query), _shape(lin("linear_keys", query)), _shape(lin("linear_values", query))
        if cache["self_keys"] is not None:
            k = torch.cat((cache["self_keys"], k), dim=2)
            v = torch.cat((cache["self_values"], v), dim=2)
        cache["self_keys"], cache["self_values"] = k, v
    else:
        q = lin("linear_query", query)
        if cache["memory_keys"] is None:
            cache["memory_keys"] = _shape(lin("linear_keys", key))
            cache["memory_values"] = _shape(lin("linear_values", value))
        k, v = cache["memory_keys"], cache["memory_values"]
    q = _shape(q) / math.sqrt(DEC_DIM // DEC_HEADS)
    scores = torch.matmul(q, k.transpose(2, 3)).float()
    if mask is not None:
        scores = scores.masked_fill(mask.unsqueeze(1), -1e18)
    attn = torch.softmax(scores, dim=-1)
    ctx = _unshape(torch.matmul(attn, v))
    return lin("final_linear", ctx)


class DecoderState:
    """The per-layer cache dict of TransformerDecoder (_init_cache, models/decoder.py:482-486)."""

    def __init__(self):
        self.layers = [dict(memory_keys=None, memory_values=None, self_keys=None, self_values=None)
                       for _ in range(DEC_LAYERS)]

    def index_select(self, idx: torch.Tensor):
        # TransformerDecoderAR.map_state, components.py:337-347
        for c in self.layers:
            for k, v in c.items():
                if v is not None:
                    c[k] = v.index_select(0, idx)


def decoder_step(dec_sd: SD, tokens: torch.Tensor, mem: torch.Tensor, state: DecoderState) -> torch.Tensor:
    """One autoregressive step for the alive rows.  tokens (R,) int64, mem (R,S,256).
    Returns dec_out (R,256) = final-LayerNorm output (the `hidden` greedy search stores).

    Embedding: Embeddings.forward / PositionalEncoding.forward (models/embedding.py:236-255,
    :42-61) are fed a batch-first (R,1,256) tensor and no `step`, so row r receives pe[r]
     -- reproduced here on purpose.
    Layers: TransformerDecoderLayer._forward (models/decoder.py:224-279)."""
    sd = _strip(dec_sd)
    R = tokens.numel()
    emb_w = sd[_P + "embeddings.make_embedding.emb_luts.0.weight"]
    pe = sd[_P + "embeddings.make_embedding.pe.pe"]
    if pe.size(0) < R:
        raise RuntimeError(f"Sequence is {R} but PositionalEncoding is limited to {pe.size(0)}")
    x = emb_w[tokens].view(R, 1, DEC_DIM) * math.sqrt(DEC_DIM) + pe[:R]
    src_pad_mask = torch.zeros((R, 1, mem.size(1)), dtype=torch.bool, device=mem.device)
    for l in range(DEC_LAYERS):
        p = f"{_P}decoder.transformer_layers.{l}."
        cache = state.layers[l]
        n1 = F.layer_norm(x, (DEC_DIM,), sd[p + "layer_norm_1.weight"], sd[p + "layer_norm_1.bias"], 1e-6)
        x = _mha(sd, p + "self_attn.", n1, n1, n1, cache, "self") + x
        n2 = F.layer_norm(x, (DEC_DIM,), sd[p + "layer_norm_2.weight"], sd[p + "layer_norm_2.bias"], 1e-6)
        x = _mha(sd, p + "context_attn.", mem, mem, n2, cache, "context", mask=src_pad_mask) + x
        # onmt PositionwiseFeedForward.forward (activation 'gelu', components.py:203)
        h = F.layer_norm(x, (DEC_DIM,), sd[p + "feed_forward.layer_norm.weight"], sd[p + "feed_forward.layer_norm.bias"], 1e-6)
        h = F.gelu(F.linear(h, sd[p + "feed_forward.w_1.weight"], sd[p + "feed_forward.w_1.bias"]))
        x = F.linear(h, sd[p + "feed_forward.w_2.weight"], sd[p + "feed_forward.w_2.bias"]) + x
    x = F.layer_norm(x, (DEC_DIM,), sd[_P + "decoder.layer_norm.weight"], sd[_P + "decoder.layer_norm.bias"], 1e-6)
    return x.view(R, DEC_DIM)


def grammar_mask(tokens: torch.Tensor, offset: int, maxx: int, maxy: int) -> torch.Tensor:
    """CharTokenizer.get_output_mask (tokenization.py:383-392) for each INPUT token: True = blocked."""
    V = offset + maxx + maxy
    ids = torch.arange(V, device=tokens.device).unsqueeze(0)
    t = tokens.view(-1, 1)
    is_x = (t >= offset) & (t < offset + maxx)
    is_y = t >= offset + maxx
    return (is_x & (ids < offset + maxx)) | (is_y & (ids >= offset))


def greedy_decode(dec_sd: SD, features: torch.Tensor, max_len: int = 480,
                  grammar=(101, 64, 64), record_logprobs: bool = False,
                  forced_ids: Optional[torch.Tensor] = None, min_length: int = 1,
                  labels: Optional[torch.Tensor] = None):
    """TransformerDecoderAR.decode with GreedySearch (components.py:253-334,
    decoding/greedy_search.py:33-128, decode_strategy.py:4-62).

    Returns a list (one per image) of dicts: ids int64 (L,) incl. <eos> and excl. <sos>;
    logp fp32 (L,); hidden fp32 (L,256); score = exp(mean(logp)).  With `record_logprobs`
    each dict also has `logprobs` (L,V): the masked log-probabilities GreedySearch.advance saw.
    `forced_ids` (B,T) teacher-forces the chosen token (all rows must then share one length;
    used only to compare per-step log-probs at tensor level).
    `labels` (B,n) long = partial prediction (components.py:256-257,286-289,305,317-318,326-332;
    greedy_search.py:83-85): given tokens are fed as step inputs, MASK_ID positions are the model's;
    `ids` come back merged with the labels (and cut to the labels' length), logp / hidden stay the
    model's own over the full decoded length."""
    sd = _strip(dec_sd)
    B = features.size(0)
    orig_labels = labels
    mem = memory_bank(sd, features)
    state = DecoderState()
    w_out, b_out = sd[_P + "output_layer.weight"], sd[_P + "output_layer.bias"]
    dev = features.device      # the restatement runs wherever its inputs live (bench.py's eager-GPU baseline)
    alive_seq = torch.full((B, 1), SOS_ID, dtype=torch.long, device=dev)
    alive_logp = torch.zeros((B, 0), device=dev)
    alive_hidden = None
    alive_lp_rec = [] if record_logprobs else None
    orig_idx = torch.arange(B, device=dev)
    results: List[Optional[dict]] = [None] * B
    with torch.no_grad():
        for step in range(max_len):
            tgt = alive_seq[:, -1]
            if labels is not None:
                label = labels[:, step]                   # IndexError past the last column, as in the reference
                mask = label.eq(MASK_ID).long()
                tgt = tgt * mask + label * (1 - mask)
            dec_out = decoder_step(sd, tgt, mem, state)
            logits = F.linear(dec_out, w_out, b_out)
            log_probs = F.log_softmax(logits, dim=-1)
            log_probs.masked_fill_(grammar_mask(tgt, *grammar), -10000)
            if alive_seq.shape[1] <= min_length:          # ensure_min_length
                log_probs[:, EOS_ID] = -1e20
            topk_scores, topk_ids = log_probs.topk(1, dim=-1)
            if forced_ids is not None:
                topk_ids = forced_ids[orig_idx, step].view(-1, 1)
                topk_scores = log_probs.gather(1, topk_ids)
            is_finished = topk_ids.eq(EOS_ID).view(-1)
            if labels is not None and step + 1 < labels.size(1):
                is_finished = labels[:, step + 1].eq(EOS_ID)
            alive_seq = torch.cat([alive_seq, topk_ids], -1)
            alive_logp = torch.cat([alive_logp, topk_scores], -1)
            h = dec_out.unsqueeze(1)
            alive_hidden = h if alive_hidden is None else torch.cat([alive_hidden, h], 1)
            if record_logprobs:
                alive_lp_rec.append(log_probs.clone())
            if alive_seq.shape[1] == max_len + 1:         # ensure_max_length
                is_finished = torch.ones_like(is_finished)
            if is_finished.any():
                # update_finished, greedy_search.py:100-127
                for b in is_finished.nonzero().view(-1).tolist():
                    r = dict(ids=alive_seq[b, 1:].clone(), logp=alive_logp[b].clone(),
                             hidden=alive_hidden[b].clone(),
                             score=float(torch.exp(torch.mean(alive_logp[b]))))
                    if record_logprobs:
                        r["logprobs"] = torch.stack([lp[b] for lp in alive_lp_rec], 0)
                    results[int(orig_idx[b])] = r
                if bool(is_finished.all()):
                    break
                alive = ~is_finished
                alive_seq, alive_logp, alive_hidden = alive_seq[alive], alive_logp[alive], alive_hidden[alive]
                if record_logprobs:
                    alive_lp_rec = [lp[alive] for lp in alive_lp_rec]
                select = alive.nonzero().view(-1)
                orig_idx = orig_idx[alive]
                mem = mem.index_select(0, select)
                if labels is not None:
                    labels = labels.index_select(0, select)
                state.index_select(select)
    if orig_labels is not None:
        for i in range(B):
            pred = results[i]["ids"]
            label = orig_labels[i][1:len(pred) + 1]
            mask = label.eq(MASK_ID).long()
            pred = pred[:len(label)]
            results[i]["ids"] = pred * mask + label * (1 - mask)
    return results


def beam_decode(dec_sd: SD, features: torch.Tensor, beam_size: int, n_best: int = 1, max_len: int = 480,
                grammar=(101, 64, 64), min_length: int = 1, min_gap: Optional[list] = None,
                follow: Optional[np.ndarray] = None, follow_stats: Optional[dict] = None,
                record: Optional[np.ndarray] = None):
    """TransformerDecoderAR.decode with BeamSearch -- PARITY UNPINNED.

    The reference's beam branch cannot run (constructor arguments are passed in
    the wrong order so max_length becomes False, `advance` is called with four arguments but
    takes two, caches are only reordered when a hypothesis finished, hidden states are not
    tracked).  This function is the *intended* algorithm of decoding/beam_search.py:84-190 with
    exactly those four defects repaired, following the OpenNMT-py translator the file was
    adapted from:
      * length-normalised score `(cum + log_prob) / (len(self) + 1)` (beam_search.py:96-101),
        cumulative log-prob recovered as `score * curr_length` (:105);
      * top-`beam` over the flattened (beam * V) scores (:80-82); ties are broken by the LOWEST
        flat index (torch.topk leaves tie order unspecified; this restatement pins it with a
        stable sort so the CUDA path has something definite to match);
      * caches, memory bank, per-hypothesis hidden states and token log-probs are reordered by
        `select_indices` EVERY step (the OpenNMT behaviour; the reference does it only
        `if any_finished`, components.py:314-319, which would corrupt the KV cache);
      * finished hypotheses get cumulative log-prob -1e10 (:136), are stored with their
        length-normalised score (:149-153); an image ends when its top beam has finished and
        it holds >= n_best hypotheses (:156-167), the best n_best by score (stable, first
        stored wins ties) are returned;
      * grammar mask keyed on each row's input token and `ensure_min_length` as in greedy
        (components.py:298-301, decode_strategy.py:51-53);
      * row r of the alive (image-major, beam-minor) batch receives pe[r].

    Returns a list (one per image) of lists (n_best) of dicts: ids (L,) int64 incl. <eos>,
    logp (L,) masked log-prob of each chosen token, hidden (L,256), score (float, the
    length-normalised log score BeamSearch stores -- not exponentiated).
    `min_gap` (a list, one float per image, filled in place) receives the smallest difference, in
    cumulative-log-prob units, between adjacent entries of the top beam+1 candidates over all
    steps: how close this image came to a different selection (near-tie diagnosis in tests).
    `follow` (T, B, >=beam) int: flat indices (beam * V + token) another implementation selected
    at each step for each ORIGINAL image.  The oracle then adopts those selections instead of its
    own after measuring how far they are from a valid descending top-`beam` of ITS scores
    (`follow_stats['max_violation']`, in cumulative-log-prob units; 0 when they are exactly a valid
    top-k) and counting the steps where they differ from its own choice
    (`follow_stats['deviations']`).  With near-ties as frequent as they are under beam search
    (candidates a few 1e-5 apart, fp32 summation order decides) this is the sound way to check a
    second implementation end to end: every selection must be optimal within tolerance on the
    oracle's own arithmetic, and everything derived from the selections must then match.
    `record` (T, B, beam) int array: receives the selections actually applied, same layout."""
    sd = _strip(dec_sd)
    B, K = features.size(0), beam_size
    V = sum(grammar)
    mem = torch.repeat_interleave(memory_bank(sd, features), K, dim=0)       # beam_search.py:35
    state = DecoderState()
    w_out, b_out = sd[_P + "output_layer.weight"], sd[_P + "output_layer.bias"]
    alive_seq = torch.full((B * K, 1), SOS_ID, dtype=torch.long)
    alive_logp = torch.zeros((B * K, 0))
    alive_hidden = torch.zeros((B * K, 0, DEC_DIM))
    topk_log_probs = torch.tensor([0.0] + [float("-inf")] * (K - 1)).repeat(B)   # beam_search.py:45-47
    batch_offset = torch.arange(B)
    top_beam_finished = torch.zeros(B, dtype=torch.bool)
    hypotheses: List[list] = [[] for _ in range(B)]
    results: List[list] = [[] for _ in range(B)]
    with torch.no_grad():
        for step in range(max_len):
            _B = alive_seq.size(0) // K
            tgt = alive_seq[:, -1]
            dec_out = decoder_step(sd, tgt, mem, state)
            log_probs = F.log_softmax(F.linear(dec_out, w_out, b_out), dim=-1)
            log_probs.masked_fill_(grammar_mask(tgt, *grammar), -10000)
            cur_len = alive_seq.shape[1]                       # len(self)
            if cur_len <= min_length:
                log_probs[:, EOS_ID] = -1e20
            token_lp = log_probs.clone()
            log_probs = log_probs + topk_log_probs.view(_B * K, 1)
            curr_length = cur_len + 1
            curr_scores = (log_probs / curr_length).reshape(_B, K * V)
            order = torch.sort(curr_scores, dim=-1, descending=True, stable=True)
            topk_scores, topk_flat = order.values[:, :K], order.indices[:, :K]
            if min_gap is not None:
                top = order.values[:, :K + 1]
                gaps = (top[:, :-1] - top[:, 1:]) * curr_length
                gaps = torch.where(torch.isfinite(gaps), gaps, torch.full_like(gaps, float("inf"))).min(dim=1).values
                for i in range(_B):
                    b = int(batch_offset[i])
                    min_gap[b] = min(min_gap[b], float(gaps[i]))
            if follow is not None:
                theirs = torch.as_tensor(np.asarray(follow[step])[batch_offset.numpy(), :K].astype(np.int64))
                their_scores = curr_scores.gather(1, theirs)
                if not torch.equal(theirs, topk_flat):
                    follow_stats["deviations"] = follow_stats.get("deviations", 0) + 1
                    fin = torch.isfinite(their_scores)
                    # (a) listed in descending order, (b) nothing left out beats the last one listed
                    desc = torch.where(fin[:, :-1] & fin[:, 1:], their_scores[:, 1:] - their_scores[:, :-1],
                                       torch.zeros_like(their_scores[:, 1:])).clamp(min=0).max() if K > 1 else torch.tensor(0.0)
                    rest = curr_scores.clone()
                    rest.scatter_(1, theirs, float("-inf"))
                    left = torch.where(fin[:, -1], rest.max(dim=1).values - their_scores[:, -1],
                                       torch.zeros_like(their_scores[:, -1])).clamp(min=0).max()
                    viol = float(max(desc, left)) * curr_length
                    if len(set(theirs.view(-1).tolist())) and any(len(set(r)) < K for r in theirs.tolist()):
                        viol = float("inf")            # a candidate selected twice
                    follow_stats["max_violation"] = max(follow_stats.get("max_violation", 0.0), viol)
                topk_flat, topk_scores = theirs, their_scores
            if record is not None:
                record[step, batch_offset.numpy(), :K] = topk_flat.numpy()
            topk_log_probs = (topk_scores * curr_length).reshape(-1)
            parent = topk_flat // V + (torch.arange(_B) * K).unsqueeze(1)
            select = parent.view(-1)
            topk_ids = topk_flat % V
            new_tok = topk_ids.view(-1, 1)
            alive_logp = torch.cat([alive_logp.index_select(0, select), token_lp[select].gather(1, new_tok)], -1)
            alive_hidden = torch.cat([alive_hidden.index_select(0, select), dec_out[select].unsqueeze(1)], 1)
            alive_seq = torch.cat([alive_seq.index_select(0, select), new_tok], -1)
            is_finished = topk_ids.eq(EOS_ID)
            if alive_seq.shape[1] == max_len + 1:
                is_finished = torch.ones_like(is_finished)
            # reorder decoder state every step (OpenNMT translator; see docstring)
            mem = mem.index_select(0, select)
            state.index_select(select)
            if not bool(is_finished.any()):
                continue
            # update_finished, beam_search.py:133-190
            topk_log_probs = topk_log_probs.view(_B, K).masked_fill(is_finished, -1e10).view(-1)
            top_beam_finished |= is_finished[:, 0]
            keep = []
            for i in range(_B):
                b = int(batch_offset[i])
                for j in is_finished[i].nonzero().view(-1).tolist():
                    r = i * K + j
                    hypotheses[b].append(dict(score=float(topk_scores[i, j]), ids=alive_seq[r, 1:].clone(),
                                              logp=alive_logp[r].clone(), hidden=alive_hidden[r].clone()))
                if bool(top_beam_finished[i]) and len(hypotheses[b]) >= n_best:
                    best = sorted(hypotheses[b], key=lambda h: h["score"], reverse=True)
                    results[b] = best[:n_best]
                else:
                    keep.append(i)
            if not keep:
                break
            if len(keep) < _B:
                kept = torch.tensor(keep)
                rows = (kept.unsqueeze(1) * K + torch.arange(K)).view(-1)
                top_beam_finished = top_beam_finished.index_select(0, kept)
                batch_offset = batch_offset.index_select(0, kept)
                topk_log_probs = topk_log_probs.index_select(0, rows)
                alive_seq, alive_logp, alive_hidden = alive_seq[rows], alive_logp[rows], alive_hidden[rows]
                mem = mem.index_select(0, rows)
                state.index_select(rows)
    return results


# --------------------------------------------------------------------------------------
# Bond head  (components.py:350-400, driver :470-491)
# --------------------------------------------------------------------------------------
def edge_probabilities(dec_sd: SD, hidden: torch.Tensor, indices: Sequence[int]) -> torch.Tensor:
    """GraphPredictor.forward on one image + softmax: (k,k,7) fp32 (components.py:365-380, :481)."""
    sd = _strip(dec_sd)
    idx = torch.as_tensor(list(indices), dtype=torch.long)
    h = hidden[idx]
    k = h.size(0)
    hh = torch.cat([h.unsqueeze(1).expand(k, k, DEC_DIM), h.unsqueeze(0).expand(k, k, DEC_DIM)], dim=2)
    y = F.linear(hh, sd["decoder.edges.mlp.0.weight"], sd["decoder.edges.mlp.0.bias"])
    y = F.linear(F.gelu(y), sd["decoder.edges.mlp.2.weight"], sd["decoder.edges.mlp.2.bias"])
    return F.softmax(y, dim=2)


def symmetrised_edge_probabilities(prob: np.ndarray) -> np.ndarray:
    """The in-place symmetrisation of get_edge_prediction (components.py:389-397) in double precision, as the
    reference runs it on Python floats obtained from `.tolist()`.  The 5/6 rule reads edge_prob[i][j][5]
    AFTER it was overwritten, exactly as the reference does."""
    n = prob.shape[0]
    p = prob.astype(np.float64).copy()
    for i in range(n):
        for j in range(i + 1, n):
            for k in range(5):
                p[i, j, k] = (p[i, j, k] + p[j, i, k]) / 2
                p[j, i, k] = p[i, j, k]
            p[i, j, 5] = (p[i, j, 5] + p[j, i, 6]) / 2
            p[i, j, 6] = (p[i, j, 6] + p[j, i, 5]) / 2
            p[j, i, 5] = p[i, j, 6]
            p[j, i, 6] = p[i, j, 5]
    return p


def get_edge_prediction(prob: np.ndarray):
    """get_edge_prediction (components.py:383-400): symmetrise, then argmax / max per atom pair."""
    if prob.shape[0] == 0:
        return [], []
    p = symmetrised_edge_probabilities(prob)
    return np.argmax(p, axis=2).tolist(), np.max(p, axis=2).tolist()


def decode(dec_sd: SD, features: torch.Tensor, tokenizer, max_len: int = 480, return_raw: bool = False):
    """Decoder.decode (components.py:443-492) for formats [chartok_coords, edges],
    compute_confidence off: list of {'chartok_coords': {...}, 'edges': [[int]]}."""
    raw = greedy_decode(dec_sd, features, max_len=max_len,
                        grammar=(tokenizer.offset, tokenizer.maxx, tokenizer.maxy))
    preds = []
    for r in raw:
        ct = tokenizer.sequence_to_smiles(r["ids"].tolist())
        if len(ct["indices"]) > 0:
            prob = edge_probabilities(dec_sd, r["hidden"], ct["indices"]).numpy()
            edges, _ = get_edge_prediction(prob)
        else:
            edges = []
        preds.append({"chartok_coords": ct, "edges": edges})
    return (preds, raw) if return_raw else preds


def predict(ckpt: dict, images: torch.Tensor, tokenizer, max_len: int = 480):
    """encoder + decoder.decode as called from molnextr.predict_images (model.py:105-108)."""
    with torch.no_grad():
        feats = encoder_features(ckpt["encoder"], images)
        return decode(ckpt["decoder"], feats, tokenizer, max_len=max_len)
