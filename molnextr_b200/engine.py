"""Python handle over the C ABI.  torch tensors are containers only: every call passes
`data_ptr()`, shapes and the current CUDA stream to libmolnextr_b200.so."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np
import torch

from . import _cabi
from .tokenization import CharTokenizer

MAX_LEN = 480          # FORMAT_INFO['chartok_coords']['max_len'], MolNexTR/utils.py:25
MAX_ATOMS = 160        # an atom needs >= 3 tokens


def token_class_table(tok: CharTokenizer) -> np.ndarray:
    """Per-id class bits consumed by the device-side atom scan (see include/molnextr_b200.h)."""
    tab = np.zeros(len(tok), np.uint8)
    for i in range(tok.offset):
        bits = 0
        if tok.is_symbol(i):
            bits |= 1
            ch = tok.itos[i]
            if tok.is_atom_token(ch):
                bits |= 2
            for bit, c in ((4, "["), (8, "]"), (16, "C"), (32, "l"), (64, "B"), (128, "r")):
                if ch == c:
                    bits |= bit
        tab[i] = bits
    return tab


def encoder_kind_of(enc_state: Optional[Dict[str, torch.Tensor]]) -> int:
    if not enc_state:
        return _cabi.ENCODER_NONE
    key = next(iter(enc_state)).replace("module.", "")
    if key.startswith("transformer."):
        return _cabi.ENCODER_SWIN_B
    if key.startswith("cnn."):
        return _cabi.ENCODER_CONVNEXT_B
    raise ValueError(f"unrecognised encoder state-dict (first key {key!r})")


def pipeline_plan(n_batches: int, batch: int, seq_len: int, max_clusters: int, num_sms: int, depth: int = 0, encoder_ctas: int = 0) -> dict:
    """Schedule of `Engine.predict_pipelined` (pure host logic, CPU-testable).

    A batch decodes on `clusters_per_batch` 8-CTA clusters of 16 rows; `depth` batches decode side by side while the encoders of
    the following ones run on the remaining SMs (`encoder_ctas` = cap of their persistent GEMM grids).  The throughput kernel
    needs every cluster of a batch co-resident (`max_clusters`) and at most 512 memory positions; otherwise the batches run one
    at a time on the automatic path (depth 1).  The last wave may be partial: its batches then use `16 // spread` rows per cluster,
    i.e. `spread` times the SMs, so the tail of a run does not idle the GPU."""
    n_clusters = (batch + 15) // 16
    wide_ok = 0 < n_clusters <= max_clusters and seq_len <= 512
    if not wide_ok:
        depth = 1
    elif depth <= 0:
        # decode and encoder share the SMs: ~3/4 of them to `depth` decode kernels balances the two at bs = 32
        # (measured on B200 with the 6.6 ms encoder, 20 batches: depth 5 / 6 / 7 / 8 -> 934 / 1000 / 1104 / 930 img/s;
        #  with round 2's first 10.6 ms encoder the optimum was 6)
        depth = max(1, min(max_clusters // n_clusters, int(0.76 * num_sms) // (8 * n_clusters)))
    depth = max(1, min(depth, n_batches))
    if encoder_ctas <= 0:
        encoder_ctas = max(16, num_sms - depth * 8 * n_clusters) if wide_ok else 32
    last_wave, tail = (n_batches - 1) // depth, n_batches - depth * ((n_batches - 1) // depth)
    spread = 1
    while wide_ok and spread * 2 <= depth // tail and spread < 4:
        spread *= 2
    return {"depth": depth, "throughput_kernel": bool(wide_ok), "encoder_ctas": encoder_ctas, "clusters_per_batch": n_clusters,
            "last_wave": last_wave, "tail": tail, "spread": spread}


class EngineError(RuntimeError):
    pass


class Engine:
    """One engine per GPU.  `checkpoint` is a dict in the reference schema
    ({'encoder': sd, 'decoder': sd, 'args': {...}}, main.py:389-398); pass encoder=None for a
    decoder-only handle."""

    def __init__(self, checkpoint: dict, tokenizer: Optional[CharTokenizer] = None, device: int = 0,
                 max_batch: int = 32, max_height: int = 384, max_width: int = 384, encoder_dim: int = 1024,
                 max_beam: int = 1):
        if not torch.cuda.is_available():
            raise EngineError("no CUDA device: molnextr_b200 has no CPU fallback")
        self.lib = _cabi.load()
        self.tok = tokenizer or CharTokenizer(64)
        self.device = torch.device("cuda", device)
        self.max_batch, self.max_height, self.max_width = max_batch, max_height, max_width
        self.encoder_dim = encoder_dim
        self.max_beam = max_beam
        self._host_out: Dict[int, Dict[str, torch.Tensor]] = {}
        self._pipe_streams = None
        self.last_pipeline: Dict[str, int] = {}
        self._host_pool: Dict[tuple, Dict[str, torch.Tensor]] = {}
        self._decode_path = self.DECODE_PATHS.get(os.environ.get("MNX_DECODE_PATH", "auto"), 0)   # mirrors mnx_create
        enc_sd = checkpoint.get("encoder")
        self.encoder_kind = encoder_kind_of(enc_sd)
        offset, maxx, maxy = self.tok.grammar_rule()
        cls = token_class_table(self.tok)
        self._cls = cls   # keep alive during create
        cfg = _cabi.MnxConfig(device=device, encoder_kind=self.encoder_kind, max_batch=max_batch,
                              max_height=max_height, max_width=max_width, max_len=MAX_LEN, vocab=len(self.tok),
                              tok_offset=offset, max_x=maxx, max_y=maxy, max_atoms=MAX_ATOMS,
                              encoder_dim=encoder_dim, max_beam=max_beam,
                              token_class=cls.ctypes.data_as(C.POINTER(C.c_uint8)))
        h = C.c_void_p()
        rc = self.lib.mnx_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise EngineError(f"mnx_create failed ({rc}): {self.lib.mnx_last_error(None).decode()}")
        self.h = h
        try:
            if enc_sd:
                self._load("encoder.", enc_sd)
            self._load("decoder.", checkpoint["decoder"])
            self._check(self.lib.mnx_finalize_weights(self.h), "mnx_finalize_weights")
        except Exception:
            self.close()
            raise

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int, what: str):
        if rc != 0:
            raise EngineError(f"{what} failed ({rc}): {self.lib.mnx_last_error(self.h).decode()}")

    def _load(self, scope: str, sd: Dict[str, torch.Tensor]):
        for k, v in sd.items():
            t = v.detach().cpu().contiguous()
            is_i64 = t.dtype == torch.int64
            if not is_i64:
                t = t.float().contiguous()
            shape = (C.c_int64 * t.dim())(*t.shape)
            self._check(self.lib.mnx_load_tensor(self.h, (scope + k).encode(), C.c_void_p(t.data_ptr()), shape,
                                                 t.dim(), int(is_i64)), f"mnx_load_tensor({k})")

    def close(self):
        if getattr(self, "h", None):
            self.lib.mnx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _same_device(self, *tensors):
        for t in tensors:
            if t is not None and t.is_cuda and t.device != self.device:
                raise EngineError(f"tensor on {t.device} passed to an engine that lives on {self.device}")

    @staticmethod
    def _p(t: Optional[torch.Tensor]) -> C.c_void_p:
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

    # ------------------------------------------------------------------ upstream: preprocessing on the device
    def preprocess(self, images, size: int = 384, pad: int = 50) -> torch.Tensor:
        """The reference's inference transform (dataset.py:158-185 + data_aug.CropWhite) for a list of RGB
        uint8 HxWx3 arrays of any size -> cuda fp32 (B, 3, size, size), bit-exact with the cv2 path in
        molnextr_b200/preprocess.py.  One pinned packed upload of the raw bytes, three kernels."""
        from .preprocess import MEAN, STD
        B = len(images)
        hs = np.array([im.shape[0] for im in images], np.int32)
        ws = np.array([im.shape[1] for im in images], np.int32)
        sizes = hs.astype(np.int64) * ws.astype(np.int64) * 3
        offsets = np.zeros(B, np.int64)
        offsets[1:] = np.cumsum((sizes[:-1] + 15) // 16 * 16)
        total = int(offsets[-1] + sizes[-1])
        packed = torch.empty((total,), dtype=torch.uint8, pin_memory=True)
        pk = packed.numpy()
        for im, o, n in zip(images, offsets, sizes):
            assert im.dtype == np.uint8 and im.ndim == 3 and im.shape[2] == 3
            pk[o:o + n] = np.ascontiguousarray(im).reshape(-1)
        dev = packed.to(self.device, non_blocking=True)
        mean255 = (MEAN * 255.0).astype(np.float32)
        inv = (1.0 / (STD * 255.0)).astype(np.float32)
        out = torch.empty((B, 3, size, size), device=self.device, dtype=torch.float32)
        self._check(self.lib.mnx_preprocess(self.h, self._p(dev), C.c_void_p(offsets.ctypes.data), C.c_void_p(hs.ctypes.data),
                                            C.c_void_p(ws.ctypes.data), B, pad, size, C.c_void_p(mean255.ctypes.data),
                                            C.c_void_p(inv.ctypes.data), self._p(out), self._stream()), "mnx_preprocess")
        return out

    # ------------------------------------------------------------------ hot path
    def seq_len(self, H: int, W: int) -> int:
        if self.encoder_kind == _cabi.ENCODER_SWIN_B:
            h, w = (H + 3) // 4, (W + 3) // 4
            for _ in range(3):
                h, w = (h + 1) // 2, (w + 1) // 2
            return h * w
        return (H // 32) * (W // 32)

    def encode(self, images: torch.Tensor) -> torch.Tensor:
        """Encoder.forward (components.py:162-174): fp32 NCHW cuda tensor -> (B,S,1024)."""
        assert images.is_cuda and images.dtype == torch.float32 and images.dim() == 4 and images.size(1) == 3
        self._same_device(images)
        images = images.contiguous()
        B, _, H, W = images.shape
        feats = torch.empty((B, self.seq_len(H, W), self.encoder_dim), device=images.device, dtype=torch.float32)
        self._check(self.lib.mnx_encode(self.h, self._p(images), B, H, W, self._p(feats), self._stream()), "mnx_encode")
        return feats

    def decode_greedy(self, features: torch.Tensor, return_hidden: bool = False, labels: Optional[torch.Tensor] = None):
        """TransformerDecoderAR.decode with GreedySearch (components.py:253-334).  `labels` (B, n) integer tensor =
        partial prediction (components.py:286-289,326-332): column 0 is <sos>, MASK_ID (4) marks the positions the
        model fills in, every other position is given; ids / lens come back merged with the labels."""
        assert features.is_cuda and features.dtype == torch.float32
        self._same_device(features)
        features = features.contiguous().view(features.size(0), -1, features.size(-1))
        B, S, _ = features.shape
        dev = features.device
        ids = torch.empty((B, MAX_LEN), device=dev, dtype=torch.int32)
        lens = torch.empty((B,), device=dev, dtype=torch.int32)
        logp = torch.empty((B, MAX_LEN), device=dev, dtype=torch.float32)
        hidden = torch.empty((B, MAX_LEN, 256), device=dev, dtype=torch.float32) if return_hidden else None
        if labels is not None:
            assert labels.dim() == 2 and labels.size(0) == B, "labels must be (B, n)"
            lab = labels.to(device=dev, dtype=torch.int32).contiguous()
            self._check(self.lib.mnx_decode_greedy_labels(self.h, self._p(features), B, S, self._p(lab), lab.size(1), self._p(ids),
                                                          self._p(lens), self._p(logp), self._p(hidden), self._stream()),
                        "mnx_decode_greedy_labels")
        else:
            self._check(self.lib.mnx_decode_greedy(self.h, self._p(features), B, S, self._p(ids), self._p(lens),
                                                   self._p(logp), self._p(hidden), self._stream()), "mnx_decode_greedy")
        out = {"ids": ids, "lens": lens, "logp": logp}
        if return_hidden:
            out["hidden"] = hidden
        return out

    def decode_beam(self, features: torch.Tensor, beam_size: int, n_best: int = 1, return_hidden: bool = False):
        """TransformerDecoderAR.decode with BeamSearch (components.py:253-334, decoding/beam_search.py),
        with the repairs listed in oracle/restate.py beam_decode -- the reference branch itself
        cannot run.  ids/logp are (B, n_best, 480), lens/scores (B, n_best), best
        hypothesis first; `hidden` (B, 480, 256) belongs to the best hypothesis."""
        assert features.is_cuda and features.dtype == torch.float32
        features = features.contiguous().view(features.size(0), -1, features.size(-1))
        B, S, _ = features.shape
        dev = features.device
        ids = torch.empty((B, n_best, MAX_LEN), device=dev, dtype=torch.int32)
        lens = torch.empty((B, n_best), device=dev, dtype=torch.int32)
        scores = torch.empty((B, n_best), device=dev, dtype=torch.float32)
        logp = torch.empty((B, n_best, MAX_LEN), device=dev, dtype=torch.float32)
        hidden = torch.empty((B, MAX_LEN, 256), device=dev, dtype=torch.float32) if return_hidden else None
        self._check(self.lib.mnx_decode_beam(self.h, self._p(features), B, S, beam_size, n_best, self._p(ids),
                                             self._p(lens), self._p(scores), self._p(logp), self._p(hidden),
                                             self._stream()), "mnx_decode_beam")
        out = {"ids": ids, "lens": lens, "scores": scores, "logp": logp}
        if return_hidden:
            out["hidden"] = hidden
        return out

    def beam_trace(self, batch: int) -> np.ndarray:
        """(480, batch, 8) int32 selections of the last decode_beam (parity tests; see the header)."""
        tr = np.empty((MAX_LEN, batch, 8), np.int32)
        self._check(self.lib.mnx_beam_trace(self.h, C.c_void_p(tr.ctypes.data), batch), "mnx_beam_trace")
        return tr

    def atom_indices(self, ids: Optional[torch.Tensor] = None, lens: Optional[torch.Tensor] = None, batch: int = 0):
        """ids/lens None: scan the engine-internal result of the last decode (`batch` rows)."""
        B = ids.size(0) if ids is not None else batch
        atom_idx = torch.full((B, MAX_ATOMS), -1, device=self.device, dtype=torch.int32)
        n_atoms = torch.empty((B,), device=self.device, dtype=torch.int32)
        self._check(self.lib.mnx_atom_indices(self.h, self._p(ids), self._p(lens), B, self._p(atom_idx),
                                              self._p(n_atoms), self._stream()), "mnx_atom_indices")
        return atom_idx, n_atoms

    def edges(self, atom_idx: torch.Tensor, n_atoms: torch.Tensor, hidden: Optional[torch.Tensor] = None,
              return_scores: bool = False):
        """GraphPredictor + get_edge_prediction (components.py:365-400) for every image of the batch."""
        self._same_device(atom_idx, n_atoms, hidden)
        B = atom_idx.size(0)
        edges = torch.zeros((B, MAX_ATOMS, MAX_ATOMS), device=atom_idx.device, dtype=torch.uint8)
        score = torch.zeros((B, MAX_ATOMS, MAX_ATOMS), device=atom_idx.device, dtype=torch.float32) if return_scores else None
        self._check(self.lib.mnx_edges(self.h, self._p(hidden), self._p(atom_idx), self._p(n_atoms), B,
                                       self._p(edges), self._p(score), self._stream()), "mnx_edges")
        return (edges, score) if return_scores else edges

    def confidence(self, ids: torch.Tensor, lens: torch.Tensor, logp: torch.Tensor, edge_score: torch.Tensor):
        """Decoder.decode's compute_confidence outputs (components.py:456-469,485-491) on the device: per-atom geometric-mean
        token score (B, max_atoms) fp32, average token score (B,) fp32, overall score (B,) fp64."""
        self._same_device(ids, lens, logp, edge_score)
        B = ids.size(0)
        atom_scores = torch.zeros((B, MAX_ATOMS), device=ids.device, dtype=torch.float32)
        seq_score = torch.empty((B,), device=ids.device, dtype=torch.float32)
        overall = torch.empty((B,), device=ids.device, dtype=torch.float64)
        self._check(self.lib.mnx_confidence(self.h, self._p(ids.contiguous()), self._p(lens.contiguous()), self._p(logp.contiguous()), B,
                                            self._p(edge_score.contiguous()), self._p(atom_scores), self._p(seq_score), self._p(overall),
                                            self._stream()), "mnx_confidence")
        return atom_scores, seq_score, overall

    def predict(self, images: torch.Tensor):
        """encoder -> greedy decode -> atom scan -> bond head, device tensors in and out."""
        assert images.is_cuda and images.dtype == torch.float32
        self._same_device(images)
        images = images.contiguous()
        B, _, H, W = images.shape
        dev = images.device
        ids = torch.empty((B, MAX_LEN), device=dev, dtype=torch.int32)
        lens = torch.empty((B,), device=dev, dtype=torch.int32)
        logp = torch.empty((B, MAX_LEN), device=dev, dtype=torch.float32)
        atom_idx = torch.empty((B, MAX_ATOMS), device=dev, dtype=torch.int32)
        n_atoms = torch.empty((B,), device=dev, dtype=torch.int32)
        edges = torch.empty((B, MAX_ATOMS, MAX_ATOMS), device=dev, dtype=torch.uint8)
        self._check(self.lib.mnx_predict(self.h, self._p(images), B, H, W, self._p(ids), self._p(lens), self._p(logp),
                                         self._p(atom_idx), self._p(n_atoms), self._p(edges), self._stream()),
                    "mnx_predict")
        return {"ids": ids, "lens": lens, "logp": logp, "atom_idx": atom_idx, "n_atoms": n_atoms, "edges": edges}

    def empty_result(self):
        """Zero-row result with the dtypes / trailing shapes of `predict` (a rank with an empty shard, parallel.py)."""
        dev = self.device
        return {"ids": torch.empty((0, MAX_LEN), device=dev, dtype=torch.int32),
                "lens": torch.empty((0,), device=dev, dtype=torch.int32),
                "logp": torch.empty((0, MAX_LEN), device=dev, dtype=torch.float32),
                "atom_idx": torch.empty((0, MAX_ATOMS), device=dev, dtype=torch.int32),
                "n_atoms": torch.empty((0,), device=dev, dtype=torch.int32),
                "edges": torch.empty((0, MAX_ATOMS, MAX_ATOMS), device=dev, dtype=torch.uint8)}

    def predict_host(self, images: torch.Tensor):
        """Same through host buffers: H2D of the images and D2H of every result inside the call.
        The result tensors are pinned staging buffers owned by the engine (allocating pinned memory
        costs tens of milliseconds): they are overwritten by the next predict_host call with the
        same batch size -- copy what must outlive it."""
        assert not images.is_cuda and images.dtype == torch.float32
        images = images.contiguous()
        B, _, H, W = images.shape
        out = self._host_out.get(B)
        if out is None:
            pin = dict(pin_memory=True)
            out = {"ids": torch.empty((B, MAX_LEN), dtype=torch.int32, **pin),
                   "lens": torch.empty((B,), dtype=torch.int32, **pin),
                   "logp": torch.empty((B, MAX_LEN), dtype=torch.float32, **pin),
                   "atom_idx": torch.empty((B, MAX_ATOMS), dtype=torch.int32, **pin),
                   "n_atoms": torch.empty((B,), dtype=torch.int32, **pin),
                   "edges": torch.empty((B, MAX_ATOMS, MAX_ATOMS), dtype=torch.uint8, **pin)}
            self._host_out[B] = out
        self._check(self.lib.mnx_predict_host(self.h, self._p(images), B, H, W, self._p(out["ids"]), self._p(out["lens"]),
                                              self._p(out["logp"]), self._p(out["atom_idx"]), self._p(out["n_atoms"]),
                                              self._p(out["edges"])),
                    "mnx_predict_host")
        return dict(out)

    # ------------------------------------------------------------------ pipelining across batches
    def _decode_all(self, feats: torch.Tensor):
        out = self.decode_greedy(feats)
        out["atom_idx"], out["n_atoms"] = self.atom_indices(out["ids"], out["lens"])
        out["edges"] = self.edges(out["atom_idx"], out["n_atoms"])
        return out

    def predict_pipelined(self, batches, host: bool = False, depth: int = 0, encoder_ctas: int = 0):
        """Greedy predictions for consecutive image batches, several batches in flight (the reference handles its
        mini-batches one after the other: model.py:102-109, main.py:273-293).

        Each batch decodes with the throughput kernel (csrc/wide.cu: one persistent launch, 8-CTA clusters of
        <= 16 rows, i.e. 16 SMs for a batch of 32) in its own decode context on its own high-priority stream;
        `depth` batches decode side by side while a lower-priority stream runs the encoders of the following
        batches with their persistent GEMM grids capped at `encoder_ctas` CTAs (default: the SMs the decode
        kernels leave free).  Results are identical to calling `predict` batch by batch -- every batch keeps its
        own row ranks.  `batches`: iterable of fp32 (B,3,H,W) tensors -- cuda tensors, or with
        host=True pinned host tensors (H2D inside; results are returned as pinned host tensors, D2H inside; those
        buffers are owned by the engine and REUSED by the next predict_pipelined(host=True) call with the same
        batch sizes: copy what must outlive it).
        Returns a list of result dicts (ids, lens, logp, atom_idx, n_atoms, edges)."""
        batches = list(batches)
        if not batches:
            return []
        B0 = max(int(x.shape[0]) for x in batches)
        H0, W0 = int(batches[0].shape[2]), int(batches[0].shape[3])
        max_cl = int(self.time_kernel(1004, 1))
        num_sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        plan = pipeline_plan(len(batches), B0, self.seq_len(H0, W0), max_cl, num_sms, depth, encoder_ctas)
        wide_ok, depth, encoder_ctas, n_clusters = plan["throughput_kernel"], plan["depth"], plan["encoder_ctas"], plan["clusters_per_batch"]
        cur = torch.cuda.current_stream(self.device)
        if self._pipe_streams is None or len(self._pipe_streams[1]) < depth:
            # torch: priority -1 = high, 0 = low
            old = self._pipe_streams[1] if self._pipe_streams else []
            self._pipe_streams = (self._pipe_streams[0] if self._pipe_streams else torch.cuda.Stream(self.device, priority=0),
                                  old + [torch.cuda.Stream(self.device, priority=-1) for _ in range(depth - len(old))])
        self.last_pipeline = dict(plan)
        enc, decs = self._pipe_streams[0], self._pipe_streams[1][:depth]
        enc.wait_stream(cur)
        for d in decs:
            d.wait_stream(cur)
        outs, keep = [], []
        prev_path = self._decode_path
        if wide_ok:
            self.reserve_contexts(depth)
            self.set_decode_path("wide")
        last_wave, spread = plan["last_wave"], plan["spread"]
        try:
            for i, x in enumerate(batches):
                dec = decs[i % depth]
                if wide_ok:
                    self._check(self.lib.mnx_set_wide_rows(self.h, 16 // spread if i // depth == last_wave else 0), "mnx_set_wide_rows")
                # the first encoders find the GPU empty: every SM; later ones share it with `depth` decode kernels
                self._check(self.lib.mnx_set_encoder_cta_limit(self.h, 0 if i == 0 else encoder_ctas), "mnx_set_encoder_cta_limit")
                with torch.cuda.stream(enc):
                    xd = x if x.is_cuda else x.to(self.device, non_blocking=True)
                    f = self.encode(xd)
                    f.record_stream(dec)
                    ev = torch.cuda.Event()
                    ev.record(enc)
                keep.append(xd)
                with torch.cuda.stream(dec):
                    dec.wait_event(ev)
                    if wide_ok:
                        self.set_context(i % depth)
                    out = self._decode_all(f)           # asynchronous: one persistent kernel + scan + bond head
                    if host:
                        # D2H into engine-owned pinned buffers (one set per batch of the call, reused by later
                        # calls: pinned allocation costs tens of milliseconds)
                        key = (int(f.shape[0]), i)
                        dst = self._host_pool.get(key)
                        if dst is None:
                            dst = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in out.items()}
                            self._host_pool[key] = dst
                        for k, v in out.items():
                            dst[k].copy_(v, non_blocking=True)
                        keep.append(out)
                        out = dict(dst)
                outs.append(out)
        finally:
            self.lib.mnx_set_encoder_cta_limit(self.h, 0)
            if wide_ok:
                self.lib.mnx_set_context(self.h, 0)
                self.lib.mnx_set_wide_rows(self.h, 0)
                self.set_decode_path(prev_path)
            for d in decs:
                cur.wait_stream(d)
            cur.wait_stream(enc)
        if host:
            for d in decs:
                d.synchronize()
        for o in outs:
            for v in o.values():
                if v.is_cuda:
                    v.record_stream(cur)
        return outs

    # ------------------------------------------------------------------ decode path / contexts
    DECODE_PATHS = {"auto": 0, "graph": 1, "cluster": 2, "cluster16": 3, "wide": 6}

    def set_decode_path(self, path) -> None:
        """"auto" (lowest single-batch latency), "graph", "cluster", "cluster16" or "wide" (the throughput kernel:
        8-CTA clusters of <= 16 rows).  Results are identical on every path."""
        code = self.DECODE_PATHS[path] if isinstance(path, str) else int(path)
        self._check(self.lib.mnx_set_decode_path(self.h, code), "mnx_set_decode_path")
        self._decode_path = code

    def reserve_contexts(self, n: int) -> None:
        self._check(self.lib.mnx_reserve_contexts(self.h, n), "mnx_reserve_contexts")

    def set_context(self, i: int) -> None:
        self._check(self.lib.mnx_set_context(self.h, i), "mnx_set_context")

    # ------------------------------------------------------------------ introspection
    def launch_count(self) -> int:
        return int(self.lib.mnx_launch_count(self.h))

    def last_decode_steps(self) -> int:
        return int(self.lib.mnx_last_decode_steps(self.h))

    def time_kernel(self, which: int, iters: int = 20) -> float:
        ms = C.c_float(0.0)
        self._check(self.lib.mnx_time_kernel(self.h, which, iters, C.byref(ms), self._stream()), "mnx_time_kernel")
        return float(ms.value)
