"""Host-side mirror of the reference's `Encoder` / `Decoder` modules for the accelerated path
(MolNexTR/components.py:110-174, :403-492): same constructor shape, same call signatures, same
return structures -- but every tensor operation happens inside libmolnextr_b200.so."""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from .engine import Engine, MAX_LEN
from .tokenization import CharTokenizer

FORMAT_INFO = {"chartok_coords": {"max_len": MAX_LEN}}   # MolNexTR/utils.py:12-26 (this format only)


class Encoder:
    """`features, hiddens = encoder(images)`.  `hiddens` (per-stage maps) is returned empty: the
    reference collects them but no consumer on the inference path reads them."""

    def __init__(self, engine: Engine):
        self.engine = engine
        self.n_features = engine.encoder_dim

    def __call__(self, images: torch.Tensor, refs=None):
        return self.forward(images, refs)

    def forward(self, images: torch.Tensor, refs=None):
        return self.engine.encode(images), []


class Decoder:
    """`predictions = decoder.decode(features, hiddens)` with the reference's output schema:
    [{'chartok_coords': {'smiles','symbols','coords','indices'}, 'edges': [[int]]}, ...]
    (+ 'atom_scores', 'edge_scores', 'overall_score' when compute_confidence is set,
    components.py:456-469,485-491)."""

    def __init__(self, engine: Engine, tokenizer: Optional[dict] = None, compute_confidence: bool = False):
        self.engine = engine
        self.tokenizer = tokenizer or {"chartok_coords": engine.tok}
        self.formats = ["chartok_coords", "edges"]
        self.compute_confidence = compute_confidence

    def decode(self, encoder_out: torch.Tensor, hiddens=None, refs=None, beam_size: int = 1, n_best: int = 1) -> List[dict]:
        eng = self.engine
        tok: CharTokenizer = self.tokenizer["chartok_coords"]
        if beam_size != 1:
            # the reference's beam branch cannot run; this is the repaired algorithm
            # of oracle/restate.py beam_decode.  As in Decoder.decode (components.py:455,477) only the
            # best hypothesis of each image goes on to the atom scan and the bond head.
            bo = eng.decode_beam(encoder_out, beam_size, n_best)
            out = {"ids": bo["ids"][:, 0].contiguous(), "lens": bo["lens"][:, 0].contiguous(),
                   "logp": bo["logp"][:, 0].contiguous()}
        else:
            out = eng.decode_greedy(encoder_out)
        atom_idx, n_atoms = eng.atom_indices(out["ids"], out["lens"])
        if self.compute_confidence:
            edges, escore = eng.edges(atom_idx, n_atoms, return_scores=True)
            # atom scores / average token score / overall score on the device (mnx_confidence); the k x k edge-score
            # lists of the result schema still come to the host
            atom_sc, _, overall = eng.confidence(out["ids"], out["lens"], out["logp"], escore)
            atom_sc, overall = atom_sc.cpu().numpy(), overall.cpu().numpy()
            escore = escore.cpu().numpy()
        else:
            edges = eng.edges(atom_idx, n_atoms)
        ids, lens, logp = out["ids"].cpu().numpy(), out["lens"].cpu().numpy(), out["logp"].cpu().numpy()
        edges, n_atoms = edges.cpu().numpy(), n_atoms.cpu().numpy()
        atom_idx = atom_idx.cpu().numpy()
        predictions = []
        for i in range(ids.shape[0]):
            L = int(lens[i])
            seq = ids[i, :L].tolist()
            ct = tok.sequence_to_smiles(seq)
            k = len(ct["indices"])
            if k != int(n_atoms[i]) or ct["indices"] != atom_idx[i, :k].tolist():
                raise RuntimeError("device atom scan disagrees with the tokenizer")   # never expected
            pred = {"chartok_coords": ct, "edges": edges[i, :k, :k].astype(int).tolist()}
            if self.compute_confidence:
                ct["atom_scores"] = atom_sc[i, :k].astype(np.float64).tolist()
                pred["edge_scores"] = escore[i, :k, :k].astype(np.float64).tolist()
                pred["overall_score"] = float(overall[i])
            predictions.append(pred)
        return predictions
