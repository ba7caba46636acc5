"""Drop-in for `MolNexTR.model.molnextr` (MolNexTR/model.py:32-196): same constructor arguments,
same `predict_*` methods and result dictionaries, backed by the CUDA engine.

Differences, all deliberate:
  * there is no CPU mode: `device` must be a CUDA device (north_star: no CPU fallback);
  * the checkpoint is loaded strictly (the reference uses strict=False and hides key mismatches);
  * graph -> SMILES/molfile post-processing (MolNexTR/chemical.py, RDKit) is outside the accelerated
    path.  If RDKit and the reference's `chemical` module are importable its per-molecule function runs
    on a persistent worker pool (postprocess.py; the reference forks 16 processes per call);
    otherwise `predicted_smiles` is the decoder's own token-stream SMILES and `predicted_molfile`
    is None, and `postprocess` in the result says so."""
from __future__ import annotations

import argparse
from typing import List, Optional

import numpy as np
import torch

from .components import Decoder, Encoder
from .engine import Engine
from .preprocess import transform
from .tokenization import get_tokenizer

BOND_TYPES = ["", "single", "double", "triple", "aromatic", "solid wedge", "dashed wedge"]


def _load_postprocessor():
    """The reference's own RDKit stage behind a persistent worker pool (postprocess.py); None without RDKit."""
    try:
        from .postprocess import GraphPostProcessor
        return GraphPostProcessor()
    except Exception:
        return None


class molnextr:
    def __init__(self, model_path, device=None, max_batch: int = 16, checkpoint: Optional[dict] = None):
        states = checkpoint if checkpoint is not None else torch.load(model_path, map_location="cpu")
        args = self._get_args(states.get("args"))
        if device is None:
            device = torch.device("cuda")
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("molnextr_b200 runs on CUDA devices only (no CPU fallback)")
        self.device = torch.device("cuda", device.index or 0)
        self.args = args
        self.tokenizer = get_tokenizer(args)
        self.engine = Engine(states, tokenizer=self.tokenizer["chartok_coords"], device=self.device.index,
                             max_batch=max_batch, max_height=args.input_size, max_width=args.input_size)
        self.encoder = Encoder(self.engine)
        self.decoder = Decoder(self.engine, self.tokenizer)
        self.input_size = args.input_size
        self._post = _load_postprocessor()

    @staticmethod
    def _get_args(args_states=None):
        """Inference defaults of the reference (model.py:50-81) overlaid with the checkpoint's args."""
        args = argparse.Namespace(
            encoder="swin_base", decoder="transformer", dec_num_layers=6, dec_hidden_size=256, dec_attn_heads=8,
            enc_pos_emb=False, continuous_coords=False, compute_confidence=False, input_size=384, vocab_file=None,
            coord_bins=64, sep_xy=True, formats=["chartok_coords", "edges"])
        for k, v in (args_states or {}).items():
            setattr(args, k, v)
        return args

    def transform(self, image, keypoints=None):
        return {"image": torch.from_numpy(transform(image, self.input_size)), "keypoints": keypoints or []}

    def predict_images(self, input_images: List, return_atoms_bonds=False, return_confidence=False, batch_size=16):
        predictions = []
        self.decoder.compute_confidence = return_confidence
        if batch_size > self.engine.max_batch:
            # not re-chunked silently: the chunk size is part of the semantics (row-rank positional encoding)
            raise ValueError(f"batch_size {batch_size} exceeds this model's engine capacity (max_batch={self.engine.max_batch}); "
                             f"construct molnextr(..., max_batch={batch_size})")
        # the reference's chunking is part of its semantics: the positional encoding of a row depends on
        # its rank inside its mini-batch, so chunks of `batch_size` are kept as is
        for idx in range(0, len(input_images), batch_size):
            batch = input_images[idx:idx + batch_size]
            # CropWhite -> Resize -> ToGray -> Normalize on the device (bit-exact with self.transform, the cv2 path)
            images = self.engine.preprocess([np.ascontiguousarray(im) for im in batch], size=self.input_size)
            features, hiddens = self.encoder(images)
            predictions += self.decoder.decode(features, hiddens)
        node_coords = [p["chartok_coords"]["coords"] for p in predictions]
        node_symbols = [p["chartok_coords"]["symbols"] for p in predictions]
        edges = [p["edges"] for p in predictions]
        if self._post is not None:
            smiles_list, molblock_list, _ = self._post(node_coords, node_symbols, edges, images=input_images)
            post = "rdkit"
        else:
            smiles_list = [p["chartok_coords"]["smiles"] for p in predictions]
            molblock_list = [None] * len(predictions)
            post = "token-stream SMILES (RDKit post-processing unavailable)"
        outputs = []
        for smiles, molfile, pred in zip(smiles_list, molblock_list, predictions):
            d = {"predicted_smiles": smiles, "predicted_molfile": molfile, "postprocess": post}
            if return_atoms_bonds:
                coords, symbols = pred["chartok_coords"]["coords"], pred["chartok_coords"]["symbols"]
                atoms = []
                for i, (symbol, coord) in enumerate(zip(symbols, coords)):
                    a = {"atom_number": f"{i}", "atom_symbol": symbol, "coords": (round(coord[0], 3), round(coord[1], 3))}
                    if return_confidence:
                        a["confidence"] = pred["chartok_coords"]["atom_scores"][i]
                    atoms.append(a)
                d["atom_sets"] = atoms
                bonds = []
                n = len(symbols)
                for i in range(n - 1):
                    for j in range(i + 1, n):
                        bt = pred["edges"][i][j]
                        if bt != 0:
                            b = {"atom_number": f"{i}", "bond_type": BOND_TYPES[bt], "endpoints": (i, j)}
                            if return_confidence:
                                b["confidence"] = pred["edge_scores"][i][j]
                            bonds.append(b)
                d["bond_sets"] = bonds
            outputs.append(d)
        return outputs

    def predict_image(self, image, return_atoms_bonds=False, return_confidence=False):
        return self.predict_images([image], return_atoms_bonds=return_atoms_bonds, return_confidence=return_confidence)[0]

    def predict_image_files(self, image_files: List, return_atoms_bonds=False, return_confidence=False):
        import cv2
        images = []
        for path in image_files:
            im = cv2.imread(path)
            if im is None:
                raise FileNotFoundError(path)
            images.append(cv2.cvtColor(im, cv2.COLOR_BGR2RGB))
        return self.predict_images(images, return_atoms_bonds=return_atoms_bonds, return_confidence=return_confidence)

    def predict_final_results(self, image_file: str, return_atoms_bonds=False, return_confidence=False):
        return self.predict_image_files([image_file], return_atoms_bonds=return_atoms_bonds,
                                        return_confidence=return_confidence)[0]
