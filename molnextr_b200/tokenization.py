"""Host-side tokenizer for the `chartok_coords` output format.

Mirrors the public surface of the reference `CharTokenizer` (MolNexTR/tokenization.py:330-515,
with the `NodeTokenizer` helpers at :113-180) for the inference path only: id classes,
the grammar mask applied between decode steps, and `sequence_to_smiles`, which turns a
greedy id stream into `{smiles, symbols, coords, indices}`.  The training-time encoders
(`smiles_to_sequence`, `nodes_to_sequence`, ...) are out of scope.

The device kernels implement the same grammar mask from three integers
(`offset`, `maxx`, `maxy`); `grammar_rule()` exposes them.
"""
from __future__ import annotations

import json
import os

PAD_ID, SOS_ID, EOS_ID, UNK_ID, MASK_ID = 0, 1, 2, 3, 4
PAD, SOS, EOS, UNK, MASK = "<pad>", "<sos>", "<eos>", "<unk>", "<mask>"

_VOCAB_DIR = os.path.join(os.path.dirname(__file__), "vocab")
FORMAT_MAX_LEN = {"chartok_coords": 480}  # MolNexTR/utils.py:25


class CharTokenizer:
    """Character-level atom tokens followed by an X-bin and a Y-bin token."""

    def __init__(self, input_size=64, path=None, sep_xy=True, continuous_coords=False):
        if path is None:
            path = os.path.join(_VOCAB_DIR, "vocab_chars.json")
        with open(path) as f:
            self.stoi = json.load(f)
        self.itos = {i: s for s, i in self.stoi.items()}
        self.maxx = input_size
        self.maxy = input_size
        self.sep_xy = sep_xy
        self.continuous_coords = continuous_coords
        self.special_tokens = [PAD, SOS, EOS, UNK, MASK]

    # ---- sizes -----------------------------------------------------------------
    @property
    def offset(self):
        return len(self.stoi)

    def __len__(self):
        if self.sep_xy:
            return self.offset + self.maxx + self.maxy
        return self.offset + max(self.maxx, self.maxy)

    @property
    def output_constraint(self):
        return not self.continuous_coords

    def grammar_rule(self):
        """(offset, maxx, maxy) consumed by the device-side mask (sep_xy layout only)."""
        if not self.sep_xy or self.continuous_coords:
            raise NotImplementedError("device grammar mask implements the sep_xy discrete layout")
        return self.offset, self.maxx, self.maxy

    # ---- id classes (tokenization.py:131-150) ------------------------------------
    def is_x(self, i):
        return self.offset <= i < self.offset + self.maxx

    def is_y(self, i):
        if self.sep_xy:
            return self.offset + self.maxx <= i
        return self.offset <= i

    def is_symbol(self, i):
        return len(self.special_tokens) <= i < self.offset or i == UNK_ID

    @staticmethod
    def is_atom_token(token):
        return token.isalpha() or token.startswith("[") or token == "*" or token == UNK

    def is_atom(self, i):
        return self.is_symbol(i) and self.is_atom_token(self.itos[i])

    def id_to_x(self, i):
        return (i - self.offset) / (self.maxx - 1)

    def id_to_y(self, i):
        if self.sep_xy:
            return (i - self.offset - self.maxx) / (self.maxy - 1)
        return (i - self.offset) / (self.maxy - 1)

    # ---- grammar mask (tokenization.py:383-392): True = forbidden next id ---------
    def get_output_mask(self, i):
        n = len(self)
        if self.continuous_coords:
            return [False] * n
        if self.is_x(i):       # after an X-bin only Y-bins may follow
            return [True] * (self.offset + self.maxx) + [False] * self.maxy
        if self.is_y(i):       # after a Y-bin only symbol/special ids may follow
            return [False] * self.offset + [True] * (self.maxx + self.maxy)
        return [False] * n

    # ---- id stream -> graph nodes (tokenization.py:464-515) -----------------------
    def sequence_to_smiles(self, sequence):
        seq = list(sequence)
        n = len(seq)
        has_coords = not self.continuous_coords
        pieces, coords, symbols, indices = [], [], [], []
        i = 0
        while i < n:
            tok = seq[i]
            if tok == EOS_ID or tok == PAD_ID:
                break
            if self.is_x(tok) or self.is_y(tok):
                i += 1
                continue
            if not self.is_atom(tok):
                pieces.append(self.itos[tok])
                i += 1
                continue
            ch = self.itos[tok]
            if ch == "[":
                j = i + 1
                while j < n and self.is_symbol(seq[j]):
                    closing = self.itos[seq[j]] == "]"
                    j += 1
                    if closing:
                        break
            else:
                j = i + 1
                if j < n and self.is_symbol(seq[j]):
                    nxt = self.itos[seq[j]]
                    if (ch == "C" and nxt == "l") or (ch == "B" and nxt == "r"):
                        j = i + 2
            token = "".join(self.itos[seq[k]] for k in range(i, j))
            pieces.append(token)
            if has_coords:
                if j + 2 < n and self.is_x(seq[j]) and self.is_y(seq[j + 1]):
                    coords.append([self.id_to_x(seq[j]), self.id_to_y(seq[j + 1])])
                    symbols.append(token)
                    indices.append(j + 2)
                    i = j + 2
                else:
                    i = j
            else:
                if j < n:
                    symbols.append(token)
                    indices.append(j)
                i = j
        out = {"smiles": "".join(pieces), "symbols": symbols, "indices": indices}
        if has_coords:
            out["coords"] = coords
        return out


def get_tokenizer(args):
    """Same contract as the reference factory (tokenization.py:518-544) for the supported format."""
    tok = {}
    for fmt in args.formats:
        if fmt == "chartok_coords":
            path = getattr(args, "vocab_file", None)
            tok[fmt] = CharTokenizer(args.coord_bins, path, args.sep_xy,
                                     continuous_coords=getattr(args, "continuous_coords", False))
        elif fmt == "edges":
            continue
        else:
            raise NotImplementedError(f"format {fmt!r} is outside the accelerated path (SURVEY.md 8f-4)")
    return tok
