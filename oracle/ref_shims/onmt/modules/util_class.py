"""onmt.modules.util_class.Elementwise restated (used at MolNexTR/models/embedding.py:149)."""
import torch
import torch.nn as nn


class Elementwise(nn.ModuleList):
    def __init__(self, merge=None, *args):
        assert merge in [None, "first", "concat", "sum", "mlp"]
        self.merge = merge
        super().__init__(*args)

    def forward(self, emb):
        emb_ = [feat.squeeze(2) for feat in emb.split(1, dim=2)]
        assert len(self) == len(emb_)
        emb_out = [f(x) for f, x in zip(self, emb_)]
        if self.merge == "first":
            return emb_out[0]
        elif self.merge == "concat" or self.merge == "mlp":
            return torch.cat(emb_out, 2)
        elif self.merge == "sum":
            return sum(emb_out)
        return emb_out
