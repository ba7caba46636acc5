"""onmt.utils.misc.sequence_mask restated (imported at MolNexTR/models/decoder.py:13, unused
on the inference path)."""
import torch


def sequence_mask(lengths, max_len=None):
    batch_size = lengths.numel()
    max_len = max_len or lengths.max()
    return (torch.arange(0, max_len, device=lengths.device).type_as(lengths)
            .repeat(batch_size, 1).lt(lengths.unsqueeze(1)))
