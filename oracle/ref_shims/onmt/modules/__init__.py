"""onmt.modules restated: MultiHeadedAttention (OpenNMT-py 2.2.0 semantics, as used at
MolNexTR/models/decoder.py:61-66,144-151,213-215,269-276)."""
import math

import torch
import torch.nn as nn


class MultiHeadedAttention(nn.Module):
    def __init__(self, head_count, model_dim, dropout=0.1, max_relative_positions=0):
        assert model_dim % head_count == 0
        super().__init__()
        self.dim_per_head = model_dim // head_count
        self.model_dim = model_dim
        self.head_count = head_count
        self.linear_keys = nn.Linear(model_dim, head_count * self.dim_per_head)
        self.linear_values = nn.Linear(model_dim, head_count * self.dim_per_head)
        self.linear_query = nn.Linear(model_dim, head_count * self.dim_per_head)
        self.softmax = nn.Softmax(dim=-1)
        self.dropout = nn.Dropout(dropout)
        self.final_linear = nn.Linear(model_dim, model_dim)
        assert max_relative_positions == 0, "relative positions are unused by the reference"
        self.max_relative_positions = max_relative_positions

    def forward(self, key, value, query, mask=None, layer_cache=None, attn_type=None):
        batch_size = key.size(0)
        dim_per_head = self.dim_per_head
        head_count = self.head_count

        def shape(x):
            return x.view(batch_size, -1, head_count, dim_per_head).transpose(1, 2)

        def unshape(x):
            return x.transpose(1, 2).contiguous().view(batch_size, -1, head_count * dim_per_head)

        if layer_cache is not None:
            if attn_type == "self":
                query, key, value = (self.linear_query(query), self.linear_keys(query),
                                     self.linear_values(query))
                key = shape(key)
                value = shape(value)
                if layer_cache["self_keys"] is not None:
                    key = torch.cat((layer_cache["self_keys"], key), dim=2)
                if layer_cache["self_values"] is not None:
                    value = torch.cat((layer_cache["self_values"], value), dim=2)
                layer_cache["self_keys"] = key
                layer_cache["self_values"] = value
            elif attn_type == "context":
                query = self.linear_query(query)
                if layer_cache["memory_keys"] is None:
                    key, value = self.linear_keys(key), self.linear_values(value)
                    key = shape(key)
                    value = shape(value)
                else:
                    key, value = layer_cache["memory_keys"], layer_cache["memory_values"]
                layer_cache["memory_keys"] = key
                layer_cache["memory_values"] = value
        else:
            key = self.linear_keys(key)
            value = self.linear_values(value)
            query = self.linear_query(query)
            key = shape(key)
            value = shape(value)

        query = shape(query)
        key_len = key.size(2)
        query_len = query.size(2)

        query = query / math.sqrt(dim_per_head)
        query_key = torch.matmul(query, key.transpose(2, 3))
        scores = query_key.float()
        if mask is not None:
            mask = mask.unsqueeze(1)
            scores = scores.masked_fill(mask, -1e18)

        attn = self.softmax(scores).to(query.dtype)
        drop_attn = self.dropout(attn)
        context_original = torch.matmul(drop_attn, value)
        context = unshape(context_original)
        output = self.final_linear(context)
        attns = attn.view(batch_size, head_count, query_len, key_len)
        return output, attns

    def update_dropout(self, dropout):
        self.dropout.p = dropout


class AverageAttention(nn.Module):
    """Import-only placeholder (self_attn_type is always 'scaled-dot' in the reference)."""

    def __init__(self, *a, **k):
        raise NotImplementedError("AverageAttention is never built by the reference path")
