"""CPU: host-side logic that needs no GPU -- the C-ABI library exports, the class table fed to the
device atom scan, preprocessing, and the loud failure when CUDA is absent."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from molnextr_b200 import _cabi
from molnextr_b200.engine import token_class_table
from molnextr_b200.tokenization import CharTokenizer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "molnextr_b200.h")).read()
    declared = set(re.findall(r"\b(mnx_[a-z0-9_]+)\s*\(", header))
    declared.discard("mnx_engine")
    assert declared == set(_cabi.EXPORTS), declared ^ set(_cabi.EXPORTS)
    if not os.path.exists(_cabi.LIB_PATH):
        from molnextr_b200 import build
        build.build()
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


def test_token_class_table_matches_tokenizer():
    tok = CharTokenizer(64)
    tab = token_class_table(tok)
    assert tab.shape == (229,) and tab[101:].sum() == 0
    for i in range(101):
        assert bool(tab[i] & 1) == tok.is_symbol(i)
        assert bool(tab[i] & 2) == tok.is_atom(i)
    assert tab[tok.stoi["["]] & 4 and tab[tok.stoi["]"]] & 8
    assert tab[tok.stoi["C"]] & 16 and tab[tok.stoi["l"]] & 32 and tab[tok.stoi["B"]] & 64 and tab[tok.stoi["r"]] & 128


def test_preprocess_shapes_and_crop():
    from molnextr_b200.preprocess import crop_white, transform
    img = np.full((200, 300, 3), 255, np.uint8)
    img[50:120, 80:200] = 0
    c = crop_white(img)
    assert c.shape == (70 + 100, 120 + 100, 3)
    assert (crop_white(np.full((40, 40, 3), 255, np.uint8)).shape == (140, 140, 3))
    t = transform(img)
    assert t.shape == (3, 384, 384) and t.dtype == np.float32
    assert np.allclose(t[0] * 0.229 + 0.485, t[1] * 0.224 + 0.456, atol=1e-5)   # three equal gray channels


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_engine_refuses_to_run_without_cuda():
    from molnextr_b200 import synth
    from molnextr_b200.engine import Engine, EngineError
    with pytest.raises(EngineError, match="no CPU fallback"):
        Engine({"decoder": synth.decoder_state(0), "encoder": None})


def test_pipeline_plan_schedule():
    """Engine.predict_pipelined's schedule is pure host logic: depth from the SM budget and the co-resident clusters, the
    encoder's CTA cap = what the decode kernels leave free, tail spreading, fallback to one batch at a time."""
    from molnextr_b200.engine import pipeline_plan
    p = pipeline_plan(20, 32, 144, 15, 148)                     # the bench: 20 batches of 32 at 384 x 384 on a B200
    assert p["throughput_kernel"] and p["clusters_per_batch"] == 2 and p["depth"] == 7 and p["encoder_ctas"] == 148 - 7 * 16
    assert p["last_wave"] == 2 and p["tail"] == 6 and p["spread"] == 1
    p = pipeline_plan(15, 32, 144, 15, 148)                     # 7 + 7 + 1: the single tail batch spreads over 4x the clusters
    assert p["last_wave"] == 2 and p["tail"] == 1 and p["spread"] == 4
    p = pipeline_plan(3, 32, 144, 15, 148)                      # fewer batches than the budget allows
    assert p["depth"] == 3 and p["encoder_ctas"] == 148 - 3 * 16
    assert pipeline_plan(8, 32, 144, 15, 148, depth=5)["depth"] == 5          # explicit depth wins
    assert pipeline_plan(8, 16, 144, 15, 148)["depth"] == 8                   # one cluster per batch: up to 14 fit, 8 batches given
    p = pipeline_plan(8, 64, 144, 15, 148)                      # 4 clusters per batch
    assert p["clusters_per_batch"] == 4 and p["depth"] == 3
    for bad in (pipeline_plan(8, 8, 1024, 15, 148), pipeline_plan(8, 256, 144, 15, 148)):   # S > 512, or 16 clusters > 15 resident
        assert not bad["throughput_kernel"] and bad["depth"] == 1 and bad["encoder_ctas"] == 32 and bad["spread"] == 1


def test_pipeline_plan_invariants_fuzz():
    """Whatever the request, the schedule never asks for more co-resident clusters than the GPU holds, always leaves the
    encoder a positive CTA budget, and only spreads the tail by 1, 2 or 4."""
    from hypothesis import given, settings, strategies as st
    from molnextr_b200.engine import pipeline_plan

    @settings(max_examples=300, deadline=None)
    @given(n=st.integers(1, 200), b=st.integers(1, 400), s=st.integers(1, 1100), cl=st.integers(0, 20), sms=st.integers(16, 160),
           depth=st.integers(0, 12))
    def check(n, b, s, cl, sms, depth):
        p = pipeline_plan(n, b, s, cl, sms, depth)
        assert 1 <= p["depth"] <= n and p["encoder_ctas"] >= 16 and p["spread"] in (1, 2, 4)
        assert p["last_wave"] == (n - 1) // p["depth"] and 1 <= p["tail"] <= p["depth"]
        if p["throughput_kernel"]:
            assert s <= 512 and p["clusters_per_batch"] <= cl
            if depth == 0:
                assert p["depth"] * p["clusters_per_batch"] <= cl
                assert p["tail"] * p["spread"] <= p["depth"]          # the spread tail uses no more clusters than a full wave
        else:
            assert p["depth"] == 1 and p["spread"] == 1

    check()
