"""Workload for an ncu launch list of the beam-search path at BASELINE configs[2] size (256 images x beam 5 = 1280 rows)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from molnextr_b200 import synth
from molnextr_b200.engine import Engine
from tests.helpers import seeded_features
B, K = int(os.environ.get("PROFILE_BATCH", "256")), 5
eng = Engine({"decoder": synth.decoder_state(0, "fixed480"), "encoder": None}, max_batch=B, max_beam=K)
f = seeded_features(7, B, 144).cuda()
out = eng.decode_beam(f, K, 1)
torch.cuda.synchronize()
print("steps", eng.last_decode_steps())
