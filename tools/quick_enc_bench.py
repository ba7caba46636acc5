"""Swin-B encoder alone at bs=32 / 384x384 (all SMs), CUDA-event timed over 10 calls after 3 warm-ups."""
import sys, torch
sys.path.insert(0, '.')
from molnextr_b200 import synth
from molnextr_b200.engine import Engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ck = synth.synthetic_checkpoint(0, "fixed480")
eng = Engine(ck, max_batch=B)
x = torch.randn((B, 3, 384, 384), generator=torch.Generator().manual_seed(0)).cuda()
for _ in range(3): eng.encode(x)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10): f = eng.encode(x)
e1.record(); torch.cuda.synchronize()
print("swin encoder ms", e0.elapsed_time(e1) / 10, "checksum", float(f.double().abs().sum()))
