import sys, torch
sys.path.insert(0, '.')
from molnextr_b200 import synth
from molnextr_b200.engine import Engine
B = 32
ck = synth.synthetic_checkpoint(0, "fixed480", encoder="convnext_base")
eng = Engine(ck, max_batch=B)
x = torch.randn((B, 3, 384, 384), generator=torch.Generator().manual_seed(0)).cuda()
for _ in range(3): eng.encode(x)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(5): eng.encode(x)
e1.record(); torch.cuda.synchronize()
print("convnext encoder ms", e0.elapsed_time(e1) / 5)
us = [1000 * eng.time_kernel(101 + s, 50) for s in range(4)]
depth = (3, 3, 27, 3)
print("dwconv us per stage", us, "total ms", sum(d * u for d, u in zip(depth, us)) / 1000)
