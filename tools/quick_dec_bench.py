import sys, time, torch
sys.path.insert(0, '.')
from molnextr_b200 import synth
from molnextr_b200.engine import Engine
from tests.helpers import seeded_features
ck = {"decoder": synth.decoder_state(0, "fixed480"), "encoder": None}
eng = Engine(ck, max_batch=32)
f = seeded_features(1, 32, 144).cuda()
for i in range(2): out = eng.decode_greedy(f)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(3): out = eng.decode_greedy(f)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/3
print("decode bs=32 T=480: %.2f ms, steps %d, %.1f us/step, %.0f img/s" % (ms, eng.last_decode_steps(), ms*1000/480, 32/ms*1000))
for w,name in [(1,'cross-attn'),(2,'self-attn t=240'),(3,'ln1+qkv'),(4,'sum+ln+W1'),(5,'W2+res'),(6,'pick')]:
    print(name, "%.2f us" % (eng.time_kernel(w, 50)*1000))
