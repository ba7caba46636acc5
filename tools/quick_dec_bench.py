import sys, time, torch
sys.path.insert(0, '.')
from molnextr_b200 import synth
from molnextr_b200.engine import Engine
from tests.helpers import seeded_features
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ck = {"decoder": synth.decoder_state(0, "fixed480"), "encoder": None}
eng = Engine(ck, max_batch=B)
print("max co-resident clusters: 8-CTA", eng.time_kernel(1000, 1), " 16-CTA", eng.time_kernel(1002, 1), " wide", eng.time_kernel(1004, 1))
f = seeded_features(1, B, 144).cuda()
for i in range(2): out = eng.decode_greedy(f)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(3): out = eng.decode_greedy(f)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/3
print("path", eng.time_kernel(1003, 1)); print("decode bs=%d T=480: %.2f ms, steps %d, %.1f us/step, %.0f img/s, launches %d" % (B, ms, eng.last_decode_steps(), ms*1000/480, B/ms*1000, eng.launch_count()))
import os
if os.environ.get("MNX_DECODE_PROFILE"):
    eng.time_kernel(1001, 1)
