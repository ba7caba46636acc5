"""Pipelined vs batch-by-batch throughput at bs = 32 (device-resident and host-buffer modes)."""
import sys, torch
sys.path.insert(0, '.')
from molnextr_b200 import synth
from molnextr_b200.engine import Engine
B = 32
K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
depths = [int(a) for a in sys.argv[2:]] or [3, 4, 5, 6]
ck = synth.synthetic_checkpoint(0, "fixed480")
eng = Engine(ck, max_batch=B)
x = torch.randn((B, 3, 384, 384), generator=torch.Generator().manual_seed(0))
xd, xh = x.cuda(), x.pin_memory()
def t(fn, n=1):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
seq = t(lambda: [eng.predict(xd) for _ in range(4)]) / 4
print("batch by batch      : %.2f ms/batch = %.0f img/s" % (seq, B / seq * 1e3), flush=True)
for d in depths:
    p = t(lambda: eng.predict_pipelined([xd] * K, depth=d)) / K
    print("pipelined depth=%d K=%d : %.2f ms/batch = %.0f img/s" % (d, K, p, B / p * 1e3), flush=True)
ph = t(lambda: eng.predict_pipelined([xh] * K, host=True)) / K
print("pipelined host bufs (auto depth) : %.2f ms/batch = %.0f img/s" % (ph, B / ph * 1e3), flush=True)
