"""Pipelined vs batch-by-batch throughput at bs = 32 (device-resident and host-buffer modes)."""
import sys, torch
sys.path.insert(0, '.')
from molnextr_b200 import synth
from molnextr_b200.engine import Engine
B, K = 32, 6
ck = synth.synthetic_checkpoint(0, "fixed480")
eng = Engine(ck, max_batch=B)
x = torch.randn((B, 3, 384, 384), generator=torch.Generator().manual_seed(0))
xd, xh = x.cuda(), x.pin_memory()
def t(fn, n=2):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
seq = t(lambda: [eng.predict(xd) for _ in range(K)]) / K
print("batch by batch      : %.2f ms/batch = %.0f img/s" % (seq, B / seq * 1e3))
for ctas in (16, 24, 32, 36, 48):
    p = t(lambda: eng.predict_pipelined([xd] * K, encoder_ctas=ctas)) / K
    print("pipelined ctas=%2d   : %.2f ms/batch = %.0f img/s" % (ctas, p, B / p * 1e3))
ph = t(lambda: eng.predict_pipelined([xh] * K, host=True, encoder_ctas=32)) / K
print("pipelined host bufs : %.2f ms/batch = %.0f img/s" % (ph, B / ph * 1e3))
sh = t(lambda: [eng.predict_host(xh) for _ in range(K)]) / K
print("predict_host        : %.2f ms/batch = %.0f img/s" % (sh, B / sh * 1e3))
