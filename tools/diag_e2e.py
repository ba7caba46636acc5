"""Where does the host-buffer (e2e) arm spend its time?  H2D bandwidth of this box, then predict vs predict_host."""
import sys, time, torch
sys.path.insert(0, '.')
from molnextr_b200 import synth
from molnextr_b200.engine import Engine
B = 32
x_host = torch.randn((B, 3, 384, 384)).pin_memory()
x_dev = torch.empty_like(x_host, device="cuda")
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: x_dev.copy_(x_host, non_blocking=True))
print("H2D pinned 56.6 MB: %.2f ms = %.1f GB/s" % (ms, x_host.numel() * 4 / ms / 1e6))
u8 = torch.randint(0, 255, (B, 384, 384), dtype=torch.uint8).pin_memory()
u8d = torch.empty_like(u8, device="cuda")
ms = t(lambda: u8d.copy_(u8, non_blocking=True))
print("H2D pinned 4.7 MB u8: %.3f ms = %.1f GB/s" % (ms, u8.numel() / ms / 1e6))
r = torch.empty((B, 160, 160), dtype=torch.uint8).pin_memory(); rd = torch.empty_like(r, device="cuda")
ms = t(lambda: r.copy_(rd, non_blocking=True))
print("D2H pinned 0.8 MB: %.3f ms" % ms)
ck = synth.synthetic_checkpoint(0, "fixed480")
eng = Engine(ck, max_batch=B)
print("predict (device in/out): %.2f ms" % t(lambda: eng.predict(x_dev), 3))
print("predict_host           : %.2f ms" % t(lambda: eng.predict_host(x_host), 3))
t0 = time.perf_counter(); eng.predict_host(x_host); print("predict_host wall: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
import subprocess
for period in ("100", "1000"):
    p = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap",
                          "--format=csv,noheader,nounits", "-lms", period], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    time.sleep(0.5)
    print("with nvidia-smi -lms %s: predict %.2f ms, predict_host %.2f ms" % (period, t(lambda: eng.predict(x_dev), 5), t(lambda: eng.predict_host(x_host), 5)))
    p.terminate(); p.wait()
