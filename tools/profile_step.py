"""Workload for ncu: one bs=32 image->ids pass (Swin-B encoder + decode); run under
  ncu --metrics gpu__time_duration.sum --clock-control none -c <N> --csv --log-file ... python tools/profile_step.py
Only the first <N> launches are profiled (encoder, K/V precompute, first decode steps)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from molnextr_b200 import synth  # noqa: E402
from molnextr_b200.engine import Engine  # noqa: E402

B = int(os.environ.get("PROFILE_BATCH", "32"))
ck = synth.synthetic_checkpoint(0, "fixed480", encoder=os.environ.get("PROFILE_ENCODER", "swin_base"))
eng = Engine(ck, max_batch=B)
x = torch.randn((B, 3, 384, 384), generator=torch.Generator().manual_seed(0)).cuda()
if os.environ.get("PROFILE_WARM", "0") == "1":
    eng.predict(x)
if os.environ.get("PROFILE_ENCODE_ONLY", "0") == "1":
    eng.encode(x)
    torch.cuda.synchronize()
    print("encode only")
    sys.exit(0)
out = eng.predict(x)
torch.cuda.synchronize()
print("steps", eng.last_decode_steps(), "launches", eng.launch_count())
