"""One NCSN++ forward (B=8, 256x256) inside a cudaProfilerStart/Stop window, for ncu:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_forward.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from storm_b200 import BackboneRegistry, ncsnpp_spec  # noqa: E402

B = int(os.environ.get("PF_B", "8"))
T = int(os.environ.get("PF_T", "256"))
name = os.environ.get("PF_MODEL", "ncsnpp")
m = BackboneRegistry.get_by_name(name)()
m.load_state_dict(ncsnpp_spec.random_state_dict(m.cfg, seed=0, init_scale=1.0))
m.cuda()
g = torch.Generator().manual_seed(0)
x = (torch.complex(torch.randn(B, 2, 256, T, generator=g), torch.randn(B, 2, 256, T, generator=g)) * 0.3).cuda()
t = torch.full((B,), 0.5, device="cuda")
for _ in range(2):
    y = m(x, t)
torch.cuda.synchronize()
torch.cuda.profiler.start()
y = m(x, t)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    y = m(x, t)
e1.record()
torch.cuda.synchronize()
print(f"forward B={B} 256x{T}: {e0.elapsed_time(e1) / 3:.2f} ms")
