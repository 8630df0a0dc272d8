"""Time the voxel stage with different library builds (PCH_LIB_PATH set by the caller)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes
from pointcloudhookup_b200 import synth, device as dv, _native
n = int(float(os.environ.get("VB_POINTS", "50e6")))
pinned = torch.empty(n * 34, dtype=torch.uint8, pin_memory=True)
synth.corridor_records(n, max(2, n // 2_000_000), "hilly", 3, out=pinned.numpy())
dl = dv.upload_records(pinned, n, 34, synth.SCALES, synth.OFFSETS)
lib = _native.lib()
for _ in range(2): dv.voxel_downsample(dl, 0.1, 500000, want=("f32",))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(3): dv.voxel_downsample(dl, 0.1, 500000, want=("f32",))
e1.record(); torch.cuda.synchronize()
wall = e0.elapsed_time(e1) / 3
lib.pch_profile_enable(1)
for _ in range(3): dv.voxel_downsample(dl, 0.1, 500000, want=("f32",))
buf = ctypes.create_string_buffer(65536); lib.pch_profile_report(buf, 65536)
out = {}
for line in buf.value.decode().splitlines():
    nm, c, t = line.split(); out[nm] = (int(c) // 3, float(t) / 3)
r = dv.voxel_downsample(dl, 0.1, 500000, want=("f32",))
chk = int(r.f32.view(torch.int32).to(torch.int64).sum().item()) ^ r.count          # identical across variants or the variant is wrong
print(os.environ.get("PCH_LIB_PATH", "default").split("_")[-1], "n", n, "M", r.count, "chk", chk, "stage_ms", round(wall, 3),
      {k: (v[0], round(v[1], 3)) for k, v in sorted(out.items(), key=lambda kv: -kv[1][1])[:7]}, flush=True)
