"""Sweep the host-buffer entry's transfer modes on the GPU box: python tools/e2e_modes.py [n_points]
(slice size, host threads, share of slices sent as whole records).  Not the bench: picks defaults."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointcloudhookup_b200 import synth, pipeline, _native

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
pinned = torch.empty(n * 34, dtype=torch.uint8, pin_memory=True)
synth.corridor_records(n, max(2, n // 2_000_000), "hilly", 3, out=pinned.numpy())
lib = _native.lib()
stage = pipeline._acquire_staging(n * 12)
for t in (4, 8, 12, 14, 15, 16, 24, 32):
    lib.pch_host_pack_xyz(pinned.data_ptr(), n, 34, stage.data_ptr(), t)
    t0 = time.perf_counter()
    lib.pch_host_pack_xyz(pinned.data_ptr(), n, 34, stage.data_ptr(), t)
    dt = time.perf_counter() - t0
    print(f"gather only: {t:2d} threads {dt*1e3:7.2f} ms  {n*34/dt/1e9:6.1f} GB/s read", flush=True)


def run(reps=4, **kw):
    pipeline.run_pipeline_from_host(pinned, n, 34, synth.SCALES, synth.OFFSETS, 0.1, 500000, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = pipeline.run_pipeline_from_host(pinned, n, 34, synth.SCALES, synth.OFFSETS, 0.1, 500000, **kw)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, len(r.towers)


print("full records:", run(pack="none"), flush=True)
for sc in (5, 10, 20):
    for threads in (0, 15, 14):
        for raw_every in (0, 8, 6, 5, 4):
            ms, tw_ = run(pack="xyz", slice_chunks=sc, threads=threads, raw_every=raw_every)
            print(f"xyz slice_chunks={sc:2d} threads={threads:2d} raw_every={raw_every}: {ms:7.2f} ms/step  ({n/ms/1e6:.2f} Gpt/s) towers={tw_}", flush=True)
