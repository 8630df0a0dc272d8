#!/usr/bin/env python
"""bench.py — points/s of the per-point LAS hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload pipeline|voxel_geoid|corridor400M|corridor1B_geo] [--points P] [--tiles T]

A "step" is one pass of the hot path over one synthetic corridor tile per GPU:
  pipeline     (default; BASELINE.json configs[2]) 100 M-point hilly corridor, 50 towers per GPU:
               decode -> voxel downsample (0.1 m, 500 k chunks) -> percentile ground filter ->
               chunked DBSCAN -> per-cluster box/centroid reduction -> tower list (+ tower merge
               across ranks).  This is the path north_star quotes its 5 Gpt/s target on.
  voxel_geoid  (configs[1]) 20 M-point flat corridor: voxel downsample + per-point EPSG:4547->4326
               and EGM96 geoid height conversion.
  corridor400M (configs[3]) 400 M-point hilly corridor in 4 spatial tiles over the N GPUs (STRONG scaling: the
               corridor is fixed, a rank holds 4/N consecutive tiles): per tile voxel downsample + grid min-z
               ground removal, then ONE DBSCAN over the whole corridor with an NCCL halo exchange between
               neighbouring ranks (tiles.py) and towers from the all-reduced per-cluster table.
  corridor1B_geo (configs[4]) 1 B points in 8 tiles, the same plus EGM2008 (the reference's simulated 0.25 deg
               grid) geoid conversion and EPSG:4547->4326 of EVERY point (test/005test.py:37-66).
`value` is whole-job input points/s with the records resident in HBM; `e2e` is the same metric through
the host-buffer call (pinned host records -> H2D -> pipeline -> D2H of the results) every step.
`--impl reference` times the CPU oracle (numpy + real scikit-learn, all host threads) on a bounded
prefix of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "input points/s through the LAS hot path (decode->downsample->ground->tower)"
UNIT = "points/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="pipeline", choices=["pipeline", "voxel_geoid", "corridor400M", "corridor1B_geo"])
    ap.add_argument("--points", type=float, default=None, help="points per GPU (pipeline, voxel_geoid) / in the whole corridor (corridor*)")
    ap.add_argument("--tiles", type=int, default=None, help="spatial tiles of the corridor workloads (default 4 / 8)")
    ap.add_argument("--no-modes", action="store_true", help="skip the extra timed regions (ground=grid, box=obb) of the default line")
    ap.add_argument("--box", default="aabb", choices=["aabb", "obb"])
    ap.add_argument("--ground", default="percentile", choices=["percentile", "grid"])
    ap.add_argument("--ref-sample", type=float, default=None, help="points in the CPU sample (reference arm)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU sample (profiling runs)")
    return ap.parse_args()


TILED = ("corridor400M", "corridor1B_geo")


def workload_config(args):
    if args.workload == "pipeline":
        n = int(args.points or 100e6)
        towers = max(2, round(n / 2e6))
        return dict(workload="pipeline: 100M-pt hilly corridor tile, 50 towers, voxel 0.1 m chunk 500k -> pct25+3 m "
                             "ground filter -> DBSCAN(8,80) on 50k chunks -> cluster reduce -> towers",
                    n=n, towers=towers, terrain="hilly", seed=3, voxel=0.1, chunk=500000)
    if args.workload == "voxel_geoid":
        n = int(args.points or 20e6)
        return dict(workload="voxel_geoid: 20M-pt flat corridor, voxel 0.1 m chunk 500k + per-point EPSG:4547->4326 + "
                             "EGM96-style geoid height", n=n, towers=max(2, round(n / 2e6)), terrain="flat", seed=2,
                    voxel=0.1, chunk=500000)
    geo = args.workload == "corridor1B_geo"
    total = int(args.points or (1e9 if geo else 400e6))
    tiles = int(args.tiles or (8 if geo else 4))
    per_tile = total // tiles
    towers = max(2, round(per_tile / 2e6))
    name = ("corridor1B_geo: 1B-pt hilly corridor in 8 spatial tiles; per tile voxel 0.1 m chunk 500k + EGM2008 "
            "(simulated 0.25 deg grid) + EPSG:4547->4326 of every point + grid min-z ground (2 m cells, 3 m); ONE "
            "DBSCAN(8,80) over the corridor with NCCL halo exchange; towers from the all-reduced cluster table"
            if geo else
            "corridor400M: 400M-pt hilly corridor in 4 spatial tiles; per tile voxel 0.1 m chunk 500k + grid min-z "
            "ground (2 m cells, 3 m); ONE DBSCAN(8,80) over the corridor with NCCL halo exchange; towers from the "
            "all-reduced cluster table")
    return dict(workload=name, n=per_tile, total=per_tile * tiles, tiles=tiles, towers=towers, terrain="hilly",
                seed=5 if geo else 4, voxel=0.1, chunk=500000, geo=geo)


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions (NVML from a thread, every 20 ms — often
    enough for several samples inside the shortest timed region, rare enough not to take the GIL from the
    thread that launches the kernels; the nvidia-smi loop of B200_PROFILING.md is the fallback)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop_flag = threading.Event()
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        try:
            self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)))
        except Exception:
            pass
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                try:
                    r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, b in bits.items():
                    if r & b:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.th.join(1.0)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "samples": len(self.sm), "reasons": sorted(self.reasons), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernel_bytes_model(cfg, info):
    """ALGORITHMIC bytes per launch for each kernel (DESIGN.md 'kernels'): every input read once,
    every output written once.  n = input points, M = voxels, G = candidates, R = 34."""
    n, M, G, R = cfg["n"], info.get("M", 0), info.get("G", 0), 34
    return {
        "k_chunk_minmax": n * (R + 16),
        "k_voxel_keys16": n * (16 + 8),
        "k_hist": None,          # mixed (voxel sort + DBSCAN sort): resolved from launches below
        "k_pass": None,
        "k_voxel_reduce": n * 8 + n * 16 + M * 12 + M * 4,    # keys + lattice gathers -> float32 cloud + z column
        "k_seq_sum_serial": M * 12,
        "k_shift": M * 12 + M * 4,
        "k_sel_hist": M * 4,
        "k_compact": M * 12 + G * 12,                         # keep flag derived from the cloud itself
        "k_compact_xyz": M * 12 + G * 12,
        "k_compact_flags": G,
        "k_column": M * 12 + M * 4,
        "k_sum_prep": M * 12,
        "k_sum_tables": M * 12,
        "k_grid_min": M * 12,
        "k_minmax_f32": M * 12,
        "k_db_bounds": G * 12,
        "k_db_labels_core": G * (16 + 1 + 4) + G * 4,
        "k_db_keys": G * 12 + G * 8,
        "k_db_cells": G * 8 + G * 12 + G * (16 + 4 + 4),
        "k_db_core": G * 16 + G,
        "k_db_labels": G * 16 + G * 4,
        "k_las_geodetic": n * R + n * 24,
    }


# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from pointcloudhookup_b200 import _native, device as dv, dist as pdist, pipeline, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = workload_config(args)
    n = cfg["n"]
    lib = _native.lib()

    # ---- workload: one tile per rank, generated straight into pinned host memory
    t0 = time.time()
    pinned = torch.empty(n * 34, dtype=torch.uint8, pin_memory=True)
    synth.corridor_records(n, cfg["towers"], cfg["terrain"], cfg["seed"] + rank,
                           s_origin=pdist.tile_for_rank(rank, cfg["towers"]), out=pinned.numpy())
    gen_s = time.time() - t0
    grid = None
    if args.workload == "voxel_geoid":
        from pointcloudhookup_b200 import geo
        lat = np.linspace(-90, 90, 721)
        lon = -180 + 0.25 * np.arange(1440)
        g = (30 * np.sin(np.radians(lat))[:, None] * np.cos(np.radians(lon))[None, :]).astype(np.float32)
        grid = geo.upload_grid(geo.HostGrid(-90.0, -180.0, 0.25, 0.25, g), dev)

    info = {}

    def step(dl, ground=None, box=None):
        if args.workload == "pipeline":
            res = pipeline.run_pipeline(dl, cfg["voxel"], cfg["chunk"], ground=ground or args.ground, box=box or args.box)
            info.update(M=res.n_voxels, G=res.n_candidates, K=res.n_clusters, towers=len(res.towers),
                        voxel_passes=(res.voxel_plan or {}).get("n_passes", 0),
                        db_passes=(res.db_plan or {}).get("n_passes", 0))
            merged = pdist.merge_towers(res.towers) if world > 1 else res.towers
            info["towers_merged"] = len(merged)
            return res.n_clusters * 56 + 64
        from pointcloudhookup_b200 import geo
        v = dv.voxel_downsample(dl, cfg["voxel"], cfg["chunk"], want=("lattice",))
        out = geo.las_to_geodetic(dl, grid, -1.0, geo.EPSG4547)
        chk = out[:: max(1, n // 1024), 2].sum().item()   # force completion, tiny D2H
        info.update(M=v.count, G=0, K=0, checksum=chk, voxel_passes=v.plan["n_passes"], db_passes=0)
        return 8 + 64

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dl = dv.upload_records(pinned, n, 34, synth.SCALES, synth.OFFSETS, dev)
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        step(dl)
    barrier()

    # ---- timed region 1: records resident in HBM; per-kernel CUDA events on the launching stream
    sampler = ClockSampler(local)
    sampler.start()
    lib.pch_profile_enable(1)
    l0 = lib.pch_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(dl)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = (lib.pch_launch_count() - l0) / args.steps
    lib.pch_profile_enable(0)
    import ctypes
    cbuf = ctypes.create_string_buffer(65536)
    lib.pch_profile_report(cbuf, 65536)
    prof = {}
    for line in cbuf.value.decode().splitlines():
        nm, cnt, tot = line.split()
        prof[nm] = (int(cnt), float(tot))
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = n * world * args.steps / (ms_max / 1e3)

    # ---- the other modes of the same step, each with its own timed region (same rules: warm-up, barrier, CUDA
    # events, max over ranks): north_star's grid min-z ground removal, and the reference's default oriented box
    modes = {}
    if args.workload == "pipeline" and not args.no_modes:
        main_info = dict(info)
        for mname, kw in (("ground_grid", dict(ground="grid")), ("box_obb", dict(box="obb"))):
            if (kw.get("ground") or args.ground) == args.ground and (kw.get("box") or args.box) == args.box:
                continue
            for _ in range(min(2, args.warmup)):
                step(dl, **kw)
            barrier()
            e0.record()
            for _ in range(args.steps):
                step(dl, **kw)
            e1.record()
            barrier()
            tm = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            modes[mname] = {"value": n * world * args.steps / (float(tm.item()) / 1e3), "unit": UNIT,
                            "ms_per_step": float(tm.item()) / args.steps, "ground": kw.get("ground", args.ground),
                            "box": kw.get("box", args.box),
                            "stage_info": {k: info.get(k) for k in ("M", "G", "K", "towers")}}
        info.clear()
        info.update(main_info)

    # ---- timed region 2: end to end through host buffers (H2D of the records + results D2H every step)
    e2e = None
    if not args.no_e2e:
        del dl
        barrier()
        d2h = 0

        def e2e_step(pack):
            if args.workload == "pipeline":
                # public host-buffer entry: sliced H2D on a copy stream overlapped with the voxel stage
                res = pipeline.run_pipeline_from_host(pinned, n, 34, synth.SCALES, synth.OFFSETS, cfg["voxel"],
                                                      cfg["chunk"], ground=args.ground, box=args.box, pack=pack)
                if world > 1:
                    pdist.merge_towers(res.towers)
                return res.n_clusters * 56 + 64
            dl2 = dv.upload_records(pinned, n, 34, synth.SCALES, synth.OFFSETS, dev)
            return step(dl2)

        def e2e_run(pack):
            e2e_step(pack)          # one untimed pass: allocator and pinned-page warm-up for this code path
            barrier()
            e0.record()
            for _ in range(args.steps):
                nbytes = e2e_step(pack)
            e1.record()
            barrier()
            tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return {"value": n * world * args.steps / (float(tt.item()) / 1e3), "unit": UNIT,
                    "h2d_bytes_per_step": n * (12 if pack == "xyz" else 34), "d2h_bytes_per_step": int(nbytes),
                    "ms_per_step": float(tt.item()) / args.steps}

        # two transfer modes of the same public call: whole 34-byte records over PCIe, or the 12 X,Y,Z bytes
        # of each record gathered into pinned staging by host threads (no arithmetic on the host).  The headline
        # e2e is the mode the public call picks by itself (pack="auto"); the other one is context.
        names = {"none": "full_records", "xyz": "xyz12_host_gather"}
        auto = pipeline.resolve_pack("auto") if args.workload == "pipeline" else "none"
        emodes = {names[auto]: e2e_run(auto)}
        if args.workload == "pipeline":
            other = "xyz" if auto == "none" else "none"
            emodes[names[other]] = e2e_run(other)
            emodes["xyz12_host_gather"]["host_threads"] = pipeline.host_threads()
        e2e = dict(emodes[names[auto]], mode=names[auto], picked_by='pack="auto"', modes=emodes)
        if args.workload == "pipeline":
            # context, not the headline: the same tiles as a STREAM (run_tiles_from_host): tile k+1 is gathered and
            # copied while tile k is in its ground / tower stages.  Every tile's records cross PCIe and every tile's
            # results are read back inside the timed region.
            def stream(k):
                nbytes = 0
                tiles = ((pinned, n, 34) for _ in range(k))
                for res in pipeline.run_tiles_from_host(tiles, synth.SCALES, synth.OFFSETS, cfg["voxel"], cfg["chunk"],
                                                        ground=args.ground, box=args.box, pack="xyz"):
                    if world > 1:
                        pdist.merge_towers(res.towers)
                    nbytes = res.n_clusters * 56 + 64
                return nbytes
            stream(2)
            barrier()
            e0.record()
            nb = stream(args.steps)
            e1.record()
            barrier()
            tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            emodes["xyz12_tile_stream"] = {"value": n * world * args.steps / (float(tt.item()) / 1e3), "unit": UNIT,
                                          "h2d_bytes_per_step": n * 12, "d2h_bytes_per_step": int(nb),
                                          "ms_per_step": float(tt.item()) / args.steps,
                                          "note": "steps pipelined: tile k+1 upload overlaps tile k ground/tower stages"}
            e2e["modes"] = emodes
        # context: the bare pinned->HBM copy of one step's records (the PCIe floor under e2e)
        buf = torch.empty(n * 34, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        e0.record()
        buf.copy_(pinned, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        e2e["h2d_copy_only_ms"] = e0.elapsed_time(e1)
        e2e["h2d_copy_only_GBps"] = n * 34 / (e2e["h2d_copy_only_ms"] / 1e3) / 1e9
        del buf
        if args.workload == "pipeline":
            # context: the bare host gather (34-byte records -> 12-byte stream in pinned staging), all slices
            stage = pipeline._acquire_staging(n * 12)
            tg = time.perf_counter()
            lib.pch_host_pack_xyz(pinned.data_ptr(), n, 34, stage.data_ptr(), pipeline.host_threads())
            tg = time.perf_counter() - tg
            pipeline._release_staging(stage)
            e2e["host_gather_only_ms"] = tg * 1e3
            e2e["host_gather_read_GBps"] = n * 34 / tg / 1e9

    clocks = sampler.stop()      # sampled through both timed regions

    # ---- roofline of the dominant kernel (from the live per-kernel events of timed region 1)
    peak, peak_src = measured_peak()
    model = kernel_bytes_model(cfg, info)
    pv, pd = info.get("voxel_passes", 0), info.get("db_passes", 0)
    real_launches = {}
    if pv + pd:
        # k_pass is launched for the widest key the chunk size allows (device-side plan, no host round trip); the
        # surplus launches exit at once.  Bytes and launch count below are those of the passes that sort.
        keys_per_launch = (n * pv + info.get("G", 0) * pd) / (pv + pd)   # k_pass: voxel sort + DBSCAN cell sort
        model["k_pass"] = 16 * keys_per_launch
        model["k_hist"] = 8 * info.get("G", 0)                           # only the DBSCAN sort still reads its keys for histograms
        real_launches["k_pass"] = pv + pd
    total_kernel_ms = sum(v[1] for v in prof.values()) or 1.0
    roof = None
    if prof:
        name, (cnt, tot) = max(prof.items(), key=lambda kv: kv[1][1])
        if name in real_launches:
            cnt = real_launches[name] * args.steps
        per_launch_ms = tot / cnt
        alg = model.get(name) or 0
        achieved = alg / (per_launch_ms / 1e3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r2_traffic_100M.json")
        if not os.path.exists(tpath):
            tpath = os.path.join(ROOT, "profiles", "r1_traffic_100M.json")
        if args.workload == "pipeline" and n == 100_000_000 and os.path.exists(tpath):
            # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full
            # capture of this same configuration (profiles/r1_ncu_summary.md)
            traffic = json.load(open(tpath)).get(name) or json.load(open(tpath)).get(name + "<2>")
        roof = {"kernel": name, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg, "launches_per_step": cnt / args.steps,
                "ms_per_launch": per_launch_ms, "share_of_kernel_time": tot / total_kernel_ms}
    kernels = {k: {"launches_per_step": v[0] / args.steps, "ms_per_step": v[1] / args.steps,
                   "GBps": (model.get(k) / ((v[1] / (real_launches.get(k, v[0] / args.steps) * args.steps)) / 1e3) / 1e9)
                   if model.get(k) else None}
               for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}

    # per-stage device time (sum of the stage's kernels, live events) -> input points/s per stage
    stage_of = {"voxel_downsample": ("k_chunk_minmax", "k_init_minmax", "k_voxel_plan", "k_voxel_keys", "k_voxel_keys16",
                                     "k_voxel_reduce"),
                "ground": ("k_sum_", "k_seq_sum", "k_centroid", "k_shift", "k_sel_", "k_compact", "k_grid_", "k_minmax_f32"),
                "tower": ("k_db_",), "geoid_crs": ("k_las_geodetic", "k_gk_inverse", "k_geoid_shift")}
    stages = {}
    sort_ms = sum(v[1] for k, v in prof.items() if k in ("k_hist", "k_scan", "k_pass")) / args.steps
    for st_name, prefixes in stage_of.items():
        ms_st = sum(v[1] for k, v in prof.items() if any(k.startswith(pf) for pf in prefixes)) / args.steps
        if st_name == "voxel_downsample" and (pv + pd):
            ms_st += sort_ms * (n * pv) / max(1, n * pv + info.get("G", 0) * pd)      # the voxel sort's share of k_pass
        if st_name == "tower" and (pv + pd):
            ms_st += sort_ms * (info.get("G", 0) * pd) / max(1, n * pv + info.get("G", 0) * pd)
        if ms_st > 0:
            stages[st_name] = {"ms_per_step": ms_st, "input_points_per_s": n / (ms_st / 1e3)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 (voxel means, distances) / f32 (tower stage) / int32 lattice",
            "data": f"synthetic corridor LAS (PDRF 3, 34 B records), seed {cfg['seed']}+rank, generated in {gen_s:.1f}s",
            "config": {"workload": cfg["workload"], "points_per_gpu": n, "voxel_size": cfg["voxel"],
                       "chunk_size": cfg["chunk"], "ground": args.ground, "box": args.box,
                       "parallelism": f"tile-per-gpu x{world}, tower merge by all_gather",
                       "l2": "inputs (3.4 GB records per step) far exceed the 126 MB L2; no flush needed"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roof,
            "modes": modes, "stages": stages, "stage_info": info, "kernels": kernels}
    if rank == 0:
        line["cpu_baseline"] = None if args.no_cpu_baseline else cpu_baseline(args, cfg, bounded=True)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
def run_tiled(args):
    """corridor400M / corridor1B_geo: a fixed corridor in T spatial tiles over the N GPUs (strong scaling)."""
    import ctypes
    import torch
    import torch.distributed as dist
    from pointcloudhookup_b200 import _native, device as dv, geo, pipeline, synth, tiles as tl

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = workload_config(args)
    T, n = cfg["tiles"], cfg["n"]
    if T % world:
        raise SystemExit(f"{T} tiles do not divide over {world} GPUs")
    mine = list(range(rank * T // world, (rank + 1) * T // world))
    lib = _native.lib()
    comm = tl.TorchComm(dev) if world > 1 else tl.SoloComm()
    axis = pipeline.corridor_axis(synth.AZIMUTH_DEG)

    t0 = time.time()
    pinned = [torch.empty(n * 34, dtype=torch.uint8, pin_memory=True) for _ in mine]

    def gen(i):
        t = mine[i]
        synth.corridor_records(n, cfg["towers"], cfg["terrain"], cfg["seed"] * 100 + t,
                               s_origin=t * cfg["towers"] * synth.SPAN, out=pinned[i].numpy())
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=max(1, min(len(mine), (os.cpu_count() or 4) // max(1, world) // 2 or 1))) as ex:
        list(ex.map(gen, range(len(mine))))         # workload generation only (numpy releases the GIL in its big loops)
    gen_s = time.time() - t0
    grid = None
    if cfg["geo"]:
        path = geo.find_grid_file("egm2008_simulated_0.25deg.npz")
        if path is None:
            raise SystemExit("egm2008_simulated_0.25deg.npz not found (pointcloudhookup_b200/data)")
        grid = geo.load_grid(path, dev)

    info = {}
    # the common float32 frame of the candidates: known to every rank from the LAS header (its offsets) — no collective
    frame = np.array([synth.OFFSETS[0], synth.OFFSETS[1], 100.0], dtype=np.float32)
    geo_sums = []

    def per_tile(dl, vres):
        if grid is None:
            return
        out = geo.las_to_geodetic(dl, grid, 1.0, geo.EPSG4547, chunk_minmax=vres.chunk_minmax)   # the reference's +multiplier=1
        geo_sums.append(out[:: max(1, dl.n // 1024), 2].sum())                                   # read back once per step, below

    def step(dls):
        geo_sums.clear()
        res = pipeline.run_pipeline_tiled(dls, comm, axis, cfg["voxel"], cfg["chunk"], ground="grid", per_tile=per_tile,
                                          origin=frame)
        if geo_sums:
            info["geo_checksum"] = float(torch.stack(geo_sums).sum().item())                     # the converted heights leave the device as a checksum
        info.update(M=res.n_voxels, G=res.n_candidates, K=res.n_clusters, towers=len(res.towers), halo=res.halo)
        return res.n_clusters * 56 + 64

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            nb = fn()
        e1.record()
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()), nb

    dls = [dv.upload_records(p, n, 34, synth.SCALES, synth.OFFSETS, dev) for p in pinned]
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        step(dls)
    sampler = ClockSampler(local)
    sampler.start()
    lib.pch_profile_enable(1)
    l0 = lib.pch_launch_count()
    p2p0 = getattr(comm, "bytes_p2p", 0)
    g0 = getattr(comm, "bytes_gather", 0)
    ms, _ = timed(lambda: step(dls), args.steps)
    launches = (lib.pch_launch_count() - l0) / args.steps
    lib.pch_profile_enable(0)
    cbuf = ctypes.create_string_buffer(65536)
    lib.pch_profile_report(cbuf, 65536)
    prof = {}
    for line in cbuf.value.decode().splitlines():
        nm, cnt, tot = line.split()
        prof[nm] = (int(cnt), float(tot))
    total_pts = cfg["total"]
    value = total_pts * args.steps / (ms / 1e3)
    coll = {"p2p_halo_bytes_per_step_this_rank": (getattr(comm, "bytes_p2p", 0) - p2p0) / args.steps,
            "allgather_bytes_per_step": (getattr(comm, "bytes_gather", 0) - g0) / args.steps,
            "halo_points": info.get("halo")}

    e2e = None
    if not args.no_e2e:
        del dls
        barrier()

        def e2e_step():
            up = [dv.upload_records_xyz(p, n, 34, synth.SCALES, synth.OFFSETS, dev) if pack == "xyz"
                  else dv.upload_records(p, n, 34, synth.SCALES, synth.OFFSETS, dev) for p in pinned]
            return step(up)
        pack = pipeline.resolve_pack("auto")
        e2e_step()
        ems, nb = timed(e2e_step, args.steps)
        e2e = {"value": total_pts * args.steps / (ems / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": len(mine) * n * (12 if pack == "xyz" else 34) * world, "d2h_bytes_per_step": int(nb) * world,
               "ms_per_step": ems / args.steps, "mode": "xyz12_host_gather" if pack == "xyz" else "full_records",
               "picked_by": 'pack="auto"', "note": "tiles uploaded one after another, then the tiled pipeline; no overlap"}
    clocks = sampler.stop()

    peak, peak_src = measured_peak()
    tot_ms = sum(v[1] for v in prof.values()) or 1.0
    name, (cnt, tot) = max(prof.items(), key=lambda kv: kv[1][1])
    pts_rank = len(mine) * n
    model = kernel_bytes_model(dict(cfg, n=n), {"M": info.get("M", 0) / max(1, len(mine)), "G": info.get("G", 0)})
    vp = 4
    model["k_pass"] = 16 * n
    cnt_real = (vp * len(mine) * args.steps) if name == "k_pass" else cnt
    per_launch_ms = tot / max(1, cnt_real)
    alg = model.get(name) or 0
    roof = {"kernel": name, "bound": "hbm", "achieved": alg / (per_launch_ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": alg / (per_launch_ms / 1e3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg, "launches_per_step": cnt_real / args.steps, "ms_per_launch": per_launch_ms,
            "share_of_kernel_time": tot / tot_ms,
            "note": "per-tile launch of this rank; k_pass counts the sorting passes of the voxel sort only"}
    kernels = {k: {"launches_per_step": v[0] / args.steps, "ms_per_step": v[1] / args.steps}
               for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:24]}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 (voxel means, distances, geoid/CRS) / f32 (ground, DBSCAN input) / int32 lattice",
            "data": f"synthetic corridor LAS (PDRF 3, 34 B records), {T} tiles, seed {cfg['seed']}*100+tile, generated in {gen_s:.1f}s",
            "config": {"workload": cfg["workload"], "points_total": total_pts, "tiles": T, "tiles_per_gpu": len(mine),
                       "points_per_tile": n, "voxel_size": cfg["voxel"], "chunk_size": cfg["chunk"], "ground": "grid",
                       "box": "aabb", "dbscan": "whole corridor, halo exchange",
                       "parallelism": f"{T} spatial tiles over {world} GPUs; NCCL send/recv halo (2*eps band), all-gather of "
                                      f"cluster equivalences, all-reduce of the cluster table",
                       "l2": "inputs (3.4 GB records per tile) far exceed the 126 MB L2; no flush needed"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roof, "collectives": coll,
            "stage_info": {k: v for k, v in info.items() if k != "halo"}, "kernels": kernels,
            "points_this_rank": pts_rank}
    if rank == 0:
        line["cpu_baseline"] = None if args.no_cpu_baseline else cpu_baseline(args, cfg, bounded=True)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
def oracle_step(rec, cfg, args, stage_s=None):
    """One pass of the CPU oracle (numpy + real scikit-learn DBSCAN, n_jobs=-1) over `rec`; per-stage wall seconds
    are added to `stage_s`."""
    from oracle import ground as og, las_io, towers as ot, voxel as ov
    from pointcloudhookup_b200 import synth
    stage_s = stage_s if stage_s is not None else {}

    def lap(name, t0):
        stage_s[name] = stage_s.get(name, 0.0) + time.perf_counter() - t0

    las = {"scales": synth.SCALES, "offsets": synth.OFFSETS, "X": np.ascontiguousarray(rec["X"]),
           "Y": np.ascontiguousarray(rec["Y"]), "Z": np.ascontiguousarray(rec["Z"]), "n": int(rec.size)}
    t0 = time.perf_counter()
    final, _ = ov.downsample_las_arrays(las, cfg["voxel"], cfg["chunk"])
    lap("voxel", t0)
    tiled = args.workload in TILED
    if args.workload == "pipeline" or tiled:
        t0 = time.perf_counter()
        q = [las_io.quantise(final[:, i], synth.SCALES[i], synth.OFFSETS[i]) for i in range(3)]
        las2 = dict(las, X=q[0], Y=q[1], Z=q[2], n=len(q[0]))
        raw, centroid, points = ot.stage_a(las2)
        if tiled or args.ground == "grid":
            mask, _ = og.grid_min_keep_mask(points, 2.0, 3.0)
        else:
            mask, _, _ = og.percentile_keep_mask(points[:, 2])
        filtered = points[mask]
        lap("ground", t0)
        t0 = time.perf_counter()
        if tiled:
            from sklearn.cluster import DBSCAN
            labels = DBSCAN(eps=8.0, min_samples=80, n_jobs=-1, algorithm="ball_tree").fit(filtered).labels_ if len(filtered) else \
                np.zeros(0, np.int32)
        else:
            labels = ot.stage_c(filtered, 8.0, 80)
        lap("dbscan", t0)
        t0 = time.perf_counter()
        k = int(labels.max()) + 1 if labels.size else 0
        n_ok = 0
        order = np.argsort(labels, kind="stable")
        bounds = np.searchsorted(labels[order], np.arange(k + 1))
        for c in range(k):                                     # grouped once: the reference's per-cluster mask scans are O(G*K)
            cp = filtered[order[bounds[c]:bounds[c + 1]]]
            ext, _, _ = ot.cluster_box(cp, "aabb" if (args.box == "aabb" or tiled) else "obb")
            n_ok += int(ext[2] > 15.0)
        lap("boxes", t0)
        out = n_ok
    if args.workload == "voxel_geoid" or cfg.get("geo"):
        from oracle import crs, geoid
        t0 = time.perf_counter()
        x, y, z = las_io.scaled(las)
        lon, lat = crs.gk_inverse(x, y)
        lt = np.linspace(-90, 90, 721)
        ln = -180 + 0.25 * np.arange(1440)
        g = {"ll_lat": -90.0, "ll_lon": -180.0, "dlat": 0.25, "dlon": 0.25, "rows": 721, "cols": 1440,
             "grid": (30 * np.sin(np.radians(lt))[:, None] * np.cos(np.radians(ln))[None, :]).astype(np.float32)}
        out = float(geoid.vgridshift(g, lon, lat, z, -1.0).sum())
        lap("geoid_crs", t0)
    return out


def cpu_baseline(args, cfg, bounded=True, steps=1, warmup=0, sample=None, budget_s=None):
    """The CPU oracle on a bounded prefix of the same synthetic workload, on this box's host cores.  Inside the
    default bench line the sample is sized for ~20-30 s of CPU work; `--impl reference` uses the 5 M-point prefix
    SURVEY.md 8(d) asks for.  budget_s: another pass is only started while the passes so far plus one more of the same
    length stay inside it (at least one pass always runs); the number of passes run is reported as `passes`."""
    from pointcloudhookup_b200 import synth
    default = 1_500_000 if args.workload == "pipeline" else (4_000_000 if args.workload == "voxel_geoid" else 2_000_000)
    sample = int(args.ref_sample or sample or default)
    sample = min(sample, cfg["n"])
    towers = max(1, round(cfg["towers"] * sample / cfg["n"]))
    rec = synth.corridor_records(sample, towers, cfg["terrain"], cfg["seed"])
    for _ in range(warmup):
        oracle_step(rec, cfg, args)
    stage_s = {}
    t0 = time.perf_counter()
    done = 0
    while done < steps:
        oracle_step(rec, cfg, args, stage_s)
        done += 1
        elapsed = time.perf_counter() - t0
        if budget_s is not None and elapsed + elapsed / done > budget_s:
            break
    steps = done
    dt = (time.perf_counter() - t0) / steps
    return {"value": sample / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "points_sampled": sample,
            "extrapolated": True,
            "sample": f"first {sample} points ({towers} tower spans, {cfg['terrain']} terrain) of the same synthetic workload, "
                      f"{dt:.1f} s per pass; numpy single-threaded + scikit-learn DBSCAN n_jobs=-1; the points/s of this "
                      f"prefix is what the ratio against the full-size GPU run extrapolates from",
            "seconds_per_pass": dt, "passes": steps, "stage_seconds_per_pass": {k: v / steps for k, v in stage_s.items()}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workload_config(args)
    steps = max(1, min(args.steps, 2))
    warm = 0                                     # a 5 M-point pass takes minutes; the first pass is as warm as numpy gets
    # a 5 M-point pass takes ~2 min on the 16 host cores of a B200 box: a second one only if the first was quick
    cpu = cpu_baseline(args, cfg, steps=steps, warmup=warm, sample=5_000_000, budget_s=150.0)
    steps = cpu["passes"]
    line = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT,
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": warm,
            "ms_per_step": cpu["seconds_per_pass"] * 1e3, "higher_is_better": True,
            "scaling": "strong" if args.workload in TILED else "weak",
            "vs_baseline": None, "dtype": "f64/f32 (numpy)", "data": "synthetic corridor LAS, bounded sample",
            "config": {"workload": cfg["workload"], "points_sampled": cpu["points_sampled"], "same_config": False,
                       "extrapolated": True, "full_workload_points_per_gpu": cfg["n"], "voxel_size": cfg["voxel"],
                       "chunk_size": cfg["chunk"], "ground": "grid" if args.workload in TILED else args.ground,
                       "box": args.box},
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload in TILED:
        run_tiled(a)
    else:
        run_b200(a)
