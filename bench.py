#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric for ref_cuda: frames/s of batched ref_soft views.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, libref_cuda.so)
  python bench.py --impl reference --gpus N ...            the reference's own C renderer on the host cores

A step = one pass of the hot path (R_RenderView semantics per view) over one batch of
synthetic views.  N=1 workload = BASELINE.json configs[1]: procedural multi-room BSP with
lightmaps and 4 mip levels, 640x480, batch of 64 views.  N>1: every rank renders its own
64-view shard (weak scaling) and the finished framebuffers are gathered to rank 0 over NCCL
inside the timed region -- the only collective on this path.

value  : frames/s, device-timed (CUDA events on the launching stream), world + outputs in HBM
e2e    : frames/s through RC_RenderViews with HOST buffers (view upload + framebuffer download)
roofline / cpu_baseline: see DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import gc
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOAD = "C2: procedural multi-room BSP (4x4 rooms, lightmaps, 4 mips), 640x480, 64 views per GPU"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--views", type=int, default=64, help="views per GPU per step")
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--scene", default="multi")
    ap.add_argument("--keep-cache", action="store_true", help="do not flush the surface cache between steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=8.0)
    ap.add_argument("--gather", default="peer-copy", choices=["peer-copy", "peer", "nccl", "nccl-groups"], help="N>1: how framebuffers reach rank 0")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra.c4 / extra.v320x240 / extra.c3 sub-records")
    ap.add_argument("--c4-views", type=int, default=65536, help="total views of the extra.c4 record (BASELINE configs[3])")
    ap.add_argument("--v320-views", type=int, default=1024, help="views per GPU of the extra.v320x240 record")
    return ap.parse_args()


# ------------------------------------------------------------------------------ scene ----
def make_scene(args, rank):
    import rcutil
    base = tempfile.mkdtemp(prefix=f"rc_bench_{rank}_")
    rcutil.make_assets(base, names=[args.scene])
    specs = rcutil.scene_views(args.scene, args.views, start=1 + rank * args.views, time=1.0)
    return base, specs


# -------------------------------------------------------------------- reference on CPU ----
def _cpu_worker(idx, nworkers, basedir, mapname, w, h, specs, barrier, seconds, passes, q):
    try:
        sys.path.insert(0, ROOT)
        from oracle.refsoft import RefSoft
        from tochris_b200.scenes.views import build_refdefs
        mine = specs[idx::nworkers]
        r = RefSoft(basedir, w, h, heap_mb=64)
        r.load_world(mapname)
        rds = build_refdefs(mine, w, h, model_resolver=r.model)
        r.render_batch(rds, False, False)            # warm-up pass (surface cache, page-in)
        barrier.wait()
        t0 = time.perf_counter()
        done = 0
        if passes:
            for _ in range(passes):
                r.render_batch(rds, False, False)
                done += 1
        else:
            while time.perf_counter() - t0 < seconds:
                r.render_batch(rds, False, False)
                done += 1
        t1 = time.perf_counter()
        q.put((idx, len(mine) * done, t0, t1))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((idx, -1, 0.0, traceback.format_exc()))


def run_reference_cpu(basedir, mapname, w, h, specs, seconds=8.0, passes=0):
    """The oracle (unmodified ref_soft C path), one engine process per host core over the
    same view batch (views split evenly), each looping R_RenderView over its slice after one
    warm-up pass (pattern of R_TimeRefresh_f, ref_soft/r_misc.c:42-58)."""
    from oracle import refsoft
    if not refsoft.available():
        return None
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        pass
    nworkers = max(1, min(cores, len(specs)))
    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(nworkers)
    q = ctx.Queue()
    procs = [ctx.Process(target=_cpu_worker, args=(i, nworkers, basedir, mapname, w, h, specs, barrier, seconds, passes, q))
             for i in range(nworkers)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for r in res:
        if r[1] < 0:
            raise RuntimeError("reference worker failed:\n" + str(r[3]))
    frames = sum(r[1] for r in res)
    wall = max(r[3] for r in res) - min(r[2] for r in res)
    return dict(frames=frames, seconds=wall, fps=frames / wall, cores=nworkers, host_cores=cores)


# ------------------------------------------------------------------------------ clocks ----
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of one GPU while the timed steps run.  In-process NVML (what nvidia-smi itself asks,
    the recipe's clocks line field for field), initialised BEFORE the clock starts: an `nvidia-smi` child per sample took
    tens of milliseconds to start -- on every rank, inside a timed region of a few milliseconds, with the driver locks
    NVML's start-up takes -- and left one sample.  Falls back to the nvidia-smi child when pynvml is missing."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop = index, [], False
        self.nv = self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nv = pynvml
            self.sample()              # (also pages in what the calls need)
            self.rows.clear()
        except Exception:  # noqa: BLE001
            self.nv = self.h = None

    def sample(self):
        nv, h = self.nv, self.h
        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        try:
            bits = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:  # noqa: BLE001
            bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        act = lambda m: "Active" if bits & m else "Not Active"   # noqa: E731
        self.rows.append([str(self.index), str(sm), str(mx), "", act(nv.nvmlClocksThrottleReasonHwSlowdown), act(nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                          act(nv.nvmlClocksThrottleReasonSwThermalSlowdown), act(nv.nvmlClocksThrottleReasonSwPowerCap)])

    def run(self):
        while not self.stop:
            try:
                if self.nv is not None:
                    self.sample()
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002 if self.nv is not None else 0.1)

    def summary(self):
        sm = sorted(int(r[1]) for r in self.rows if len(r) > 2 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) > 2 and r[2].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------ checks ----
def golden(name):
    import numpy as np
    try:
        return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"), allow_pickle=False)
    except Exception:  # noqa: BLE001
        return None


def sha_dev(t):
    import hashlib
    return hashlib.sha256(t.contiguous().cpu().numpy().tobytes()).hexdigest()


# ------------------------------------------------------------------------------ extras ----
def run_extra(kind, args, rank, world, local):
    """One sub-record of the line: another BASELINE configuration, sharded by view over the ranks, device timed, with the
    gather of the frame buffers to rank 0 inside the timed region (CUDA-IPC mirror) and a check of what arrived.
      c4        BASELINE configs[3]: --c4-views (65,536) independent 160x120 views in total, sharded by shard_bounds
      v320x240  the resolution north_star's "1M 320x240 views/s on 8 GPUs" is quoted on: --v320-views per GPU
      c3        BASELINE configs[2]: 256 alias models + 8192 particles per view, 1280x720, 32 views (one GPU only)"""
    import numpy as np
    import torch
    import torch.distributed as dist
    import rcutil
    from tochris_b200 import shard
    from tochris_b200.api import RefCuda, status_names
    from tochris_b200.scenes.views import build_refdefs
    dev = torch.device("cuda", local)
    if kind == "c4":
        scene, W, H, total, steps, t0, dt, ent = "multi", 160, 120, args.c4_views, 3, 0.0, 0.01, {}
        gname, gviews = "multi_c4_160x120", 256
    elif kind == "v320x240":
        scene, W, H, total, steps, t0, dt, ent = "multi", 320, 240, args.v320_views * world, 5, 0.0, 0.01, {}
        gname, gviews = "multi_320x240", 64
    else:
        scene, W, H, total, steps, t0, dt, ent = "full", 1280, 720, 32, 5, 2.5, 0.5, dict(entities=256, particles=8192)
        gname, gviews = "full_c3_1280x720", 32
    px1 = W * H
    lo, hi = shard.shard_bounds(total, rank, world)
    n = hi - lo
    base = tempfile.mkdtemp(prefix=f"rc_bench_{kind}_{rank}_")
    rcutil.make_assets(base, names=[scene])

    def specs_of(a, b):
        return rcutil.scene_views(scene, b - a, start=1 + a, time=t0 + a * dt, time_step=dt, **ent)
    r = RefCuda(base, W, H, device=local, host_threads=max(1, (os.cpu_count() or 1) // world))
    r.load_world(f"maps/{scene}.bsp")
    rds = build_refdefs(specs_of(lo, hi), W, H, model_resolver=r.model)
    fb = torch.empty((n, H, W), dtype=torch.uint8, device=dev)
    zb = torch.empty((n, H, W), dtype=torch.int16, device=dev)
    root_ptr = 0
    if world > 1:
        hbuf = torch.zeros(64, dtype=torch.uint8, device=dev)
        if rank == 0:
            root_ptr = r.device_alloc(total * px1)
            hbuf.copy_(torch.frombuffer(bytearray(r.ipc_export(root_ptr)), dtype=torch.uint8))
        dist.broadcast(hbuf, src=0)
        if rank != 0:
            root_ptr = r.ipc_import(bytes(hbuf.cpu().numpy().tobytes()))
        r.set_mirror_output(root_ptr + lo * px1)
        r.set_mirror_deferred(True)          # the copy of step k beside the kernels of step k+1; fence before the clock stops
        fb_alt = torch.empty_like(fb)
    l2flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(dev)          # the L2 flush, the events and the render on one explicit stream
    torch.cuda.set_stream(stream)
    status = 0
    stepno = [0]

    def step(last=False):
        r.flush_caches()
        l2flush.zero_()
        stepno[0] += 1
        target = fb_alt if (world > 1 and stepno[0] & 1) else fb
        st = r.render_device(rds, n, target.data_ptr(), zb.data_ptr(), stream.cuda_stream)
        if world > 1 and last:
            r.mirror_fence(stream.cuda_stream)
        return st
    for _ in range(3):
        status |= step(last=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    gc.collect()
    gc.disable()
    t_w0 = time.perf_counter()
    for k, (a, b) in enumerate(evs):
        a.record(stream)
        status |= step(last=(k == steps - 1))
        b.record(stream)
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t_w0) * 1e3
    gc.enable()
    if world > 1:
        dist.barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(t[0]), float(t[1])
    prof = None
    if kind == "c3":
        r.set_profiling(True)
        acc = {}
        for _ in range(steps):
            status |= step()
            for k, v in r.stats().items():
                acc[k] = acc.get(k, 0) + v
        r.set_profiling(False)
        prof = {k[3:]: round(v / steps, 4) for k, v in acc.items() if k.startswith("ms_") and k not in ("ms_h2d", "ms_d2h")}
        frag = {k: int(acc.get(k, 0) / steps) for k in ("alias_fragments", "alias_fragments_front", "alias_spans")}
    # ---- what arrived on rank 0 against (a) the committed golden hash of the first views, (b) rank 0's own render of
    # every other rank's shard (outside the timed region)
    verified, how = None, []
    if world > 1:
        r.set_mirror_deferred(False)
        r.set_mirror_output(0)
        if stepno[0] & 1:                     # the last timed step drew into the alternate buffer
            fb.copy_(fb_alt)
    if rank == 0:
        import ctypes as C
        g = golden(gname)
        if world > 1:
            arrived = torch.empty((total, H, W), dtype=torch.uint8, device=dev)
            C.CDLL("libcudart.so").cudaMemcpy(C.c_void_p(arrived.data_ptr()), C.c_void_p(root_ptr), C.c_size_t(total * px1), 3)
        else:
            arrived = fb
        verified = True
        if g is not None and int(g["n"]) == gviews and total >= gviews and (int(g["w"]), int(g["h"])) == (W, H):
            ok = sha_dev(arrived[:gviews]) == str(g["fb_sha"])
            verified &= ok
            how.append(f"first {gviews} frames == tests/golden/{gname}.npz (oracle): {ok}")
        if world > 1:
            ok = True
            chunk = max(1, min(n, 8192))
            tmp = torch.empty((chunk, H, W), dtype=torch.uint8, device=dev)
            tz = torch.empty((chunk, H, W), dtype=torch.int16, device=dev)
            for a in range(n, total, chunk):
                b = min(total, a + chunk)
                rd2 = build_refdefs(specs_of(a, b), W, H, model_resolver=r.model)
                status |= r.render_device(rd2, b - a, tmp.data_ptr(), tz.data_ptr(), stream.cuda_stream)
                torch.cuda.synchronize()
                ok &= bool(torch.equal(tmp[:b - a], arrived[a:b]))
            ok &= bool(torch.equal(fb, arrived[:n]))
            verified &= ok
            how.append(f"all {total} gathered frames == rank 0's own render of every shard: {ok}")
    if world > 1:
        dist.barrier()
        if rank == 0:
            r.device_free(root_ptr)
        else:
            r.ipc_close(root_ptr)
    r.shutdown()
    del fb, zb, l2flush
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    out = {"workload": {"c4": f"C4: {total} independent 160x120 views of the multi-room BSP, sharded by view over {world} GPU(s)",
                        "v320x240": f"multi-room BSP, 320x240, {total // world} views per GPU",
                        "c3": "C3: 256 alias-model entities + 8192 particles per view, 1280x720, batch 32 (full scene: sky, water, doors)"}[kind],
           "value": round(total * steps / (dev_ms / 1e3), 1), "unit": "views/s" if kind != "c3" else "frames/s", "n_gpus": world,
           "views_total": total, "views_per_gpu": n, "steps": steps, "ms_per_step": round(dev_ms / steps, 4),
           "wall_ms_per_step": round(wall_ms / steps, 4), "mpixel_per_s": round(total * steps * px1 / (dev_ms / 1e3) / 1e6, 1),
           "gather": ("CUDA-IPC mirror into rank 0's buffer inside the timed region (deferred: the copy of step k runs beside step k+1, fence before the clock stops)" if world > 1 else "none (one GPU)"),
           "verified": verified, "verified_how": how, "status": status_names(status),
           "timing": "CUDA events per step on the launching stream, max over ranks; surface cache and L2 flushed before every step"}
    if prof:
        ent_px = None
        out["kernel_ms_per_step"] = prof
        # alias / particle pixels cost 6 B when they pass the z test (2 B z read, 1 B texel, 1 B colour, 2 B z
        # write) and 2 B when they fail; "pass" is counted against the world's z-buffer (what the span kernel tests first)
        model_bytes = 6 * frag["alias_fragments_front"] + 2 * (frag["alias_fragments"] - frag["alias_fragments_front"])
        alias_ms = prof.get("alias") or 0.0
        out["entity_stage"] = {"alias_ms": prof.get("alias"), "particles_ms": prof.get("particles"), "zresolve_ms": prof.get("zresolve"),
                               "alias_spans_per_step": frag["alias_spans"], "alias_fragments_per_step": frag["alias_fragments"],
                               "alias_fragments_in_front_of_world": frag["alias_fragments_front"],
                               "alias_fragments_per_view_pixel": round(frag["alias_fragments"] / float(total * px1), 2),
                               "alias_model_bytes_per_step": int(model_bytes),
                               "alias_model_gbs": round(model_bytes / (alias_ms * 1e-3) / 1e9, 1) if alias_ms else None,
                               "note": "k_alias_verts + k_alias_tris + k_alias_spans + k_sprites | k_particles | k_zresolve, serialised by the profiling events; "
                                       "fragments = pixels of alias spans the rasteriser looked at (overdraw included: the C3 scene stacks 256 models, many of them "
                                       "near the camera, in front of every view), bytes by SURVEY.md 8(d)'s 6 B / 2 B model"}
    return out


# ------------------------------------------------------------------------------- ours ----
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from tochris_b200 import shard
    from tochris_b200.api import RefCuda, status_names
    from tochris_b200.scenes.views import build_refdefs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- ref_cuda has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    base, specs = make_scene(args, rank)
    W, H, n = args.width, args.height, args.views
    r = RefCuda(base, W, H, device=local, host_threads=max(1, (os.cpu_count() or 1) // world))
    r.load_world(f"maps/{args.scene}.bsp")
    rds = build_refdefs(specs, W, H, model_resolver=r.model)
    dev = torch.device("cuda", local)
    fb = torch.empty((n, H, W), dtype=torch.uint8, device=dev)
    zb = torch.empty((n, H, W), dtype=torch.int16, device=dev)
    # N>1: the path's only cross-GPU step is bringing the framebuffers to rank 0.
    #   peer-copy (default): rank 0 exports its gather buffer (CUDA IPC); every rank renders locally and copies each
    #                   finished group of 16 views into its slice over NVLink while the next group is drawn
    #   peer:           every rank renders straight into its slice: the draw kernel's stores cross NVLink
    #   nccl:           render locally, then one NCCL gather to rank 0
    #   nccl-groups:    NCCL gather per group of 16 views on a side stream while the next group is drawn
    side = torch.cuda.Stream(dev) if world > 1 else None
    gbufs = {}
    gather = None
    fb_target = fb.data_ptr()
    px1 = H * W
    if world > 1 and args.gather in ("peer", "peer-copy"):
        hbuf = torch.zeros(64, dtype=torch.uint8, device=dev)
        if rank == 0:
            root_ptr = r.device_alloc(world * n * px1)
            hbuf.copy_(torch.frombuffer(bytearray(r.ipc_export(root_ptr)), dtype=torch.uint8))
        dist.broadcast(hbuf, src=0)
        if rank != 0:
            root_ptr = r.ipc_import(bytes(hbuf.cpu().numpy().tobytes()))
        if args.gather == "peer":
            fb_target = root_ptr + rank * n * px1
        else:
            # deferred mirror: the copy of step k into rank 0's buffer runs beside the kernels of step k+1 (two frame
            # buffers in rotation); RC_MirrorFence before the end of the timed region
            r.set_mirror_output(root_ptr + rank * n * px1)
            r.set_mirror_deferred(True)
            fb_alt = torch.empty_like(fb)
    elif world > 1 and args.gather == "nccl":
        gather = [torch.empty_like(fb) for _ in range(world)] if rank == 0 else None
    elif world > 1:
        def on_group(first, num):
            with torch.cuda.stream(side):
                if rank == 0 and first not in gbufs:
                    gbufs[first] = [torch.empty((num, H, W), dtype=torch.uint8, device=dev) for _ in range(world)]
                dist.gather(fb[first:first + num], gbufs.get(first), dst=0)
        r.set_group_callback(on_group, side.cuda_stream, 4)
    l2flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # an explicit non-default stream carries the L2 flush, the timing events and the render, in that order (a NULL stream
    # means the legacy default stream to RC_RenderViewsDevice, and CUDA graphs cannot be captured there)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    px = n * W * H

    deferred = world > 1 and args.gather == "peer-copy"
    stepno = [0]

    def step(timed_events=None, last=False):
        if not args.keep_cache:
            r.flush_caches()               # D_FlushCaches: every step rebuilds its surface blocks
        l2flush.zero_()                    # evict the previous step's data from the 126 MB L2
        if timed_events is not None:
            timed_events[0].record(stream)
        stepno[0] += 1
        target = (fb_alt.data_ptr() if stepno[0] & 1 else fb.data_ptr()) if deferred else fb_target
        st = r.render_device(rds, n, target, zb.data_ptr(), stream.cuda_stream)
        if deferred and last:
            r.mirror_fence(stream.cuda_stream)     # every frame of every step has arrived on rank 0 before the clock stops
        if world > 1 and args.gather == "nccl":
            shard.gather_frames(fb, n * world, rank, world, dst=0, bufs=gather, concat=False)
        elif world > 1 and args.gather == "nccl-groups":
            stream.wait_stream(side)       # the step ends when the last group has arrived on rank 0
        if timed_events is not None:
            timed_events[1].record(stream)
        return st

    status = 0
    for _ in range(max(args.warmup, 3)):
        status |= step(last=True)
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None       # (rank 0 prints the line: its GPU's clocks; NVML starts before the barrier)
    if world > 1:
        dist.barrier()
    if sampler is not None:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = r.stats()["kernel_launches"]
    launches = 0
    torch.cuda.synchronize()
    # (a step is 0.2 ms of device work and the host enqueues it while the L2 flush runs: a collection of the interpreter's
    # garbage inside the K steps would leave the device idle between a step's two events -- on one of N ranks, which the
    # max over ranks then reports for all)
    gc.collect()
    gc.disable()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        status |= step(evs[k], last=(k == args.steps - 1))
        launches += r.stats()["kernel_launches"]
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    gc.enable()
    if world > 1:
        dist.barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    if sampler is not None:
        sampler.stop = True
    if world > 1 and args.gather == "nccl-groups":
        r.set_group_callback(None)
    gather_check = None
    if world > 1 and args.gather in ("peer", "peer-copy"):
        r.set_mirror_deferred(False)
        r.set_mirror_output(0)
        # once, outside the timed region: what arrived on rank 0 over NVLink equals what each rank renders locally
        r.render_device(rds, n, fb.data_ptr(), zb.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        wts = torch.arange(1, n * px1 + 1, device=dev, dtype=torch.int64) % 65521
        mine = torch.stack([(fb.view(-1).to(torch.int64) * wts).sum(), fb.view(-1).to(torch.int64).sum()])
        allsums = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allsums, mine)
        if rank == 0:
            import ctypes as C
            arrived = torch.empty(world * n * px1, dtype=torch.uint8, device=dev)
            C.CDLL("libcudart.so").cudaMemcpy(C.c_void_p(arrived.data_ptr()), C.c_void_p(root_ptr), C.c_size_t(world * n * px1), 3)
            gather_check = True
            for q in range(world):
                sl = arrived[q * n * px1:(q + 1) * n * px1].to(torch.int64)
                gather_check &= bool(((sl * wts).sum() == allsums[q][0]) and (sl.sum() == allsums[q][1]))

    # ---- e2e: the reference-facing C-ABI calls with HOST buffers ----
    # (1) blocking RC_RenderViews, one call per step, timed call by call
    fb_host = torch.empty((n, H, W), dtype=torch.uint8, pin_memory=True)
    fb_np = fb_host.numpy()
    e2e_sync_s = 0.0
    for k in range(3 + args.steps):
        if not args.keep_cache:
            r.flush_caches()
        l2flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, _, st = r.render(rds, n, want_z=False, fb=fb_np)
        dt = time.perf_counter() - t0
        status |= st
        if k >= 3:
            e2e_sync_s += dt
    ref_sum = int(fb_np.astype(np.uint64).sum())
    # (2) the pipelined form a stream of batches uses (RC_RenderViewsAsync / RC_WaitViews, three batches outstanding, three
    # pinned output buffers): every step still uploads its refdefs, flushes the surface cache and L2, renders and brings
    # its frame buffers to the host inside the timed region; the copy of step k overlaps the kernels of step k+1.
    fb_host2 = [torch.empty((n, H, W), dtype=torch.uint8, pin_memory=True) for _ in range(3)]
    fb_np2 = [t.numpy() for t in fb_host2]
    e2e_ok = True

    def pipelined(steps):
        nonlocal status
        for k in range(steps):
            if k >= 3:
                st = r.wait()
                status |= st
            if not args.keep_cache:
                r.flush_caches()
            l2flush.zero_()
            r.render_async(rds, n, fb_np2[k % 3])
        for k in range(min(3, steps)):
            status |= r.wait()
    pipelined(3)
    torch.cuda.synchronize()
    gc.collect()
    gc.disable()
    t0 = time.perf_counter()
    pipelined(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    gc.enable()
    e2e_ok &= all(int(b.astype(np.uint64).sum()) == ref_sum for b in fb_np2)
    # ---- the PCIe floor under e2e: the same bytes, device plane -> the same pinned buffer, nothing else running
    cevs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
    for a, b in cevs:
        a.record(); fb_host.copy_(fb.view(n, H, W), non_blocking=True); b.record()
    torch.cuda.synchronize()
    d2h_ms = sorted(a.elapsed_time(b) for a, b in cevs[2:])[len(cevs[2:]) // 2]
    # N>1: the same copy by ALL ranks at once -- what the host takes in when every GPU delivers its frame buffers at the same
    # time (one VM, one NUMA node: profiles/r02o_d2h_probe_8gpu.json); the floor of an e2e step at N GPUs
    d2h_all_ms = None
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        t0c = time.perf_counter()
        for _ in range(8):
            fb_host.copy_(fb.view(n, H, W), non_blocking=True)
        torch.cuda.synchronize()
        tc = torch.tensor([(time.perf_counter() - t0c) / 8 * 1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        d2h_all_ms = float(tc[0])
    # ---- extra: the same steps with the surface cache KEPT across steps, which is how ref_soft (and the CPU arm
    # below, after its warm-up pass) normally runs: D_CacheSurface hits.  Reported beside `value`, not as it.
    warm_ms = None
    if world == 1:
        wevs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        r.render_device(rds, n, fb.data_ptr(), zb.data_ptr(), stream.cuda_stream)
        for k in range(args.steps):
            l2flush.zero_()
            wevs[k][0].record(stream)
            r.render_device(rds, n, fb.data_ptr(), zb.data_ptr(), stream.cuda_stream)
            wevs[k][1].record(stream)
        torch.cuda.synchronize()
        warm_ms = sum(a.elapsed_time(b) for a, b in wevs) / args.steps
    # ---- extra: the end of the frame for the batch (RC_PresentViewsDevice: every view's 8-bit plane -> 32-bit display
    # pixels through its own R_UpdatePalette palette).  Streaming kernel: 1 B read + 4 B written per pixel.
    present = None
    if world == 1:
        rgbx = torch.empty((n, H, W), dtype=torch.int32, device=dev)
        pevs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        pstream = torch.cuda.Stream(dev)      # the call is asynchronous: time it on the stream it is given
        torch.cuda.synchronize()
        with torch.cuda.stream(pstream):
            for k in range(3 + args.steps):
                l2flush.zero_()
                if k >= 3:
                    pevs[k - 3][0].record(pstream)
                r.present_device(rds, n, fb.data_ptr(), rgbx.data_ptr(), pstream.cuda_stream)
                if k >= 3:
                    pevs[k - 3][1].record(pstream)
            torch.cuda.synchronize()
            kevs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            for k in range(args.steps):       # the kernel alone: palettes of the previous call, nothing uploaded
                l2flush.zero_()
                kevs[k][0].record(pstream)
                r.present_device(None, n, fb.data_ptr(), rgbx.data_ptr(), pstream.cuda_stream)
                kevs[k][1].record(pstream)
            # eight launches between one pair of events: the figure of a kernel timed alone carries its launch latency
            # (a third of it for a 17 us kernel)
            l2flush.zero_()
            b2b = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            b2b[0].record(pstream)
            for _ in range(8):
                r.present_device(None, n, fb.data_ptr(), rgbx.data_ptr(), pstream.cuda_stream)
            b2b[1].record(pstream)
        torch.cuda.synchronize()
        bms = b2b[0].elapsed_time(b2b[1]) / 8.0
        pms = sum(a.elapsed_time(b) for a, b in pevs) / args.steps
        kms = sum(a.elapsed_time(b) for a, b in kevs) / args.steps
        present = {"kernel": "k_present (palette expansion, R_EndFrame for batches)", "ms_per_step": round(kms, 4),
                   "alg_bytes_per_launch": int(5 * px), "achieved_gbs": round(5 * px / (kms * 1e-3) / 1e9, 1),
                   "call_ms_with_palette_upload": round(pms, 4),
                   "back_to_back": {"ms_per_launch": round(bms, 4), "achieved_gbs": round(5 * px / (bms * 1e-3) / 1e9, 1),
                                    "note": "eight launches between one pair of events (98 MB per launch, L2 flushed before the first)"},
                   "note": "1 B read + 4 B written per pixel; kernel alone (palettes already on the device), and the whole call with its 64 KB palette upload"}
    # ---- per-kernel times (profiling mode puts events between the kernels) ----
    r.set_profiling(True)
    prof = {}
    for k in range(args.steps):
        if not args.keep_cache:
            r.flush_caches()
        l2flush.zero_()
        r.render_device(rds, n, fb.data_ptr(), zb.data_ptr(), stream.cuda_stream)
        s = r.stats()
        for key, v in s.items():
            prof[key] = prof.get(key, 0) + v
    r.set_profiling(False)
    torch.cuda.synchronize()

    t = torch.tensor([dev_ms, e2e_s, e2e_sync_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s, e2e_sync_s = float(t[0]), float(t[1]), float(t[2])
    # the timed `value` frames against the committed golden vectors of the same batch (tests/golden/multi_c2_640x480.npz,
    # made by the oracle): rank 0 renders views 1..64, which is that batch
    value_check = None
    if rank == 0 and (args.scene, W, H, n) == ("multi", 640, 480, 64):
        g = golden("multi_c2_640x480")
        if g is not None:
            r.render_device(rds, n, fb.data_ptr(), zb.data_ptr(), stream.cuda_stream)
            torch.cuda.synchronize()
            value_check = {"fb_sha_matches_golden": sha_dev(fb) == str(g["fb_sha"]), "z_sha_matches_golden": sha_dev(zb) == str(g["z_sha"]),
                           "golden": "tests/golden/multi_c2_640x480.npz (the oracle's frames of the same 64 views)"}
    # ---- the other BASELINE configurations as sub-records (every rank takes part: the views are sharded)
    r.shutdown()
    del fb, zb, l2flush
    torch.cuda.empty_cache()
    extra = {}
    if not args.no_extras:
        for kind in ("v320x240", "c4", "c3"):
            if kind == "c3" and world > 1:
                continue
            try:
                extra[kind] = run_extra(kind, args, rank, world, local)
            except Exception as e:  # noqa: BLE001
                import traceback
                extra[kind] = {"error": str(e), "trace": traceback.format_exc()[-600:]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    frames = n * world * args.steps
    ms_per_step = dev_ms / args.steps
    value = frames / (dev_ms / 1e3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # the dominant kernel: k_draw alone (events around it in profiling mode).  It draws the screen cells one span covers
    # completely with its straight-line code; cells that hold a span boundary (k_ends) and the turbulent / sky / solid
    # cells it queues (k_slowcells) are other launches, so its algorithmic bytes are those of its own pixels.
    draw_ms = prof["ms_cells"] / args.steps
    stage_ms = (prof["ms_draw"] - prof.get("ms_spanseg", 0.0)) / args.steps
    kdraw_px = 8.0 * (prof["fullcells"] - prof["slowcells"]) / args.steps
    alg_bytes = 4.0 * kdraw_px                # 1 B texel + 1 B colour + 2 B z per pixel
    achieved = alg_bytes / (draw_ms * 1e-3) / 1e9 if draw_ms > 0 else 0.0
    traffic, traffic_src, ncu_us = None, None, 0.0   # DRAM bytes per launch of k_draw (and its duration) from the newest committed ncu --set full capture
    try:
        pdir = os.path.join(ROOT, "profiles")
        for fn in sorted(f for f in os.listdir(pdir) if f.endswith("_ncu_summary.json")):
            for l in json.load(open(os.path.join(pdir, fn)))["launches"]:
                if "k_draw" in l.get("kernel", "") and (W, H, n) == (640, 480, 64) and l.get("grid", 0) >= 2368:
                    traffic = int((l["dram_read_mb"] + l["dram_write_mb"]) * 1e6)
                    traffic_src = f"profiles/{fn}"
                    ncu_us = float(l.get("time_us", 0.0))
    except Exception:  # noqa: BLE001
        pass
    out = {
        "metric": "frames/s, batched ref_soft views (pixel-exact 8-bit framebuffer + int16 z-buffer)",
        "value": round(value, 1), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int32 fixed-point + f32 setup (fmad off)", "data": "synthetic",
        "config": {"workload": WORKLOAD if (args.scene, W, H, n) == ("multi", 640, 480, 64) else f"{args.scene} {W}x{H} x{n} views per GPU",
                   "width": W, "height": H, "views_per_gpu": n, "global_batch": n * world},
        "method": {"parallelism": f"dp{world} by view", "l2": "flushed between steps (256 MiB write)",
                   "surface_cache": "kept across steps" if args.keep_cache else "flushed before every step (all visible blocks rebuilt inside the timed region)",
                   "gather_verified": gather_check, "collective": ({"peer-copy": "none: rank 0 exports its gather buffer (CUDA IPC peer memory); every rank copies each finished batch into its slice over NVLink on a copy stream while the next step is rendered (deferred mirror, two frame buffers in rotation), RC_MirrorFence before the end of the timed region; NCCL only for the handle broadcast and the closing barrier", "peer": "none: every rank renders straight into rank 0's gather buffer (CUDA IPC peer memory), the draw kernel's stores cross NVLink as they are produced; NCCL only for the handle broadcast and the closing barrier", "nccl": "NCCL gather of framebuffers to rank 0 inside the timed region", "nccl-groups": "NCCL gather per group of 16 views, overlapped with the draw of the next group"}[args.gather] if world > 1 else "none")},
        "mpixel_per_s": round(value * W * H / 1e6, 1),
        "warm_surface_cache": ({"value": round(n / (warm_ms / 1e3), 1), "ms_per_step": round(warm_ms, 4),
                                "note": "same steps without the per-step D_FlushCaches (the reference's and the CPU arm's steady state); informational"}
                               if warm_ms else None),
        "present": (dict(present, frac=round(present["achieved_gbs"] / peak, 4)) if present else None),
        "wall_ms_per_step": round(t_wall / args.steps * 1e3, 4),
        "e2e": {"value": round(n * world * args.steps / min(e2e_s, e2e_sync_s), 1), "unit": "frames/s",
                "h2d_bytes_per_step": int(n * 648), "d2h_bytes_per_step": int(px), "ms_per_step": round(min(e2e_s, e2e_sync_s) / args.steps * 1e3, 4),
                "api": "RC_RenderViewsAsync + RC_WaitViews" if e2e_s <= e2e_sync_s else "RC_RenderViews",
                "pipelined_call": {"value": round(n * world * args.steps / e2e_s, 1), "ms_per_step": round(e2e_s / args.steps * 1e3, 4)},
                "note": "COLOUR PLANES ONLY: the 8-bit frame buffers come back to the host, the int16 z-buffers stay on the device (north_star gathers frame buffers; with z the PCIe bytes would triple).  refdef_t[] in host memory -> 8-bit framebuffers in pinned host memory; every step uploads, flushes the surface cache and L2, renders and downloads inside the timed region.  Both host APIs are measured, the faster one is `value`: the blocking RC_RenderViews (timed call by call) and the pipelined RC_RenderViewsAsync / RC_WaitViews (up to three batches outstanding, the PCIe copy of step k beside the kernels of step k+1; timed over all steps).  The PCIe copy of the frame buffers bounds both (d2h_copy_alone is that copy measured by itself)",
                "verified": bool(e2e_ok),
                "d2h_copy_alone": {"ms": round(d2h_ms, 4), "gbs": round(px / d2h_ms / 1e6, 1),
                                   "note": "median of 6 plain device->pinned-host copies of one step's frame buffers, measured here with nothing else running: the floor of any e2e step on this box"},
                "d2h_copy_all_ranks_at_once": ({"ms": round(d2h_all_ms, 4), "aggregate_gbs": round(world * px / d2h_all_ms / 1e6, 1),
                                                "floor_frames_per_s": round(n * world / (d2h_all_ms / 1e3), 1),
                                                "note": "every rank copying one step's frame buffers to its own pinned buffer at the same time: the host's aggregate ingest (one VM, one NUMA node) bounds e2e at N GPUs, not the GPUs or their PCIe links"}
                                               if d2h_all_ms else None),
                "blocking_call": {"value": round(n * world * args.steps / e2e_sync_s, 1), "ms_per_step": round(e2e_sync_s / args.steps * 1e3, 4),
                                  "note": "RC_RenderViews, one blocking call per step, timed call by call"}},
        "value_verified": value_check, "gather_verified": gather_check,
        "extra": extra,
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "status": status_names(status),
        "roofline": {"bound": "hbm", "kernel": "k_draw (span texture mapping + z-buffer of the screen cells one span covers)", "achieved": round(achieved, 1), "peak": peak,
                     "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                     "alg_bytes_per_launch": int(alg_bytes), "kernel_ms": round(draw_ms, 4),
                     "pixels_per_launch": int(kdraw_px), "pixel_share": round(kdraw_px / px if px else 0.0, 4),
                     "ncu_capture": ({"kernel_us": ncu_us, "frac": round(alg_bytes / (ncu_us * 1e-6) / 1e9 / peak, 4), "source": traffic_src,
                                      "note": "gpu__time_duration of the same launch in the committed ncu capture (cold cache, serialised, no launch gap): the event figure above "
                                              "carries the launch latency of a kernel timed alone; `frac` is the event figure's"} if ncu_us else None),
                     "draw_stage_ms": round(stage_ms, 4),
                     "draw_stage_note": "k_draw + k_ends (cells that hold a span boundary) + k_slowcells (turbulent / sky / solid cells), serialised by the profiling events; they overlap on two streams in the timed steps",
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s"},
        "surface_cache_kernel": {"kernel": "k_cache (R_BuildLightMap + R_DrawSurfaceBlock8_mip0..3 of every block the step's views need; the cache is flushed before each step)",
                                 "texels_per_launch": int(prof["cache_texels"] / args.steps), "alg_bytes_per_launch": int(2 * prof["cache_texels"] / args.steps),
                                 "kernel_ms": round(prof["ms_cache"] / args.steps, 4),
                                 "achieved_gbs": round(2 * prof["cache_texels"] / (prof["ms_cache"] * 1e-3) / 1e9, 1) if prof["ms_cache"] else None,
                                 "frac": round(2 * prof["cache_texels"] / (prof["ms_cache"] * 1e-3) / 1e9 / peak, 4) if prof["ms_cache"] else None,
                                 "note": "2 B per rebuilt texel (SURVEY.md 8d); event-timed alone; in the timed steps it runs beside k_spanseg on a second stream. "
                                         "13 MB of algorithmic bytes per launch: the kernel is far too short to approach the HBM roofline -- an empty launch "
                                         "of it measures 5 us, and its work is chains of dependent loads (DESIGN.md section 3)"},
        "kernel_ms_per_step": {k[3:]: round(v / args.steps, 4) for k, v in prof.items() if k.startswith("ms_") and k not in ("ms_h2d", "ms_d2h")},
        "per_step_counts": {k: int(prof[k] / args.steps) for k in ("faces", "edges", "spans", "blocks", "cache_jobs", "cache_texels")},
    }
    if world == 1 and not args.no_cpu_baseline:
        cb = run_reference_cpu(base, f"maps/{args.scene}.bsp", W, H, specs, seconds=args.cpu_seconds)
        if cb:
            out["cpu_baseline"] = {"value": round(cb["fps"], 1), "unit": "frames/s", "cores": cb["cores"], "kind": "reference",
                                   "sample": f"the same {n}-view batch split over {cb['cores']} engine processes (host has {cb['host_cores']} cores), looped for {cb['seconds']:.1f} s after a warm-up pass; oracle/_ref = unmodified ref_soft C path, gcc -O2"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """The reference arm: the UNMODIFIED reference renderer (oracle/_ref, built from /root/reference by oracle/Makefile) on
    the box's host cores, one engine process per core, over the same views our arm renders at this N (64 views per GPU:
    rank q's views are start = 1 + 64 q), so `config` is identical in both arms.  Rank 0 alone runs it.  A step is
    `passes_per_step` passes over that batch, sized in the warm-up so that a step lasts >= 0.15 s: the K timed steps
    then take seconds, not the tens of milliseconds a single pass over 64 views would (freshly spawned processes, other
    ranks of the launcher still exiting)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import rcutil
    W, H, n = args.width, args.height, args.views
    base = tempfile.mkdtemp(prefix="rc_bench_ref_")
    rcutil.make_assets(base, names=[args.scene])
    specs = []
    for q in range(world):
        specs += rcutil.scene_views(args.scene, n, start=1 + q * n, time=1.0)
    from oracle import refsoft
    if not refsoft.available():
        if os.path.isdir("/root/reference"):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
        else:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/librefsoft.so missing and /root/reference absent"}))
            return
    if world > 1:
        time.sleep(3.0)            # the launcher's other ranks exit at once; let their start-up noise pass
    warm = run_reference_cpu(base, f"maps/{args.scene}.bsp", W, H, specs, passes=max(args.warmup, 1))
    pass_s = warm["seconds"] / max(args.warmup, 1)
    ppstep = max(1, int(0.15 / pass_s + 0.999))
    cb = run_reference_cpu(base, f"maps/{args.scene}.bsp", W, H, specs, passes=args.steps * ppstep)
    value = cb["fps"]
    out = {
        "impl": "reference",
        "metric": "frames/s, batched ref_soft views (pixel-exact 8-bit framebuffer + int16 z-buffer)",
        "value": round(value, 1), "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(cb["seconds"] / args.steps * 1e3, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int32 fixed-point + f32 setup", "data": "synthetic",
        "config": {"workload": WORKLOAD if (args.scene, W, H, n) == ("multi", 640, 480, 64) else f"{args.scene} {W}x{H} x{n} views per GPU",
                   "width": W, "height": H, "views_per_gpu": n, "global_batch": n * world},
        "method": {"parallelism": f"{cb['cores']} host processes by view (the whole host, whatever N is: the reference has no GPU path)",
                   "passes_per_step": ppstep, "timed_seconds": round(cb["seconds"], 3),
                   "surface_cache": "warm (the reference's steady state: D_CacheSurface hits after the warm-up pass)"},
        "mpixel_per_s": round(value * W * H / 1e6, 1),
        "cpu_baseline": {"value": round(value, 1), "unit": "frames/s", "cores": cb["cores"], "kind": "reference",
                         "sample": f"{args.steps} steps x {ppstep} passes over the {n * world}-view batch ({n} views per GPU of our arm) split over {cb['cores']} engine processes (host has {cb['host_cores']} cores), {cb['seconds']:.2f} s timed; oracle/_ref = unmodified ref_soft C path, gcc -O2"},
        "e2e": {"value": round(value, 1), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


if __name__ == "__main__":
    a = parse()
    # the contract is ONE JSON line on stdout: libraries that print there while they initialise (NCCL announces its
    # version on stdout on some boxes) are sent to stderr, and only the result line goes to the real stdout
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    _buf = []
    _print = print

    def print(*args, **kw):  # noqa: A001
        _buf.append(" ".join(str(x) for x in args))
    try:
        if a.impl == "reference":
            run_reference(a)
        else:
            run_ours(a)
    finally:
        sys.stdout.flush()
        for line in _buf:
            os.write(_real_stdout, (line + "\n").encode())
