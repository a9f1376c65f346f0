"""TEST INFRASTRUCTURE ONLY: runs the compiled reference (oracle/_ref) in a fresh process.
ref_soft keeps all state in globals and its buffers are sized at init, so every
(asset dir, resolution, cvar set) needs its own process."""
from __future__ import annotations

import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(q, basedir, mapname, width, height, specs, cvars, want_z, want_dump, repeat, want_palette=False, overlays=None, clear_each=False):
    try:
        sys.path.insert(0, ROOT)
        from oracle.refsoft import RefSoft
        from tochris_b200.scenes.views import build_refdefs
        import ctypes as C
        r = RefSoft(basedir, width, height, heap_mb=96)
        for k, v in (cvars or {}).items():
            if not isinstance(v, str):
                r.set_cvar(k, v)
        r.load_world(mapname)
        for k, v in (cvars or {}).items():
            if isinstance(v, str):
                r.set_cvar(k, v)            # r_skyname: loads the pictures (needs the file system up)
        cm = C.addressof(C.c_char.in_dll(r.lib, "vid")) + 8     # vid.colormap (client/vid.h:36: after `byte *buffer`)
        rds = build_refdefs(specs, width, height, model_resolver=r.model, colormap_ptr=cm)
        out = {}
        if overlays is not None:
            # one view at a time, the view's Draw_* commands on top of it
            fbs = []
            for i in range(len(specs)):
                r.render(rds[i], False)
                fbs.append(r.draw2d(overlays[i]))
            out["fb"] = np.stack(fbs); out["z"] = None
        elif clear_each:
            # one view at a time into cleared buffers (colour 0, z 0): what a view leaves untouched -- everything but the
            # entities of an RDF_NOWORLDMODEL view, the screen outside a sub-rectangle -- then compares as 0
            fbs, zs = [], []
            for i in range(len(specs)):
                r.clear(0, 0)
                fb, z = r.render(rds[i], True)
                fbs.append(fb); zs.append(z)
            out["fb"] = np.stack(fbs); out["z"] = np.stack(zs)
        elif want_palette:
            # one view at a time, R_EndFrame after each: frame buffer + the palette the reference uploaded
            fbs, pals = [], []
            for i in range(len(specs)):
                fb, _ = r.render(rds[i], False)
                fbs.append(fb)
                pals.append(r.end_frame())
            out["fb"] = np.stack(fbs); out["z"] = None; out["palettes"] = pals
        elif want_dump:
            r.dump_enable(True)
            fbs, zs, dumps = [], [], []
            for i in range(len(specs)):
                fb, z = r.render(rds[i], True)
                fbs.append(fb); zs.append(z)
                e, s, sp = r.dump_get()
                dumps.append((e, s, sp, r.counters()))
            out["fb"] = np.stack(fbs); out["z"] = np.stack(zs); out["dumps"] = dumps
        else:
            secs = 0.0
            for _ in range(max(1, repeat)):
                fb, z, secs = r.render_batch(rds, True, want_z)
            out["fb"] = fb; out["z"] = z; out["seconds"] = secs
        q.put(("ok", out))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put(("err", traceback.format_exc()))


def render(basedir, mapname, width, height, specs, cvars=None, want_z=True, want_dump=False, repeat=1, timeout=600, want_palette=False, overlays=None, clear_each=False):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_worker, args=(q, basedir, mapname, width, height, specs, cvars, want_z, want_dump, repeat, want_palette, overlays, clear_each))
    p.start()
    try:
        # a worker that dies (a crash inside the C library) never answers: notice it instead of waiting out the timeout
        import queue as _queue
        import time as _time
        deadline = _time.monotonic() + timeout
        while True:
            try:
                kind, payload = q.get(timeout=1.0)
                break
            except _queue.Empty:
                if not p.is_alive():
                    try:
                        kind, payload = q.get(timeout=1.0)
                        break
                    except _queue.Empty:
                        kind, payload = "err", f"worker process died (exit code {p.exitcode})"
                        break
                if _time.monotonic() > deadline:
                    kind, payload = "err", f"no answer within {timeout} s"
                    break
    finally:
        p.join(timeout=30)
        if p.is_alive():
            p.kill()
    if kind != "ok":
        raise RuntimeError("oracle failed:\n" + payload)
    return payload
