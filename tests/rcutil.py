"""Shared helpers for the test-suite: synthetic scenes and renderer drivers."""
from __future__ import annotations

import ctypes as C
import hashlib
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tochris_b200.scenes import views, worlds  # noqa: E402

_INFO = {}


def scene_defs():
    return {
        "box": lambda: worlds.box_room(),
        "multi": lambda: worlds.multi_room(4),
        "small": lambda: worlds.multi_room(2, room=256, height=160),
        "full": lambda: worlds.multi_room(3, water=True, sky=True, anim=True, doors=2),
    }


def make_assets(basedir: str, names=None):
    views.write_base_assets(basedir)
    for name, fn in scene_defs().items():
        if names and name not in names:
            continue
        data, info = fn()
        views.write_file(basedir, f"maps/{name}.bsp", data)
        _INFO[name] = info
    return basedir


def scene_info(name: str) -> dict:
    if name not in _INFO:
        _INFO[name] = scene_defs()[name]()[1]
    return _INFO[name]


def scene_views(name: str, n: int, dlights: int = 0, underwater_every: int = 0, **kw):
    if name == "box":
        vs = views.timerefresh_views((0, 0, 0), 128)[:n] if n <= 128 else views.timerefresh_views((0, 0, 0), n)
    else:
        vs = views.halton_views(scene_info(name), n, **kw)
    if dlights:
        add_dlights(vs, dlights)
    if underwater_every:
        for i, sp in enumerate(vs):
            if i % underwater_every == 0:
                sp.flags |= 2          # RDF_UNDERWATER -> r_dowarp (r_misc.c:446)
    return vs


def add_dlights(specs, k: int):
    """k dynamic lights per view, radius 200-300, scattered around the camera."""
    for i, sp in enumerate(specs):
        sp.dlights = []
        for j in range(k):
            h = i * 31 + j * 7 + 1
            ox = (views.halton(h, 2) * 2 - 1) * 260.0
            oy = (views.halton(h, 3) * 2 - 1) * 260.0
            oz = (views.halton(h, 5) * 2 - 1) * 60.0
            sp.dlights.append((sp.origin[0] + ox, sp.origin[1] + oy, sp.origin[2] + oz, 200.0 + 100.0 * views.halton(h, 7)))


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---- ref_cuda (product library or its host-sim build) in a fresh process -----------------
def _rc_worker(q, libpath, basedir, mapname, width, height, specs, cvars, want_debug):
    try:
        from tochris_b200.api import RefCuda
        r = RefCuda(basedir, width, height, libpath=libpath, **(cvars or {}))
        r.load_world(mapname)
        rds = views.build_refdefs(specs, width, height, model_resolver=r.model)
        fb, z, st = r.render(rds)
        out = dict(fb=fb, z=z, status=st, stats=r.stats())
        if want_debug:
            out["edges"] = [r.debug_read(0, i) for i in range(len(specs))]
            out["surfs"] = [r.debug_read(1, i) for i in range(len(specs))]
            out["spans"] = r.debug_read(2)
        r.shutdown()
        q.put(("ok", out))
    except Exception:  # noqa: BLE001
        import traceback
        q.put(("err", traceback.format_exc()))


def render_rc(libpath, basedir, mapname, width, height, specs, cvars=None, want_debug=False, timeout=900):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_rc_worker, args=(q, libpath, basedir, mapname, width, height, specs, cvars, want_debug))
    p.start()
    try:
        kind, payload = q.get(timeout=timeout)
    finally:
        p.join(timeout=30)
        if p.is_alive():
            p.kill()
    if kind != "ok":
        raise RuntimeError("ref_cuda run failed:\n" + payload)
    return payload


def diff_fraction(a: np.ndarray, b: np.ndarray) -> float:
    return float((a != b).sum()) / a.size
