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
    from tochris_b200.scenes import mdlgen
    from tochris_b200.scenes import assets as assets_mod
    views.write_base_assets(basedir)
    for suf, pcx in assets_mod.skybox_pcx("testsky").items():
        views.write_file(basedir, f"env/testsky{suf}.pcx", pcx)
    from tochris_b200.scenes import sprgen
    for k in range(5):        # one sprite model per orientation type (spritegn.h:82-86)
        views.write_file(basedir, f"progs/spr{k}.spr", sprgen.sprite(sprtype=k, width=32 + 8 * k, height=24 + 8 * k, nframes=2,
                                                                        group=(k % 2 == 0), beamlength=(6.0 if k == 1 else 0.0),
                                                                        synctype=(1 if k == 2 else 0), seed=3 + k))
    views.write_file(basedir, "progs/ball.mdl", mdlgen.icosphere_mdl())
    views.write_file(basedir, "progs/blob.mdl", mdlgen.icosphere_mdl(seed=77, subdiv=1, radius=14.0, nframes=2,
                                                                       frame_group=False, skin_group=True, skin_w=64, skin_h=96))
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


def scene_views(name: str, n: int, dlights: int = 0, underwater_every: int = 0, entities: int = 0, particles: int = 0,
                doors: int = 0, sprites: int = 0, **kw):
    if name == "box":
        vs = views.timerefresh_views((0, 0, 0), 128)[:n] if n <= 128 else views.timerefresh_views((0, 0, 0), n)
    else:
        vs = views.halton_views(scene_info(name), n, **kw)
    if dlights:
        add_dlights(vs, dlights)
    if doors:
        add_doors(vs, scene_info(name), doors)
    if entities:
        add_entities(vs, entities)
    if sprites:
        add_sprites(vs, sprites)
    if particles:
        add_particles(vs, particles)
    if underwater_every:
        for i, sp in enumerate(vs):
            if i % underwater_every == 0:
                sp.flags |= 2          # RDF_UNDERWATER -> r_dowarp (r_misc.c:446)
    return vs


def add_doors(specs, info, mode: int = 1):
    """The map's door brush models ('*1', '*2', ...) as entities: mode 1 = in their doorways (clipped
    to the world BSP), mode 2 = additionally slid / rotated per view, plus one floating in a room."""
    import math
    for i, sp in enumerate(specs):
        ents = []
        for d, org in enumerate(info.get("door_origins", [])):
            o = list(org)
            ang = (0.0, 0.0, 0.0)
            if mode >= 2:
                o[2] += 40.0 * views.halton(i * 7 + d + 1, 3)
                if (i + d) % 3 == 1:
                    ang = (0.0, 30.0 * views.halton(i + d + 1, 5), 0.0)
                if (i + d) % 4 == 2:
                    # carried into the middle of the camera's room: lies in a single leaf
                    o = [sp.origin[0] + 90.0 * math.cos(math.radians(sp.angles[1])),
                         sp.origin[1] + 90.0 * math.sin(math.radians(sp.angles[1])), 30.0]
            ents.append(views.EntitySpec(f"*{d + 1}", tuple(o), ang, frame=(i + d) % 2))
        sp.entities = ents + list(sp.entities)


def add_entities(specs, k: int, near: float = 40.0):
    """k alias entities per view on a jittered ring in front of / around the camera."""
    import math
    for i, sp in enumerate(specs):
        sp.entities = list(sp.entities)
        yaw = math.radians(sp.angles[1])
        for j in range(k):
            h = i * 53 + j * 11 + 3
            ang = yaw + (views.halton(h, 2) - 0.5) * 2.2
            dist = near + views.halton(h, 3) * 260.0
            ox = sp.origin[0] + math.cos(ang) * dist
            oy = sp.origin[1] + math.sin(ang) * dist
            oz = sp.origin[2] - 20.0 + views.halton(h, 5) * 60.0
            model = "progs/ball.mdl" if j % 3 else "progs/blob.mdl"
            nfr = 4 if model.endswith("ball.mdl") else 2
            sp.entities.append(views.EntitySpec(model, (ox, oy, oz), (views.halton(h, 7) * 30 - 15, views.halton(h, 11) * 360, 0.0),
                                                frame=j % nfr, oldframe=(j + 1) % nfr, framelerp=views.halton(h, 13),
                                                flags=(8 if j % 5 == 0 else 0) | (16 if j % 7 == 3 else 0)))


def add_sprites(specs, k: int, near: float = 30.0):
    """k sprite entities per view (all five orientation types, single frames and groups), interleaved with the
    alias entities already in the list so that the shared draw order matters."""
    import math
    for i, sp in enumerate(specs):
        ents = list(sp.entities)
        yaw = math.radians(sp.angles[1])
        for j in range(k):
            h = i * 61 + j * 17 + 7
            ang = yaw + (views.halton(h, 2) - 0.5) * 1.6
            dist = near + views.halton(h, 3) * 220.0
            org = (sp.origin[0] + math.cos(ang) * dist, sp.origin[1] + math.sin(ang) * dist,
                   sp.origin[2] - 16.0 + views.halton(h, 5) * 48.0)
            e = views.EntitySpec(f"progs/spr{(i + j) % 5}.spr", org,
                                 (views.halton(h, 7) * 40 - 20, views.halton(h, 11) * 360, views.halton(h, 13) * 90 - 45),
                                 frame=j % 2, syncbase=views.halton(h, 17))
            ents.insert((j * 3) % (len(ents) + 1), e)
        sp.entities = ents


def add_particles(specs, k: int):
    """k particles per view, uniform in a cone in front of the camera."""
    import math
    for i, sp in enumerate(specs):
        sp.particles = []
        yaw = math.radians(sp.angles[1])
        for j in range(k):
            h = i * 97 + j + 5
            ang = yaw + (views.halton(h, 2) - 0.5) * 1.4
            dist = 12.0 + views.halton(h, 3) * 300.0
            sp.particles.append((sp.origin[0] + math.cos(ang) * dist, sp.origin[1] + math.sin(ang) * dist,
                                 sp.origin[2] + (views.halton(h, 5) - 0.5) * 0.8 * dist, int(views.halton(h, 7) * 255)))


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
        cvars = dict(cvars or {})
        skyname = cvars.pop("r_skyname", None)
        demo_batch = cvars.pop("_timedemo_batch", 0)
        r = RefCuda(basedir, width, height, libpath=libpath, **cvars)
        r.load_world(mapname)
        if skyname:
            r.set_skybox(skyname)
        rds = views.build_refdefs(specs, width, height, model_resolver=r.model)
        if demo_batch:
            # the frames go through the client's scene lists (RC_Scene*) and are rendered by RC_TimeDemo
            scene, frames, nfr, dropped = r.record_scene(rds)
            fb = np.zeros((nfr, height, width), dtype=np.uint8)
            td = r.timedemo(frames, nfr, demo_batch, fb)
            td_dev = r.timedemo(frames, nfr, demo_batch, None)       # frames stay on the device
            r.scene_free(scene)
            out = dict(fb=fb, z=None, status=td["status"], stats=r.stats(), timedemo=td, timedemo_device=td_dev, dropped=dropped)
            r.shutdown()
            q.put(("ok", out))
            return
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
        raise RuntimeError("ref_cuda run failed:\n" + payload)
    return payload


def diff_fraction(a: np.ndarray, b: np.ndarray) -> float:
    return float((a != b).sum()) / a.size


RC_STAT_ALIAS_RANGE = 512
TOLERANCE = 1e-4        # BASELINE.json north_star: differing pixels per frame


def assert_parity(o, r, specs, w, h, allow_alias_range=True):
    """Bit-exact colour and z.  Two documented exceptions:
    - views rendered under water only define the 320x200 warp rectangle of the z-buffer
      (r_misc.c:448-454); the rest is whatever the previous frame left there in the reference;
    - when the alias rasteriser reports RC_STAT_ALIAS_RANGE the reference itself read outside the
      skin / colormap (clipped triangles stepping one texel row off the skin, d_polyse.c:385-447);
      those pixels are undefined in the reference, so the frame is held to north_star's 1e-4.
      The number of such events (rc_stats_t.alias_range_events) and of differing pixels is printed and returned;
      allow_alias_range=False makes the bit a failure (the cases that pin "it does not fire")."""
    from tochris_b200 import api
    status = r["status"] & ~RC_STAT_ALIAS_RANGE
    assert status == 0, api.status_names(r["status"])
    loose = bool(r["status"] & RC_STAT_ALIAS_RANGE)
    events = int(r.get("stats", {}).get("alias_range_events", -1))
    assert allow_alias_range or not loose, f"RC_STAT_ALIAS_RANGE fired ({events} events)"
    worst = (0, 0)
    total = [0, 0]
    for i, sp in enumerate(specs):
        hh, ww = (min(h, 200), min(w, 320)) if sp.flags & 2 else (h, w)
        dz = int((o["z"][i, :hh, :ww] != r["z"][i, :hh, :ww]).sum())
        dfb = int((o["fb"][i] != r["fb"][i]).sum())
        total[0] += dfb; total[1] += dz
        worst = max(worst, (dfb, dz))
        if loose:
            assert dfb <= TOLERANCE * w * h and dz <= TOLERANCE * w * h, (i, dfb, dz)
        else:
            assert dfb == 0 and dz == 0, (i, dfb, dz)
    info = dict(alias_range=loose, alias_range_events=events, diff_fb=total[0], diff_z=total[1], worst_view=worst)
    if loose:
        print(f"RC_STAT_ALIAS_RANGE: {events} events; differing pixels over {len(specs)} views: colour {total[0]}, z {total[1]} "
              f"(worst view {worst}, tolerance {TOLERANCE * w * h:.0f} per view)")
    return info
