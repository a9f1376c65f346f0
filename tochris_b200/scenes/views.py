"""Camera sets for the BASELINE.json configs, asset-directory writer and
the picklable view description both renderers are driven from."""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

from . import assets
from ..refdef import (refdef_t, entity_t, dlight_t, particle_t, cshift_t, default_lightstyles)


@dataclass
class EntitySpec:
    model: str
    origin: Tuple[float, float, float]
    angles: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    frame: int = 0
    oldframe: int = 0
    framelerp: float = 0.0
    skinnum: int = 0
    flags: int = 0
    syncbase: float = 0.0


@dataclass
class ViewSpec:
    origin: Tuple[float, float, float]
    angles: Tuple[float, float, float]
    time: float = 1.0
    fov_x: float = 90.0
    flags: int = 0
    rect: Optional[Tuple[int, int, int, int]] = None     # x, y, w, h (default: full screen)
    entities: List[EntitySpec] = field(default_factory=list)
    dlights: List[Tuple[float, float, float, float]] = field(default_factory=list)   # x, y, z, radius
    particles: List[Tuple[float, float, float, int]] = field(default_factory=list)   # x, y, z, colour
    cshifts: List[Tuple[int, int, int, int]] = field(default_factory=list)           # r, g, b, percent (ref.h:49-53)
    fov_y: Optional[float] = None                        # default: CalcFov (fov_x, w, h); a replayed frame brings its own


def write_base_assets(basedir: str):
    """<basedir>/id1/gfx/{palette,colormap}.lmp  (common/host.c:827-832 load these)."""
    gfx = os.path.join(basedir, "id1", "gfx")
    os.makedirs(gfx, exist_ok=True)
    os.makedirs(os.path.join(basedir, "id1", "maps"), exist_ok=True)
    os.makedirs(os.path.join(basedir, "id1", "progs"), exist_ok=True)
    pal = assets.make_palette()
    with open(os.path.join(gfx, "palette.lmp"), "wb") as f:
        f.write(pal.tobytes())
    with open(os.path.join(gfx, "colormap.lmp"), "wb") as f:
        f.write(assets.make_colormap(pal))
    # 2D assets: host.c:815 loads gfx.wad, Draw_ConsoleBackground / the menus load qpic lumps with Draw_CachePic
    with open(os.path.join(basedir, "id1", "gfx.wad"), "wb") as f:
        f.write(assets.gfx_wad())
    with open(os.path.join(gfx, "conback.lmp"), "wb") as f:
        f.write(assets.make_qpic(320, 200, 21))
    with open(os.path.join(gfx, "menuplyr.lmp"), "wb") as f:
        f.write(assets.make_qpic(48, 56, 22, transparent=True))


def write_file(basedir: str, relname: str, data: bytes):
    path = os.path.join(basedir, "id1", relname)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        f.write(data)


def halton(i: int, base: int) -> float:
    f, r = 1.0, 0.0
    while i > 0:
        f /= base
        r += f * (i % base)
        i //= base
    return r


def calc_fov_y(fov_x: float, width: int, height: int) -> float:
    """client/cl_screen.c CalcFov: fov_y from fov_x and the window shape."""
    x = width / math.tan(fov_x / 360.0 * math.pi)
    return math.atan(height / x) * 360.0 / math.pi


def build_refdefs(specs: Sequence[ViewSpec], width: int, height: int,
                  model_resolver: Optional[Callable[[str], int]] = None, colormap_ptr: int = 0,
                  lightstyles=None):
    """refdef_t[n] for `specs`.  The returned ctypes array keeps every sub-array alive."""
    n = len(specs)
    ls = lightstyles if lightstyles is not None else default_lightstyles()
    rds = (refdef_t * n)()
    keep = [ls]
    cache = {}
    for i, sp in enumerate(specs):
        rd = rds[i]
        x, y, w, h = sp.rect if sp.rect else (0, 0, width, height)
        rd.x, rd.y, rd.width, rd.height = x, y, w, h
        rd.time = float(sp.time)
        rd.oldtime = float(sp.time)
        rd.flags = sp.flags
        rd.fov_x = sp.fov_x
        rd.fov_y = calc_fov_y(sp.fov_x, w, h) if sp.fov_y is None else sp.fov_y
        for k in range(3):
            rd.vieworg[k] = float(sp.origin[k])
            rd.viewangles[k] = float(sp.angles[k])
        rd.lightstyles = ls
        if sp.entities:
            ents = (entity_t * len(sp.entities))()
            for j, e in enumerate(sp.entities):
                if e.model not in cache:
                    cache[e.model] = model_resolver(e.model)
                ents[j].model = cache[e.model]
                ents[j].framelerp = e.framelerp
                ents[j].number = j + 1
                ents[j].flags = e.flags
                ents[j].skinnum = e.skinnum
                ents[j].colormap = colormap_ptr
                for k in range(3):
                    ents[j].origin[k] = float(e.origin[k])
                    ents[j].angles[k] = float(e.angles[k])
                ents[j].frame = e.frame
                ents[j].oldframe = e.oldframe
                ents[j].syncbase = e.syncbase
            rd.num_entities = len(sp.entities)
            rd.entities = ents
            keep.append(ents)
        if sp.dlights:
            dl = (dlight_t * len(sp.dlights))()
            for j, d in enumerate(sp.dlights):
                dl[j].origin[0], dl[j].origin[1], dl[j].origin[2], dl[j].radius = d
            rd.num_dlights = len(sp.dlights)
            rd.dlights = dl
            keep.append(dl)
        if sp.particles:
            pt = (particle_t * len(sp.particles))()
            for j, p in enumerate(sp.particles):
                pt[j].org[0], pt[j].org[1], pt[j].org[2] = p[0], p[1], p[2]
                pt[j].color = int(p[3])
            rd.num_particles = len(sp.particles)
            rd.particles = pt
            keep.append(pt)
        if sp.cshifts:
            cs = (cshift_t * len(sp.cshifts))()
            for j, c in enumerate(sp.cshifts):
                cs[j].destcolor[0], cs[j].destcolor[1], cs[j].destcolor[2], cs[j].percent = c
            rd.num_cshifts = len(sp.cshifts)
            rd.cshifts = cs
            keep.append(cs)
    rds._keep = keep
    return rds


def timerefresh_views(origin=(0.0, 0.0, 0.0), frames: int = 128, time: float = 1.0) -> List[ViewSpec]:
    """The reference's own benchmark path: yaw = i/128*360 (ref_soft/r_misc.c:42-58)."""
    return [ViewSpec(tuple(origin), (0.0, i / 128.0 * 360.0, 0.0), time) for i in range(frames)]


def halton_views(info: dict, n: int, start: int = 1, time: float = 1.0, time_step: float = 0.0) -> List[ViewSpec]:
    """Halton(2,3,5) positions inside rooms x Halton(7) yaw, pitch in [-20, 20];
    view i is rendered at time + i*time_step (C4: 0.01 s apart)."""
    spawn = info["spawn"]
    room = info.get("room", 400)
    out = []
    for i in range(n):
        k = start + i
        sx, sy, _ = spawn[k % len(spawn)]
        r = room * 0.5 - 24.0
        ox = sx + (halton(k, 2) * 2 - 1) * r
        oy = sy + (halton(k, 3) * 2 - 1) * r
        oz = 24.0 + halton(k, 5) * 80.0
        yaw = halton(k, 7) * 360.0
        pitch = (halton(k, 11) * 2 - 1) * 20.0
        out.append(ViewSpec((ox, oy, oz), (pitch, yaw, 0.0), time + i * time_step))
    return out
