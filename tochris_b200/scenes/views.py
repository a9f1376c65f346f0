"""Camera sets for the BASELINE.json configs and asset-directory writer."""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import numpy as np

from . import assets
from ..refdef import refdef_t, default_lightstyles


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


def make_refdefs(n: int, width: int, height: int, cams: Sequence[Tuple[Tuple[float, float, float], Tuple[float, float, float]]],
                 times: Sequence[float], fov_x: float = 90.0, lightstyles=None):
    """n refdef_t structs sharing one lightstyle table (the table must outlive the array)."""
    ls = lightstyles if lightstyles is not None else default_lightstyles()
    rds = (refdef_t * n)()
    import math
    for i in range(n):
        rd = rds[i]
        rd.x, rd.y, rd.width, rd.height = 0, 0, width, height
        rd.time = float(times[i]); rd.oldtime = float(times[i])
        rd.flags = 0
        rd.fov_x = fov_x
        # client/cl_screen.c CalcFov: fov_y from fov_x and the window shape
        x = width / math.tan(fov_x / 360.0 * math.pi)
        rd.fov_y = math.atan(height / x) * 360.0 / math.pi
        org, ang = cams[i]
        for k in range(3):
            rd.vieworg[k] = float(org[k]); rd.viewangles[k] = float(ang[k])
        rd.lightstyles = ls
    rds._keep = ls
    return rds


def timerefresh_cams(origin=(0.0, 0.0, 0.0), frames: int = 128):
    """The reference's own benchmark path: yaw = i/128*360 (ref_soft/r_misc.c:42-58)."""
    return [(tuple(origin), (0.0, i / 128.0 * 360.0, 0.0)) for i in range(frames)]


def halton_cams(info: dict, n: int, start: int = 1):
    """Halton(2,3,5) positions inside rooms x Halton(7) yaw, pitch in [-20, 20]."""
    spawn = info["spawn"]; room = info.get("room", 400)
    cams = []
    for i in range(n):
        k = start + i
        sx, sy, sz = spawn[k % len(spawn)]
        r = room * 0.5 - 24.0
        ox = sx + (halton(k, 2) * 2 - 1) * r
        oy = sy + (halton(k, 3) * 2 - 1) * r
        oz = 24.0 + halton(k, 5) * 80.0
        yaw = halton(k, 7) * 360.0
        pitch = (halton(k, 11) * 2 - 1) * 20.0
        cams.append(((ox, oy, oz), (pitch, yaw, 0.0)))
    return cams
