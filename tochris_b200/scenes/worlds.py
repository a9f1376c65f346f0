"""The synthetic maps the BASELINE.json configs name.

box_room():    C1  one convex 512x512x256 room.
multi_room():  C2/C3/C4  n x n grid of rooms joined by door openings, optional water pool,
               sky-textured ceilings, animated '+0' textures and door brush models.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from . import assets
from .bspgen import BspBuilder, BrushModel, EMPTY, SOLID, WATER, Face


def box_room(seed: int = assets.SEED) -> Tuple[bytes, dict]:
    xs = [-272, -256, 256, 272]
    ys = [-272, -256, 256, 272]
    zs = [-144, -128, 128, 144]
    content = np.zeros((3, 3, 3), dtype=np.int8)
    content[1, 1, 1] = EMPTY
    b = BspBuilder(xs, ys, zs, content, seed=seed, multi_style_every=0)
    data = b.build()
    info = dict(stats=b.stats, spawn=[(0.0, 0.0, 0.0)], bounds=((-256, -256, -128), (256, 256, 128)))
    return data, info


def multi_room(n: int = 4, room: int = 320, wall: int = 32, height: int = 192, door_w: int = 96,
               door_h: int = 128, seed: int = assets.SEED, water: bool = False, sky: bool = False,
               anim: bool = False, doors: int = 0) -> Tuple[bytes, dict]:
    """n x n rooms of `room` units separated by `wall`-thick walls with centred door openings.
    Floors at z=0.  Rooms alternate ceiling heights to vary sightlines and mip levels."""
    rng = np.random.RandomState(seed)
    # x / y grid lines: per room [start, door_lo, door_hi, end] so door openings align with cells
    def axis_lines():
        lines = [-wall]
        pos = 0
        for i in range(n):
            d0 = pos + (room - door_w) // 2
            lines += [pos, d0, d0 + door_w, pos + room]
            pos += room + wall
        lines.append(pos - wall + wall)
        return sorted(set(lines))
    xs = axis_lines(); ys = axis_lines()
    zs = sorted(set([-32, -16, 0, 64, door_h, height - 48, height, height + 16]))
    content = np.zeros((len(xs) - 1, len(ys) - 1, len(zs) - 1), dtype=np.int8)
    xi = {v: i for i, v in enumerate(xs)}; yi = {v: i for i, v in enumerate(ys)}; zi = {v: i for i, v in enumerate(zs)}

    spawn: List[Tuple[float, float, float]] = []
    room_origin = lambda i: i * (room + wall)
    ceil_of = {}
    for rx in range(n):
        for ry in range(n):
            x0 = room_origin(rx); y0 = room_origin(ry)
            ceil = height if (rx + ry) % 2 == 0 else height - 48
            ceil_of[(rx, ry)] = ceil
            content[xi[x0]:xi[x0 + room], yi[y0]:yi[y0 + room], zi[0]:zi[ceil]] = EMPTY
            spawn.append((x0 + room / 2.0, y0 + room / 2.0, 56.0))
    # door openings through the walls between neighbouring rooms
    for rx in range(n):
        for ry in range(n):
            x0 = room_origin(rx); y0 = room_origin(ry)
            d0 = (room - door_w) // 2
            if rx + 1 < n:
                content[xi[x0 + room]:xi[x0 + room + wall], yi[y0 + d0]:yi[y0 + d0 + door_w], zi[0]:zi[door_h]] = EMPTY
            if ry + 1 < n:
                content[xi[x0 + d0]:xi[x0 + d0 + door_w], yi[y0 + room]:yi[y0 + room + wall], zi[0]:zi[door_h]] = EMPTY
    pool = None
    if water:
        # sunken pool in room (1,1) (or (0,0) if n==1): floor dug to z=-16.. and filled to z=0.. with water
        rx = ry = min(1, n - 1)
        x0 = room_origin(rx); y0 = room_origin(ry)
        d0 = (room - door_w) // 2
        content[xi[x0 + d0]:xi[x0 + d0 + door_w], yi[y0 + d0]:yi[y0 + d0 + door_w], zi[-16]:zi[0]] = WATER
        pool = (x0 + d0, y0 + d0, x0 + d0 + door_w, y0 + d0 + door_w)

    sky_rooms = set()
    if sky:
        sky_rooms = {(rx, ry) for rx in range(n) for ry in range(n) if (rx * 3 + ry) % 5 == 0}

    def tex_for(f: Face) -> str:
        if f.axis == 2 and f.side == 1 and sky_rooms:
            # ceiling: which room is it over?
            cx = (f.rect[0] + f.rect[1]) / 2; cy = (f.rect[2] + f.rect[3]) / 2
            rx = int(cx // (room + wall)); ry = int(cy // (room + wall))
            if (rx, ry) in sky_rooms and cx - rx * (room + wall) < room and cy - ry * (room + wall) < room:
                return "sky1"
        if f.axis == 2:
            return "floor" if f.side == 0 else "ceil"
        k = int(f.coord) // (room + wall)
        if anim and k % 3 == 1 and f.axis == 0:
            return "+0panel"
        return ["wall_a", "wall_b", "wall_c"][(k + f.axis) % 3]

    b = BspBuilder(xs, ys, zs, content, seed=seed, tex_for_face=tex_for)
    if anim:
        b.add_texture("+0panel", assets.make_texture_pixels(seed + 101, 64, 64, "stripes"))
        b.add_texture("+1panel", assets.make_texture_pixels(seed + 102, 64, 64, "tile"))
        b.add_texture("+2panel", assets.make_texture_pixels(seed + 103, 64, 64, "brick"))
    door_origins = []
    for d in range(doors):
        # a door slab standing in the doorway between room (d,0) and (d+1,0) (wraps over rows)
        rx = d % max(1, n - 1); ry = (d // max(1, n - 1)) % n
        x0 = room_origin(rx); y0 = room_origin(ry)
        d0 = (room - door_w) // 2
        b.add_brush_model(BrushModel((-8.0, -door_w / 2.0, 0.0), (8.0, door_w / 2.0, float(door_h)), "door"))
        door_origins.append((x0 + room + wall / 2.0, y0 + d0 + door_w / 2.0, 0.0))
    data = b.build()
    total = n * room + (n - 1) * wall
    info = dict(stats=b.stats, spawn=spawn, bounds=((0, 0, 0), (total, total, height)), room=room, wall=wall,
                n=n, pool=pool, door_origins=door_origins, ceil_of=ceil_of, door_w=door_w, door_h=door_h)
    return data, info
