"""Procedural MDL6 alias models (format: ref_soft/modelgen.h:30-121, loader r_model.c:1349-1572).

icosphere(): a subdivided icosahedron, several keyframes (pulsating / squashed), optional
frame group and skin group, front/back skin halves with on-seam vertices."""
from __future__ import annotations

import struct

import numpy as np

from . import assets


def _icosphere(subdiv: int):
    t = (1.0 + 5 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache = {}
        nf = []

        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[k] = len(v) - 1
            return cache[k]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return np.array(v), f


def icosphere_mdl(seed: int = assets.SEED, subdiv: int = 2, radius: float = 24.0, nframes: int = 3,
                  skin_w: int = 128, skin_h: int = 128, frame_group: bool = True, skin_group: bool = False,
                  flip: bool = False) -> bytes:
    rng = np.random.RandomState(seed)
    verts, tris = _icosphere(subdiv)
    nv, nt = len(verts), len(tris)
    # keyframes: sphere, squashed, bulged
    shapes = []
    for k in range(nframes + (2 if frame_group else 0)):
        s = np.array([1.0 + 0.25 * np.sin(k * 1.3), 1.0 + 0.2 * np.cos(k * 0.7), 1.0 - 0.3 * np.sin(k * 0.9)])
        bump = 1.0 + 0.08 * np.sin(verts[:, 0] * 5 + k) * np.cos(verts[:, 2] * 4)
        shapes.append(verts * s * bump[:, None] * radius)
    allp = np.concatenate(shapes)
    mins, maxs = allp.min(axis=0), allp.max(axis=0)
    scale = (maxs - mins) / 255.0
    origin = mins
    nidx = rng.randint(0, 162, size=nv)

    def quant(p):
        q = np.clip(np.round((p - origin) / scale), 0, 255).astype(np.uint8)
        return q

    # skin: two halves (front: y < 0), noise + stripes, no fullbrights
    yy, xx = np.mgrid[0:skin_h, 0:skin_w]
    skins = []
    for k in range(2 if skin_group else 1):
        n = rng.randint(0, 6, size=(skin_h, skin_w))
        skin = (((xx // 8 + yy // 8 + k) % 12 + 1) * 16 + 6 + n).astype(np.uint8)
        skins.append(np.minimum(skin, 223).astype(np.uint8))
    half = skin_w // 2
    st = []
    for i in range(nv):
        x, y, z = verts[i]
        s = int((x * 0.5 + 0.5) * (half - 1))
        t = int((0.5 - z * 0.5) * (skin_h - 1))
        onseam = 0x20 if abs(y) < 0.12 else 0
        st.append((onseam, s, t))
    out = bytearray()
    out += b"IDPO" + struct.pack("<i", 6)
    out += struct.pack("<3f3ff3f", *scale, *origin, float(radius * 1.4), 0.0, 0.0, 0.0)
    numframes = nframes + (1 if frame_group else 0)
    out += struct.pack("<iiiiiiiif", len([0]) if not skin_group else 1, skin_w, skin_h, nv, nt, numframes, 0, 0, 10.0)
    if skin_group:
        out += struct.pack("<ii", 1, len(skins)) + struct.pack("<%df" % len(skins), *[0.3 * (i + 1) for i in range(len(skins))])
        for s_ in skins:
            out += s_.tobytes()
    else:
        out += struct.pack("<i", 0) + skins[0].tobytes()
    for o, s, t in st:
        out += struct.pack("<iii", o, s, t)
    for a, b, c in tris:
        cy = (verts[a][1] + verts[b][1] + verts[c][1]) / 3.0
        if flip:
            a, b, c = a, c, b
        out += struct.pack("<iiii", 1 if cy < 0 else 0, a, b, c)

    def frame(shape, name):
        q = quant(shape)
        d = bytes([int(q[:, 0].min()), int(q[:, 1].min()), int(q[:, 2].min()), 0,
                   int(q[:, 0].max()), int(q[:, 1].max()), int(q[:, 2].max()), 0])
        d += name.encode()[:15].ljust(16, b"\0")
        body = np.concatenate([q, nidx[:, None].astype(np.uint8)], axis=1)
        return d + body.tobytes(), q

    fi = 0
    if frame_group:
        # a frame GROUP must sit at model frame 0 (r_alias.c:110,678 index the group with the model frame)
        qs = [quant(shapes[nframes]), quant(shapes[nframes + 1])]
        allq = np.concatenate(qs)
        out += struct.pack("<i", 1)
        out += struct.pack("<i", 2) + bytes([int(allq[:, 0].min()), int(allq[:, 1].min()), int(allq[:, 2].min()), 0,
                                             int(allq[:, 0].max()), int(allq[:, 1].max()), int(allq[:, 2].max()), 0])
        out += struct.pack("<2f", 0.25, 0.5)
        for k in range(2):
            fr, _ = frame(shapes[nframes + k], f"grp{k}")
            out += fr
        fi += 1
    for k in range(nframes):
        fr, _ = frame(shapes[k], f"frame{k}")
        out += struct.pack("<i", 0) + fr
    return bytes(out)
