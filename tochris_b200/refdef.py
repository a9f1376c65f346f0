"""ctypes mirrors of the renderer-boundary structs in /root/reference/client/ref.h:37-108
(LP64 layout; sizes asserted against the compiled reference in tests/test_oracle.py:
refdef_t 136 B, entity_t 80 B, particle_t 16 B, dlight_t 16 B, lightstyle_t 68 B)."""
from __future__ import annotations

import ctypes as C

MAX_LIGHTSTYLES = 64      # common/quakedef.h:105
MAX_STYLESTRING = 64      # common/quakedef.h:111
RF_VIEWERMODEL, RF_WEAPONMODEL, RF_DEPTHSCALE, RF_MINLIGHT = 1, 2, 4, 8   # quakedef.h:228-234
RF_FULLBRIGHT, RF_NOSHADOW, RF_CULLVIS = 16, 32, 64
RDF_NOWORLDMODEL, RDF_UNDERWATER = 1, 2                                   # quakedef.h:238-239

vec3 = C.c_float * 3


class particle_t(C.Structure):
    _fields_ = [("org", vec3), ("color", C.c_int)]


class dlight_t(C.Structure):
    _fields_ = [("origin", vec3), ("radius", C.c_float)]


class cshift_t(C.Structure):
    _fields_ = [("destcolor", C.c_int * 3), ("percent", C.c_int)]


class lightstyle_t(C.Structure):
    _fields_ = [("length", C.c_int), ("map", C.c_char * MAX_STYLESTRING)]


class entity_t(C.Structure):
    _fields_ = [
        ("framelerp", C.c_float),
        ("model", C.c_void_p),
        ("number", C.c_int),
        ("flags", C.c_int),
        ("skinnum", C.c_int),
        ("colormap", C.c_void_p),
        ("origin", vec3),
        ("angles", vec3),
        ("frame", C.c_int),
        ("oldframe", C.c_int),
        ("syncbase", C.c_float),
    ]


class refdef_t(C.Structure):
    _fields_ = [
        ("x", C.c_int), ("y", C.c_int), ("width", C.c_int), ("height", C.c_int),
        ("time", C.c_float), ("oldtime", C.c_float),
        ("flags", C.c_int),
        ("fov_x", C.c_float), ("fov_y", C.c_float),
        ("vieworg", vec3), ("viewangles", vec3),
        ("lightstyles", C.POINTER(lightstyle_t)),
        ("num_entities", C.c_int), ("entities", C.POINTER(entity_t)),
        ("num_dlights", C.c_int), ("dlights", C.POINTER(dlight_t)),
        ("num_particles", C.c_int), ("particles", C.POINTER(particle_t)),
        ("num_cshifts", C.c_int), ("cshifts", C.POINTER(cshift_t)),
    ]


def default_lightstyles():
    """Style 0 'm' (normal), 1 flicker, 2 slow pulse, 3 candle, 32 switchable 'm'
    -- the strings id's progs install (style 0-11), abbreviated."""
    arr = (lightstyle_t * MAX_LIGHTSTYLES)()
    table = {0: b"m", 1: b"mmnmmommommnonmmonqnmmo", 2: b"abcdefghijklmnopqrstuvwxyzyxwvutsrqponmlkjihgfedcba",
             3: b"mmmmmaaaaammmmmaaaaaabcdefgabcdefg", 32: b"m"}
    for k, v in table.items():
        arr[k].length = len(v)
        arr[k].map = v
    return arr
