"""ctypes mirrors of the client-side producer API (include/ref_cuda.h, tochris_b200/csrc/rc_client.c): the structs
CL_AddEntities reads (rc_cent_t / rc_clframe_t), the dynamic-light record and the effect numbers; `bind(lib)` declares
the RC_Fx* / RC_ClientAddEntities prototypes on a loaded library (product or host simulator)."""
from __future__ import annotations

import ctypes as C

from .refdef import cshift_t, vec3

(FX_ENTITY_PARTICLES, FX_POINT, FX_EXPLOSION, FX_EXPLOSION2, FX_BLOB_EXPLOSION, FX_RUN_EFFECT, FX_ROCKET_SPLASH,
 FX_LAVA_SPLASH, FX_TELEPORT_SPLASH, FX_RAIL_TRAIL, FX_ROCKET_TRAIL, FX_GRENADE_TRAIL, FX_BLOOD_TRAIL,
 FX_SLIGHT_BLOOD_TRAIL, FX_TRACER_TRAIL, FX_VOOR_TRAIL) = range(16)
EF_BRIGHTFIELD, EF_MUZZLEFLASH, EF_BRIGHTLIGHT, EF_DIMLIGHT = 1, 2, 4, 8                  # common/quakedef.h:210-213
MF_ROCKET, MF_GRENADE, MF_GIB, MF_ROTATE, MF_TRACER, MF_ZOMGIB, MF_TRACER2, MF_TRACER3 = 1, 2, 4, 8, 16, 32, 64, 128
MAX_DLIGHTS = 32
IT_INVISIBILITY, IT_INVULNERABILITY, IT_SUIT, IT_QUAD = 524288, 1048576, 2097152, 4194304     # common/quakedef.h:154-157


class rc_fxlight_t(C.Structure):      # cdlight_t, client/client.h:91-98
    _fields_ = [("origin", vec3), ("radius", C.c_float), ("die", C.c_float), ("decay", C.c_float), ("key", C.c_int)]


class rc_cent_t(C.Structure):
    _fields_ = [("number", C.c_int), ("model", C.c_void_p), ("modelflags", C.c_int), ("colormap", C.c_void_p),
                ("prev_origin", vec3), ("prev_angles", vec3), ("prev_frame", C.c_int),
                ("cur_origin", vec3), ("cur_angles", vec3), ("cur_frame", C.c_int), ("effects", C.c_int), ("skin", C.c_int),
                ("startlerp", C.c_float), ("deltalerp", C.c_float), ("frametime", C.c_float), ("syncbase", C.c_float),
                ("lerp_origin", vec3)]


class rc_clframe_t(C.Structure):
    _fields_ = [("time", C.c_double), ("maxclients", C.c_int), ("viewentity", C.c_int), ("chase_active", C.c_int),
                ("bobbingitems", C.c_int), ("viewangles", vec3)]


class rc_viewstate_t(C.Structure):
    _fields_ = [("time", C.c_double), ("oldtime", C.c_double), ("host_frametime", C.c_double),
                ("lerp_origin", vec3), ("current_origin", vec3), ("current_angles", vec3),
                ("velocity", vec3), ("viewangles", vec3), ("punchangle", vec3),
                ("viewheight", C.c_float),
                ("onground", C.c_int), ("health", C.c_int), ("items", C.c_int), ("intermission", C.c_int), ("chase_active", C.c_int),
                ("vrect", C.c_int * 4), ("fov_x", C.c_float),
                ("cl_rollangle", C.c_float), ("cl_rollspeed", C.c_float), ("cl_bob", C.c_float), ("cl_bobcycle", C.c_float),
                ("cl_bobup", C.c_float), ("v_kicktime", C.c_float),
                ("v_idlescale", C.c_float), ("v_iyaw_cycle", C.c_float), ("v_iroll_cycle", C.c_float), ("v_ipitch_cycle", C.c_float),
                ("v_iyaw_level", C.c_float), ("v_iroll_level", C.c_float), ("v_ipitch_level", C.c_float),
                ("gun_model", C.c_void_p), ("gun_frame", C.c_int), ("gun_oldframe", C.c_int), ("gun_frametime", C.c_float),
                ("colormap", C.c_void_p)]


def default_viewstate():
    """the cvar defaults of cl_view.c:33-57"""
    v = rc_viewstate_t()
    v.cl_rollangle, v.cl_rollspeed, v.cl_bob, v.cl_bobcycle, v.cl_bobup, v.v_kicktime = 2.0, 200.0, 0.02, 0.6, 0.5, 0.5
    v.v_idlescale, v.v_iyaw_cycle, v.v_iroll_cycle, v.v_ipitch_cycle = 0.0, 2.0, 0.5, 1.0
    v.v_iyaw_level, v.v_iroll_level, v.v_ipitch_level = 0.3, 0.1, 0.3
    v.fov_x, v.viewheight, v.health = 90.0, 22.0, 100
    return v


def bind(L):
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    L.RC_FxNew.argtypes = [C.c_int]
    L.RC_FxNew.restype = C.c_void_p
    L.RC_FxFree.argtypes = [C.c_void_p]
    L.RC_FxFree.restype = None
    L.RC_FxClear.argtypes = [C.c_void_p]
    L.RC_FxClear.restype = None
    L.RC_FxEffect.argtypes = [C.c_void_p, C.c_int, fp, fp, C.c_int, C.c_int, C.c_double]
    L.RC_FxAddParticles.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double]
    L.RC_FxParticleCount.argtypes = [C.c_void_p]
    L.RC_FxParticles.argtypes = [C.c_void_p, C.c_int, fp, fp, fp, fp, ip, ip]
    L.RC_FxAllocDlight.argtypes = [C.c_void_p, C.c_int, C.c_double]
    L.RC_FxAllocDlight.restype = C.POINTER(rc_fxlight_t)
    L.RC_FxLights.argtypes = [C.c_void_p]
    L.RC_FxLights.restype = C.POINTER(rc_fxlight_t)
    L.RC_FxAddDlights.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double]
    L.RC_ClientAddEntities.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(rc_cent_t), C.c_int, C.POINTER(rc_clframe_t)]
    L.RC_FxSetContentsColor.argtypes = [C.c_void_p, C.c_int]
    L.RC_FxSetContentsColor.restype = None
    L.RC_FxSetEmptyColor.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.RC_FxSetEmptyColor.restype = None
    L.RC_FxDamage.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.RC_FxDamage.restype = None
    L.RC_FxBonusFlash.argtypes = [C.c_void_p]
    L.RC_FxBonusFlash.restype = None
    L.RC_FxClearCshifts.argtypes = [C.c_void_p]
    L.RC_FxClearCshifts.restype = None
    L.RC_FxAddCshifts.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double]
    L.RC_FxCshifts.argtypes = [C.c_void_p]
    L.RC_FxCshifts.restype = C.POINTER(cshift_t)
    L.RC_FxDamageKick.argtypes = [C.c_void_p, C.c_int, C.c_int, fp, fp, fp, C.c_float, C.c_float, C.c_float]
    L.RC_FxDamageKick.restype = None
    L.RC_ClientCalcRefdef.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(rc_viewstate_t), C.c_void_p]
    L.RC_ClientSetViewContents.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.RC_ClientSetViewContents.restype = None
    L.RC_FxBeam.argtypes = [C.c_void_p, C.c_int, C.c_void_p, fp, fp, C.c_double]
    L.RC_ClientAddTempEntities.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int, fp, C.c_void_p, C.POINTER(C.c_void_p)]
    return L
