"""The producers in front of the scene lists (tochris_b200/csrc/rc_client.c): the client's particle
system, dynamic lights and CL_AddEntities against the UNMODIFIED client/cl_effects.c + client/cl_main.c of the reference
(oracle/_ref/libclfx.so, `make -C oracle clfx`).  Both sides get the same calls after the same srand (); compared bit for
bit: what reaches the scene every frame (V_AddParticle / V_AddDlight / V_AddEntity order included), the particle state
(position, velocity, ramp, die, colour, type in list order) and the 32 light records."""
import ctypes as C
import os
import random
import shutil

import numpy as np
import pytest

from tochris_b200 import api, client
from tochris_b200.refdef import cshift_t, entity_t, refdef_t, vec3

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLFX = os.path.join(ROOT, "oracle", "_ref", "libclfx.so")
libc = C.CDLL(None)
CAP = 4096
fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)


def V(*x):
    return vec3(*[float(v) for v in x])


@pytest.fixture
def ref(tmp_path):
    """a private copy of the oracle library: its function-static tracer counter starts at 0 for every test"""
    if not os.path.exists(CLFX):
        if not os.path.isdir("/root/reference"):
            pytest.skip("oracle/_ref/libclfx.so not built")
        import subprocess
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "clfx"])
    p = str(tmp_path / "libclfx_private.so")
    shutil.copy(CLFX, p)
    L = C.CDLL(p)
    L.clfx_init.argtypes = [C.c_uint, C.c_int]
    L.clfx_set_time.argtypes = [C.c_double, C.c_double]
    L.clfx_particles.argtypes = [fp, fp, fp, fp, ip, ip, C.c_int]
    L.clfx_rec_particles.argtypes = [fp, ip, C.c_int]
    L.clfx_rec_dlights.argtypes = [fp, fp, C.c_int]
    L.clfx_rec_entities.argtypes = [C.POINTER(entity_t), C.c_int]
    L.clfx_dlights.argtypes = [C.POINTER(client.rc_fxlight_t)]
    L.clfx_add_entities.argtypes = [C.POINTER(client.rc_cent_t), C.c_int, C.POINTER(client.rc_clframe_t)]
    L.clfx_vid_colormap.restype = C.c_void_p
    L.CL_AllocDlight.argtypes = [C.c_int]
    L.CL_AllocDlight.restype = C.POINTER(client.rc_fxlight_t)
    for name in ("CL_ParticleExplosion", "CL_BlobExplosion", "CL_LavaSplash", "CL_TeleportSplash"):
        getattr(L, name).argtypes = [fp]
    L.CL_ParticleExplosion2.argtypes = [fp, C.c_int, C.c_int]
    L.CL_RunParticleEffect.argtypes = [fp, fp, C.c_int, C.c_int]
    L.CL_RocketSplash.argtypes = [fp, fp, C.c_int]
    for name in ("CL_RailTrail", "CL_RocketTrail", "CL_GrenadeTrail", "CL_BloodTrail", "CL_SlightBloodTrail", "CL_VoorTrail"):
        getattr(L, name).argtypes = [fp, fp]
    L.CL_TracerTrail.argtypes = [fp, fp, C.c_int]
    L.CL_EntityParticles.argtypes = [C.POINTER(entity_t)]
    L.clfx_msg_set.argtypes = [fp, C.c_int]
    L.clfx_rec_cshifts.argtypes = [C.POINTER(cshift_t), C.c_int]
    L.clfx_cshifts.argtypes = [C.POINTER(cshift_t)]
    L.clfx_set_bolts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.clfx_set_frame.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, fp]
    L.V_SetContentsColor.argtypes = [C.c_int]
    L.clfx_set_contents.argtypes = [C.c_int]
    L.clfx_set_kick.argtypes = [C.c_float, C.c_float]
    L.clfx_set_viewer.argtypes = [C.c_int, fp, fp, C.c_float]
    L.clfx_calc_refdef.argtypes = [C.POINTER(client.rc_viewstate_t), C.POINTER(refdef_t), C.POINTER(entity_t)]
    L.CL_ParseBeam.argtypes = [C.c_void_p]
    return L


def _libs(sim_lib):
    out = [("sim", sim_lib)]
    if os.path.exists(api.LIBPATH):
        out.append(("product", api.LIBPATH))     # host code only: no GPU is touched
    return out


def _dump(n, a):
    return [np.ctypeslib.as_array(x)[: n * k].copy() for x, k in a]


class RefSide:
    """the reference: globals, one call per effect"""
    def __init__(self, L, seed, cap):
        self.L = L
        L.clfx_init(seed, cap)
        self.time = self.old = 0.0

    def set_time(self, t, old):
        self.time, self.old = t, old
        self.L.clfx_set_time(t, old)

    def effect(self, kind, a, b, i0, i1):
        L, F = self.L, client
        a, b = V(*a), V(*b)      # the trail functions walk `start` along: copies
        if kind == F.FX_ENTITY_PARTICLES:
            e = entity_t()
            e.origin = a
            L.CL_EntityParticles(C.byref(e))
        elif kind == F.FX_EXPLOSION: L.CL_ParticleExplosion(a)
        elif kind == F.FX_EXPLOSION2: L.CL_ParticleExplosion2(a, i0, i1)
        elif kind == F.FX_BLOB_EXPLOSION: L.CL_BlobExplosion(a)
        elif kind == F.FX_RUN_EFFECT: L.CL_RunParticleEffect(a, b, i0, i1)
        elif kind == F.FX_ROCKET_SPLASH: L.CL_RocketSplash(a, b, i0)
        elif kind == F.FX_LAVA_SPLASH: L.CL_LavaSplash(a)
        elif kind == F.FX_TELEPORT_SPLASH: L.CL_TeleportSplash(a)
        elif kind == F.FX_RAIL_TRAIL: L.CL_RailTrail(a, b)
        elif kind == F.FX_ROCKET_TRAIL: L.CL_RocketTrail(a, b)
        elif kind == F.FX_GRENADE_TRAIL: L.CL_GrenadeTrail(a, b)
        elif kind == F.FX_BLOOD_TRAIL: L.CL_BloodTrail(a, b)
        elif kind == F.FX_SLIGHT_BLOOD_TRAIL: L.CL_SlightBloodTrail(a, b)
        elif kind == F.FX_TRACER_TRAIL: L.CL_TracerTrail(a, b, i0)
        elif kind == F.FX_VOOR_TRAIL: L.CL_VoorTrail(a, b)
        else: raise ValueError(kind)

    def light(self, key, origin, radius, life, decay):
        dl = self.L.CL_AllocDlight(key)
        dl.contents.origin = V(*origin); dl.contents.radius = radius; dl.contents.die = self.time + life; dl.contents.decay = decay

    def add_entities(self, cents, frame):
        arr = (client.rc_cent_t * len(cents))(*cents)
        self.L.clfx_add_entities(arr, len(cents), C.byref(frame))
        return [V(*c.lerp_origin) for c in arr]

    def end_frame(self):
        """CL_AddParticles + CL_AddDlights; returns what reached the scene and the state afterwards"""
        L = self.L
        L.CL_AddParticles()
        L.CL_AddDlights()
        o, c = (C.c_float * (3 * CAP))(), (C.c_int * CAP)()
        n = L.clfx_rec_particles(o, c, CAP)
        lo, lr = (C.c_float * (3 * CAP))(), (C.c_float * CAP)()
        nl = L.clfx_rec_dlights(lo, lr, CAP)
        ents = (entity_t * 1024)()
        ne = L.clfx_rec_entities(ents, 1024)
        L.clfx_rec_clear()
        so, sv, sr, sd, sc, st = ((C.c_float * (3 * CAP))(), (C.c_float * (3 * CAP))(), (C.c_float * CAP)(), (C.c_float * CAP)(),
                                  (C.c_int * CAP)(), (C.c_int * CAP)())
        ns = L.clfx_particles(so, sv, sr, sd, sc, st, CAP)
        lights = (client.rc_fxlight_t * client.MAX_DLIGHTS)()
        L.clfx_dlights(lights)
        return dict(scene_p=_dump(n, [(o, 3), (c, 1)]), scene_l=_dump(nl, [(lo, 3), (lr, 1)]), ents=bytes(ents)[: ne * C.sizeof(entity_t)],
                    state=_dump(ns, [(so, 3), (sv, 3), (sr, 1), (sd, 1), (sc, 1), (st, 1)]), lights=bytes(lights), n=(n, nl, ne, ns))


class OurSide:
    """rc_client.c: one rc_fx_t, the scene recorder behind it"""
    def __init__(self, libpath, seed, cap):
        self.L = client.bind(C.CDLL(libpath))
        L = self.L
        L.RC_SceneNew.restype = C.c_void_p
        L.RC_SceneFree.argtypes = [C.c_void_p]
        L.RC_SceneClear.argtypes = [C.c_void_p]
        L.RC_SceneReset.argtypes = [C.c_void_p]
        L.RC_SceneEndFrame.argtypes = [C.c_void_p, C.POINTER(refdef_t), ip]
        L.RC_SceneFrames.argtypes = [C.c_void_p, ip]
        L.RC_SceneFrames.restype = C.POINTER(refdef_t)
        libc.srand(seed)
        self.fx = L.RC_FxNew(cap)
        self.scene = L.RC_SceneNew()
        L.RC_SceneClear(self.scene)
        self.time = self.old = 0.0

    def close(self):
        self.L.RC_FxFree(self.fx)
        self.L.RC_SceneFree(self.scene)

    def set_time(self, t, old):
        self.time, self.old = t, old

    def effect(self, kind, a, b, i0, i1):
        assert self.L.RC_FxEffect(self.fx, kind, V(*a), V(*b), i0, i1, self.time) >= 0

    def light(self, key, origin, radius, life, decay):
        dl = self.L.RC_FxAllocDlight(self.fx, key, self.time)
        dl.contents.origin = V(*origin); dl.contents.radius = radius; dl.contents.die = self.time + life; dl.contents.decay = decay

    def add_entities(self, cents, frame):
        arr = (client.rc_cent_t * len(cents))(*cents)
        assert self.L.RC_ClientAddEntities(self.fx, self.scene, arr, len(cents), C.byref(frame)) >= 0
        return [V(*c.lerp_origin) for c in arr]

    def end_frame(self):
        L = self.L
        L.RC_FxAddParticles(self.fx, self.scene, self.time, self.old)
        L.RC_FxAddDlights(self.fx, self.scene, self.time, self.old)
        rd = refdef_t()
        rd.width, rd.height, rd.fov_x, rd.fov_y = 320, 200, 90.0, 60.0
        k = L.RC_SceneEndFrame(self.scene, C.byref(rd), None)
        cnt = C.c_int()
        fr = L.RC_SceneFrames(self.scene, C.byref(cnt))[k]
        n, nl, ne = fr.num_particles, fr.num_dlights, fr.num_entities
        sp = [np.array([list(fr.particles[i].org) for i in range(n)], np.float32).reshape(-1), np.array([fr.particles[i].color for i in range(n)], np.int32)]
        sl = [np.array([list(fr.dlights[i].origin) for i in range(nl)], np.float32).reshape(-1), np.array([fr.dlights[i].radius for i in range(nl)], np.float32)]
        ents = b"".join(bytes(fr.entities[i]) for i in range(ne))
        L.RC_SceneReset(self.scene)
        L.RC_SceneClear(self.scene)
        so, sv, sr, sd, sc, st = ((C.c_float * (3 * CAP))(), (C.c_float * (3 * CAP))(), (C.c_float * CAP)(), (C.c_float * CAP)(),
                                  (C.c_int * CAP)(), (C.c_int * CAP)())
        ns = L.RC_FxParticles(self.fx, CAP, so, sv, sr, sd, sc, st)
        assert ns == L.RC_FxParticleCount(self.fx)
        lights = bytes((client.rc_fxlight_t * client.MAX_DLIGHTS).from_address(C.addressof(L.RC_FxLights(self.fx).contents)))
        return dict(scene_p=sp, scene_l=sl, ents=ents, state=_dump(ns, [(so, 3), (sv, 3), (sr, 1), (sd, 1), (sc, 1), (st, 1)]), lights=lights, n=(n, nl, ne, ns))


def _same(a, b, what):
    assert a["n"] == b["n"], (what, a["n"], b["n"])
    for key in ("scene_p", "scene_l", "state"):
        for i, (x, y) in enumerate(zip(a[key], b[key])):
            assert x.tobytes() == y.tobytes(), f"{what}: {key}[{i}] differs at {np.flatnonzero(x.view(np.uint32) != y.view(np.uint32))[:8]}"
    assert a["ents"] == b["ents"], f"{what}: entity lists differ"
    assert a["lights"] == b["lights"], f"{what}: light records differ"


def _particle_script(seed, frames):
    """per frame: (dt, [effects], [lights]); frame times from half a millisecond (particles of one frame's life survive
    into the next pass: the stale-ramp case) to a tenth of a second; bursts that overfill the 2048 slots"""
    rnd = random.Random(seed)
    F = client
    script = []
    for k in range(frames):
        dt = rnd.choice([0.0005, 0.004, 0.009, 0.013, 0.02, 0.05, 0.1])
        fx = []
        for _ in range(rnd.choice([0, 1, 1, 2, 3])):
            kind = rnd.randrange(2, 16) if rnd.random() > 0.15 else F.FX_ENTITY_PARTICLES
            a = [rnd.uniform(-500, 500) for _ in range(3)]
            b = [a[j] + rnd.uniform(-120, 120) for j in range(3)] if kind >= F.FX_RAIL_TRAIL else [rnd.uniform(-2, 2) for _ in range(3)]
            if kind == F.FX_RAIL_TRAIL and rnd.random() < 0.2:
                b = [a[0], a[1], a[2] + rnd.choice([-64.0, 64.0])]          # straight up / down: Vector2Angles' special case
            i0, i1 = 0, 0
            if kind == F.FX_EXPLOSION2: i0, i1 = rnd.randrange(0, 240), rnd.randrange(1, 16)
            if kind == F.FX_RUN_EFFECT: i0, i1 = rnd.randrange(0, 256), rnd.randrange(1, 60)
            if kind == F.FX_ROCKET_SPLASH: i0 = rnd.randrange(0, 256)
            if kind == F.FX_TRACER_TRAIL: i0 = rnd.choice([52, 230])
            fx.append((kind, a, b, i0, i1))
        lights = [(rnd.choice([0, 0, 3, 7, 12]), [rnd.uniform(-500, 500) for _ in range(3)], rnd.uniform(0, 400), rnd.choice([0.001, 0.05, 0.5, 3.0]), rnd.choice([0.0, 100.0, 300.0, 2000.0]))
                  for _ in range(rnd.choice([0, 0, 1, 2, 6]))]
        script.append((dt, fx, lights))
    return script


def _run(side, script, cents_per_frame=None):
    out, t = [], 1.0
    for k, (dt, fx, lights) in enumerate(script):
        old, t = t, t + dt
        side.set_time(t, old)
        for e in fx:
            side.effect(*e)
        for l in lights:
            side.light(*l)
        lerp = None
        if cents_per_frame:
            cents, frame = cents_per_frame(k, t)
            if k > 0:
                for c, lo in zip(cents, out[-1]["lerp"]):
                    c.lerp_origin = lo
            lerp = side.add_entities(cents, frame)
        r = side.end_frame()
        r["lerp"] = lerp
        out.append(r)
    return out


@pytest.mark.parametrize("seed,cap", [(1, 0), (2, 0), (3, 600), (4, 0)])
def test_particles_and_dlights_match_the_reference(ref, sim_lib, seed, cap):
    script = _particle_script(seed, 60)
    want = _run(RefSide(ref, 1000 + seed, cap), script)
    assert max(w["n"][3] for w in want) == (cap if cap else 2048), "the script never filled the particle slots"
    assert any(w["n"][1] > 0 for w in want) and sum(w["n"][0] for w in want) > 20000
    for name, lib in _libs(sim_lib):
        ours = OurSide(lib, 1000 + seed, cap)
        got = _run(ours, script)
        ours.close()
        for k, (a, b) in enumerate(zip(want, got)):
            _same(a, b, f"{name} frame {k}")


def _cent_maker(seed, colormap):
    rnd = random.Random(seed)
    F = client
    mflags = [0, 0, F.MF_ROTATE, F.MF_ROCKET, F.MF_GRENADE, F.MF_GIB, F.MF_ZOMGIB, F.MF_TRACER, F.MF_TRACER2, F.MF_TRACER3,
              F.MF_ROTATE | F.MF_ROCKET, F.MF_GIB | F.MF_TRACER]
    models = (C.c_int * len(mflags))(*mflags)          # the "models": Mod_Flags reads the int they point at
    N = 40
    base = []
    for i in range(N):
        base.append(dict(number=i + 1 if i < 36 else 100 + i, mi=rnd.randrange(len(mflags)) if i % 9 else None,
                         effects=rnd.choice([0, 0, 0, 1, 2, 4, 8, 2 | 8, 1 | 4]), skin=rnd.randrange(3),
                         p0=[rnd.uniform(-400, 400) for _ in range(3)], v=[rnd.uniform(-300, 300) for _ in range(3)],
                         a0=[rnd.uniform(-360, 360) for _ in range(3)], av=[rnd.uniform(-400, 400) for _ in range(3)],
                         deltalerp=rnd.choice([0.0, 0.1, 0.1, 0.05]), syncbase=rnd.random()))

    def make(k, t):
        cents = []
        for b in base:
            c = client.rc_cent_t()
            c.number = b["number"]
            if b["mi"] is not None:
                c.model = C.addressof(models) + 4 * b["mi"]
                c.modelflags = mflags[b["mi"]]
            c.colormap = colormap
            t0, t1 = 0.1 * int((t - 0.03) / 0.1), 0.1 * int((t - 0.03) / 0.1) + 0.1
            c.prev_origin = V(*[b["p0"][j] + b["v"][j] * t0 for j in range(3)]); c.cur_origin = V(*[b["p0"][j] + b["v"][j] * t1 for j in range(3)])
            c.prev_angles = V(*[(b["a0"][j] + b["av"][j] * t0) % 360 for j in range(3)]); c.cur_angles = V(*[(b["a0"][j] + b["av"][j] * t1) % 360 for j in range(3)])
            c.prev_frame, c.cur_frame = int(t0 * 10) % 7, int(t1 * 10) % 7
            c.effects, c.skin = b["effects"], b["skin"]
            c.startlerp, c.deltalerp, c.frametime, c.syncbase = t0 + 0.03, b["deltalerp"], t1 - 0.02 * (b["number"] % 3), b["syncbase"]
            cents.append(c)
        fr = client.rc_clframe_t()
        fr.time, fr.maxclients, fr.viewentity = t, 4, 2
        fr.chase_active, fr.bobbingitems = (k // 5) % 2, (k // 3) % 2
        fr.viewangles = V(10.0 + k, 30.0 * k, 0.0)
        return cents, fr
    make.keep = models
    return make


@pytest.mark.parametrize("seed", [11, 12])
def test_add_entities_matches_the_reference(ref, sim_lib, seed):
    """CL_AddEntities (cl_main.c:320-470): 40 entities per frame over 30 frames -- empty slots, every model-flag trail,
    every effect bit, players (RF_MINLIGHT), rotating / bobbing items, the viewer's own model with and without the
    chase camera -- with other effects and lights going on around them"""
    script = _particle_script(seed, 30)
    script = [(dt, fx[:1], lights[:1]) for dt, fx, lights in script]
    cmap = ref.clfx_vid_colormap()
    maker = _cent_maker(seed, cmap)            # one set of "models" for both sides: entity_t carries the pointer
    want = _run(RefSide(ref, 2000 + seed, 0), script, maker)
    assert sum(w["n"][2] for w in want) > 800 and any(w["n"][1] > 3 for w in want)
    flags = set()
    for w in want:
        ents = (entity_t * w["n"][2]).from_buffer_copy(w["ents"])
        flags |= {e.flags for e in ents}
    assert flags >= {0, 8, 9}, flags          # plain, RF_MINLIGHT (players), the viewer -- a player -- from the chase camera: + RF_VIEWERMODEL
    for name, lib in _libs(sim_lib):
        ours = OurSide(lib, 2000 + seed, 0)
        got = _run(ours, script, maker)
        ours.close()
        for k, (a, b) in enumerate(zip(want, got)):
            _same(a, b, f"{name} frame {k}")
            assert [bytes(x) for x in a["lerp"]] == [bytes(x) for x in b["lerp"]], f"{name} frame {k}: lerp_origin"


def test_fx_argument_errors(sim_lib):
    L = client.bind(C.CDLL(sim_lib))
    fx = L.RC_FxNew(100)                      # below the floor: 512 slots, as -particles 100 gives
    a = V(0, 0, 0)
    assert L.RC_FxEffect(fx, client.FX_EXPLOSION, a, None, 0, 0, 1.0) == 512
    assert L.RC_FxEffect(fx, client.FX_EXPLOSION, a, None, 0, 0, 1.0) == 0          # the free list is empty
    assert L.RC_FxEffect(fx, client.FX_EXPLOSION2, a, None, 0, 0, 1.0) == -1        # colorLength 0: a division by zero there
    assert L.RC_FxEffect(fx, client.FX_RAIL_TRAIL, a, None, 0, 0, 1.0) == -1
    assert L.RC_FxEffect(fx, 99, a, None, 0, 0, 1.0) == -1
    assert L.RC_FxAddParticles(fx, None, 7.0, 6.9) == 0 and L.RC_FxParticleCount(fx) == 0   # all dead at t = 7
    L.RC_FxFree(fx)


def test_colour_shifts_match_the_reference(ref, sim_lib):
    """V_SetContentsColor, CL_ParseDamage's colour half, V_BonusFlash_f, V_CalcPowerupCshift and CL_AddCshifts
    (cl_view.c:321-550): the shifts reach the scene only in frames in which one changed; damage and bonus fade by the
    frame time (an int minus a double, truncated)"""
    rnd = random.Random(5)
    F = client
    script = []
    for k in range(120):
        ev = []
        if rnd.random() < 0.25: ev.append(("contents", rnd.choice([-1, -2, -3, -4, -5, -6])))
        if rnd.random() < 0.2: ev.append(("damage", rnd.choice([0, 0, 5, 40, 100]), rnd.choice([0, 3, 20, 90])))
        if rnd.random() < 0.15: ev.append(("bonus",))
        if k == 70: ev.append(("clear",))
        items = rnd.choice([0, 0, 0, F.IT_QUAD, F.IT_SUIT, F.IT_INVISIBILITY, F.IT_INVULNERABILITY, F.IT_QUAD | F.IT_SUIT]) if k % 7 == 0 else None
        script.append((ev, items, rnd.choice([0.0, 0.001, 0.013, 0.02, 0.05, 0.1])))

    def run_ref():
        L = ref
        L.clfx_init(1, 0)
        out, items = [], 0
        for ev, it, dt in script:
            if it is not None: items = it
            L.clfx_set_frame(1.0, 1.0, dt, items, 1, None)
            for e in ev:
                if e[0] == "contents": L.V_SetContentsColor(e[1])
                elif e[0] == "damage":
                    L.clfx_msg_set((C.c_float * 5)(e[1], e[2], 10.0, 20.0, 30.0), 5)
                    L.CL_ParseDamage()
                elif e[0] == "bonus": L.V_BonusFlash_f()
                else: L.CL_ClearCshifts()
            L.clfx_rec_clear()
            L.CL_AddCshifts()
            sent = (cshift_t * 16)()
            n = L.clfx_rec_cshifts(sent, 16)
            state = (cshift_t * 4)()
            L.clfx_cshifts(state)
            out.append((n, bytes(sent)[: n * 16], bytes(state)))
        return out

    def run_ours(lib):
        L = client.bind(C.CDLL(lib))
        L.RC_SceneNew.restype = C.c_void_p
        for nm in ("RC_SceneFree", "RC_SceneClear", "RC_SceneReset"): getattr(L, nm).argtypes = [C.c_void_p]
        L.RC_SceneEndFrame.argtypes = [C.c_void_p, C.POINTER(refdef_t), ip]
        L.RC_SceneFrames.argtypes = [C.c_void_p, ip]
        L.RC_SceneFrames.restype = C.POINTER(refdef_t)
        fx, sc = L.RC_FxNew(0), L.RC_SceneNew()
        out, items = [], 0
        for ev, it, dt in script:
            if it is not None: items = it
            for e in ev:
                if e[0] == "contents": L.RC_FxSetContentsColor(fx, e[1])
                elif e[0] == "damage": L.RC_FxDamage(fx, e[1], e[2])
                elif e[0] == "bonus": L.RC_FxBonusFlash(fx)
                else: L.RC_FxClearCshifts(fx)
            L.RC_SceneReset(sc)
            n = L.RC_FxAddCshifts(fx, sc, items, dt)
            rd = refdef_t()
            k = L.RC_SceneEndFrame(sc, C.byref(rd), None)
            cnt = C.c_int()
            fr = L.RC_SceneFrames(sc, C.byref(cnt))[k]
            assert fr.num_cshifts == n
            sent = b"".join(bytes(fr.cshifts[i]) for i in range(n))
            state = bytes((cshift_t * 4).from_address(C.addressof(L.RC_FxCshifts(fx).contents)))
            out.append((n, sent, state))
        L.RC_FxFree(fx); L.RC_SceneFree(sc)
        return out

    want = run_ref()
    assert 10 < sum(1 for w in want if w[0] == 0) < 110 and any(w[0] == 4 for w in want)      # frames with and without a change
    for name, lib in _libs(sim_lib):
        got = run_ours(lib)
        for k, (a, b) in enumerate(zip(want, got)):
            assert a == b, f"{name} frame {k}: {a[0]} / {b[0]} shifts sent"


def test_beams_match_the_reference(ref, sim_lib):
    """CL_ParseBeam + CL_AddTempEntities (cl_tent.c:60-105, 303-356): replaced / new / expired beams, the viewer's own beam
    starting at the viewer's drawn position, full-bright bolt models, the 128-entity limit, straight-up beams
    (Vector2Angles' special case) and the reference's shared loop counter"""
    rnd = random.Random(9)
    models = (C.c_int * 4)(0, 0, 0, 0)
    mp = [C.addressof(models) + 4 * i for i in range(4)]
    frames = []
    t = 1.0
    for k in range(40):
        # (at most four beams alive, in slots 0-3: with a drawn beam in a higher slot the reference's shared loop counter
        # walks it past the end of cl_beams[] -- undefined there, not reproduced)
        t += 1.0 if k in (25, 26) else rnd.choice([0.02, 0.05, 0.15, 0.3])
        beams = []
        for _ in range(rnd.choice([0, 1, 1, 2, 3])):
            a = [rnd.uniform(-300, 300) for _ in range(3)]
            b = [a[0], a[1], a[2] + rnd.choice([-200.0, 150.0])] if rnd.random() < 0.15 else [a[j] + rnd.uniform(-400, 400) for j in range(3)]
            beams.append((rnd.choice([1, 1, 2, 5, 9]), rnd.randrange(4), a, b))
        if k == 25:
            beams = [(100 + j, 0, [0.0, 0.0, 0.0], [3000.0, 100.0 * j, 50.0]) for j in range(3)]     # 100 segments each: the limit
        frames.append((t, beams, [rnd.uniform(-50, 50) for _ in range(3)]))

    def run_ref():
        L = ref
        L.clfx_init(77, 0)
        L.clfx_set_bolts(mp[0], mp[1], mp[2])
        out = []
        for t, beams, vorg in frames:
            L.clfx_set_frame(t, t, 0.0, 0, 1, V(*vorg))
            for ent, mi, a, b in beams:
                L.clfx_msg_set((C.c_float * 7)(ent, *a, *b), 7)
                L.CL_ParseBeam(mp[mi])
            L.clfx_rec_clear()
            L.CL_AddTempEntities()
            ents = (entity_t * 256)()
            n = L.clfx_rec_entities(ents, 256)
            out.append((n, bytes(ents)[: n * C.sizeof(entity_t)]))
        return out, L.clfx_vid_colormap()

    def run_ours(lib, cmap):
        L = client.bind(C.CDLL(lib))
        L.RC_SceneNew.restype = C.c_void_p
        for nm in ("RC_SceneFree", "RC_SceneClear", "RC_SceneReset"): getattr(L, nm).argtypes = [C.c_void_p]
        L.RC_SceneEndFrame.argtypes = [C.c_void_p, C.POINTER(refdef_t), ip]
        L.RC_SceneFrames.argtypes = [C.c_void_p, ip]
        L.RC_SceneFrames.restype = C.POINTER(refdef_t)
        libc.srand(77)
        fx, sc = L.RC_FxNew(0), L.RC_SceneNew()
        bolts = (C.c_void_p * 3)(mp[0], mp[1], mp[2])
        out = []
        for t, beams, vorg in frames:
            for ent, mi, a, b in beams:
                assert L.RC_FxBeam(fx, ent, mp[mi], V(*a), V(*b), t) in (0, 1)
            L.RC_SceneReset(sc)
            n = L.RC_ClientAddTempEntities(fx, sc, t, 1, V(*vorg), cmap, bolts)
            rd = refdef_t()
            k = L.RC_SceneEndFrame(sc, C.byref(rd), None)
            cnt = C.c_int()
            fr = L.RC_SceneFrames(sc, C.byref(cnt))[k]
            assert fr.num_entities == n
            out.append((n, b"".join(bytes(fr.entities[i]) for i in range(n))))
        L.RC_FxFree(fx); L.RC_SceneFree(sc)
        return out

    want, cmap = run_ref()
    assert max(w[0] for w in want) == 128 and sum(w[0] for w in want) > 600
    for name, lib in _libs(sim_lib):
        got = run_ours(lib, cmap)
        for k, (a, b) in enumerate(zip(want, got)):
            assert a[0] == b[0], f"{name} frame {k}: {a[0]} / {b[0]} beam segments"
            assert a[1] == b[1], f"{name} frame {k}: segment entities differ"


def test_calc_refdef_matches_the_reference(ref, sim_lib):
    """V_CalcRefdef / V_CalcIntermissionRefdef with V_CalcBob, V_CalcRoll, V_CalcViewRoll (damage kicks from
    CL_ParseDamage), V_AddIdle, V_BoundOffsets, the stair smoothing, V_AddGun and the contents at the eye
    (cl_view.c:159-210, 331-391, 646-886): a 200-frame walk with steps up and down, strafing, damage from all sides, idle
    sway on and off, death, invisibility, an intermission -- refdef_t, the gun entity and the medium's colour shift bit for bit"""
    rnd = random.Random(21)
    gun = (C.c_int * 1)(0)
    cmap = ref.clfx_vid_colormap()
    frames = []
    t, z = 5.0, 24.0
    for k in range(200):
        dt = rnd.choice([0.0, 0.008, 0.013, 0.02, 0.05, 0.1])
        old, t = t, t + dt
        if rnd.random() < 0.2: z += rnd.choice([16.0, 18.0, 8.0, -16.0, 30.0])        # stairs, drops
        v = client.default_viewstate()
        v.time, v.oldtime, v.host_frametime = t, old, dt
        v.lerp_origin = V(10.0 * k, -3.0 * k, z)
        v.current_origin = V(10.0 * k + 1, -3.0 * k, z); v.current_angles = V(5.0, 90.0 + k, 0.0)
        v.velocity = V(rnd.uniform(-400, 400), rnd.uniform(-400, 400), rnd.uniform(-100, 100)) if k % 9 else V(0, 0, 0)
        v.viewangles = V(rnd.uniform(-60, 60), rnd.uniform(0, 360), rnd.uniform(-5, 5))
        v.punchangle = V(rnd.uniform(-4, 0), 0.0, 0.0) if k % 5 == 0 else V(0, 0, 0)
        v.onground = int(k % 4 != 3)
        v.health = -5 if 120 <= k < 130 else 100
        v.items = client.IT_INVISIBILITY if 60 <= k < 70 else 0
        v.intermission = int(150 <= k < 160)
        v.vrect[:] = [0, 0, 320, 200] if k % 2 else [16, 8, 288, 160]
        v.fov_x = rnd.choice([90.0, 110.0, 45.0])
        v.v_idlescale = 1.5 if 80 <= k < 110 else 0.0
        v.viewheight = rnd.choice([22.0, 22.0, 12.0])
        v.gun_model = C.addressof(gun) if k % 11 else None
        v.gun_frame, v.gun_oldframe, v.gun_frametime = k % 6, (k + 5) % 6, t - rnd.choice([0.0, 0.03, 0.07, 0.2])
        v.colormap = cmap
        dmg = (rnd.choice([0, 30, 80]), rnd.choice([5, 25, 90]), [10.0 * k + rnd.uniform(-100, 100), -3.0 * k + rnd.uniform(-100, 100), z + 20]) if rnd.random() < 0.15 else None
        frames.append((v, dmg, rnd.choice([-1, -1, -2, -3, -4, -5])))

    def run_ref():
        L = ref
        L.clfx_init(3, 0)
        L.clfx_set_kick(0.6, 0.6)
        out = []
        for v, dmg, contents in frames:
            if dmg:
                L.clfx_set_viewer(1, v.lerp_origin, v.viewangles, v.v_kicktime)
                L.clfx_msg_set((C.c_float * 5)(dmg[0], dmg[1], *dmg[2]), 5)
                L.CL_ParseDamage()
            L.clfx_set_contents(contents)
            rd, g = refdef_t(), entity_t()
            n = L.clfx_calc_refdef(C.byref(v), C.byref(rd), C.byref(g))
            cs = (cshift_t * 4)()
            L.clfx_cshifts(cs)
            out.append((n, [rd.x, rd.y, rd.width, rd.height, rd.flags], bytes(rd)[28:60], bytes(g) if n else b"", bytes(cs)[:16]))
        return out

    def run_ours(lib):
        L = client.bind(C.CDLL(lib))
        L.RC_SceneNew.restype = C.c_void_p
        for nm in ("RC_SceneFree", "RC_SceneClear", "RC_SceneReset"): getattr(L, nm).argtypes = [C.c_void_p]
        L.RC_SceneEndFrame.argtypes = [C.c_void_p, C.POINTER(refdef_t), ip]
        L.RC_SceneFrames.argtypes = [C.c_void_p, ip]
        L.RC_SceneFrames.restype = C.POINTER(refdef_t)
        fx, sc = L.RC_FxNew(0), L.RC_SceneNew()
        out = []
        for v, dmg, contents in frames:
            if dmg:
                L.RC_FxDamage(fx, dmg[0], dmg[1])
                L.RC_FxDamageKick(fx, dmg[0], dmg[1], V(*dmg[2]), v.lerp_origin, v.viewangles, 0.6, 0.6, v.v_kicktime)
            L.RC_SceneReset(sc)
            rd = refdef_t()
            n = L.RC_ClientCalcRefdef(fx, sc, C.byref(v), C.byref(rd))
            assert n in (0, 1)
            L.RC_ClientSetViewContents(fx, C.byref(rd), contents)
            k = L.RC_SceneEndFrame(sc, C.byref(refdef_t()), None)
            cnt = C.c_int()
            fr = L.RC_SceneFrames(sc, C.byref(cnt))[k]
            assert fr.num_entities == n
            g = bytes(fr.entities[0]) if n else b""
            cs = bytes((cshift_t * 4).from_address(C.addressof(L.RC_FxCshifts(fx).contents)))
            out.append((n, [rd.x, rd.y, rd.width, rd.height, rd.flags], bytes(rd)[28:60], g, cs[:16]))
        L.RC_FxFree(fx); L.RC_SceneFree(sc)
        return out

    want = run_ref()
    assert {w[0] for w in want} == {0, 1} and {w[1][4] for w in want} == {0, 2}
    for name, lib in _libs(sim_lib):
        got = run_ours(lib)
        for k, (a, b) in enumerate(zip(want, got)):
            assert a[1] == b[1], f"{name} frame {k}: rectangle / flags {a[1]} {b[1]}"
            assert a[2] == b[2], (f"{name} frame {k}: fov / vieworg / viewangles " +
                                  f"{np.frombuffer(a[2], np.float32)} {np.frombuffer(b[2], np.float32)}")
            assert a[0] == b[0] and a[3] == b[3], f"{name} frame {k}: the gun entity differs"
            assert a[4] == b[4], f"{name} frame {k}: contents colour"
