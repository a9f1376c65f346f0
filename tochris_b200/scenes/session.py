"""A scripted play session for the producers of rc_client.c (the step in front of the rendering path): a viewer walking
through the `full` map with a gun in hand while rockets and grenades fly (trails, their lights), explosions go off,
lightning beams reach out, damage and pick-up flashes tint the view.  `produce` turns it into frames through
RC_ClientCalcRefdef / RC_ClientAddEntities / RC_FxAddParticles / RC_FxAddDlights / RC_FxAddCshifts /
RC_ClientAddTempEntities and the scene lists; `to_specs` turns the frames back into ViewSpecs (what the oracle renders).
Used by tests/test_client_replay.py and tools/replay_session.py."""
import ctypes as C
import math
import random

from .. import client
from ..refdef import default_lightstyles, refdef_t, vec3
from . import views

W, H, NFRAMES = 320, 240, 48


def V(*x):
    return vec3(*[float(v) for v in x])


def produce(r, L, cams, names=("progs/ball.mdl", "progs/blob.mdl"), nframes=NFRAMES, width=W, height=H):
    """the session; returns the scene (keeps the frames alive), the frame array and its length"""
    rnd = random.Random(4)
    libc = C.CDLL(None)
    libc.srand(11)
    models = {n: r.model(n) for n in names}
    flagsof = {"progs/ball.mdl": client.MF_ROCKET, "progs/blob.mdl": client.MF_GRENADE}
    fx, sc = L.RC_FxNew(0), L.RC_SceneNew()
    bolts = (C.c_void_p * 3)(models["progs/blob.mdl"], 0, 0)
    movers = [dict(num=10 + i, model=["progs/ball.mdl", "progs/blob.mdl"][i % 2], off=[rnd.uniform(-60, 60), rnd.uniform(-60, 60), rnd.uniform(-10, 30)],
                   vel=[rnd.uniform(-200, 200), rnd.uniform(-200, 200), rnd.uniform(-20, 20)], lerp=None) for i in range(6)]
    t = 1.0
    styles = default_lightstyles()                        # cl_lightstyle: what the oracle's frames get as well
    for k in range(nframes):
        dt = rnd.choice([0.013, 0.02, 0.033, 0.05])
        old, t = t, t + dt
        cam = cams[(k // 4) % len(cams)]                                # the viewer stays around a camera spot for four frames
        vs = client.default_viewstate()
        vs.time, vs.oldtime, vs.host_frametime = t, old, dt
        vs.lerp_origin = V(cam.origin[0] + 3.0 * (k % 4), cam.origin[1], cam.origin[2] - 22 + (8 if k % 8 >= 4 else 0))
        vs.velocity = V(150.0, 40.0 * math.sin(k), 0.0)
        vs.viewangles = V(cam.angles[0], cam.angles[1] + 2.0 * (k % 4), 0.0)
        vs.onground, vs.health = 1, 100
        vs.vrect[:] = [0, 0, width, height]
        vs.fov_x = 90.0
        vs.v_idlescale = 1.0 if k >= 30 else 0.0
        vs.gun_model, vs.gun_frame, vs.gun_oldframe, vs.gun_frametime = models["progs/blob.mdl"], k % 2, (k + 1) % 2, t - 0.04
        L.RC_SceneClear(sc)
        if k % 9 == 3:
            L.RC_FxDamage(fx, 20, 30)
            L.RC_FxDamageKick(fx, 20, 30, V(cam.origin[0] + 50, cam.origin[1] - 80, cam.origin[2]), vs.lerp_origin, vs.viewangles, 0.6, 0.6, 0.5)
        if k % 11 == 5:
            L.RC_FxBonusFlash(fx)
        rd = refdef_t()
        assert L.RC_ClientCalcRefdef(fx, sc, C.byref(vs), C.byref(rd)) == 1
        L.RC_ClientSetViewContents(fx, C.byref(rd), r.point_contents(list(rd.vieworg)))
        rd.time, rd.oldtime = t, old
        rd.lightstyles = styles
        # what flies around the viewer
        fwd = (math.cos(math.radians(cam.angles[1])), math.sin(math.radians(cam.angles[1])))
        cents = []
        for m in movers:
            c = client.rc_cent_t()
            c.number, c.model, c.modelflags = m["num"], models[m["model"]], flagsof[m["model"]]
            base = [cam.origin[0] + fwd[0] * 120 + m["off"][0], cam.origin[1] + fwd[1] * 120 + m["off"][1], cam.origin[2] + m["off"][2]]
            c.prev_origin = V(*[base[j] + m["vel"][j] * (old - 1.0) for j in range(3)])
            c.cur_origin = V(*[base[j] + m["vel"][j] * (t - 1.0) for j in range(3)])
            c.prev_angles = V(0, 10.0 * k, 0); c.cur_angles = V(0, 10.0 * k + 10, 0)
            c.prev_frame, c.cur_frame = k % 2, (k + 1) % 2
            c.startlerp, c.deltalerp, c.frametime = old, dt, t - 0.05
            c.effects = client.EF_DIMLIGHT if m["num"] == 12 else (client.EF_BRIGHTFIELD if m["num"] == 13 and k % 6 == 0 else 0)
            c.lerp_origin = m["lerp"] if m["lerp"] is not None and k % 4 else c.prev_origin     # (a new spot: no trail across the map)
            cents.append(c)
        arr = (client.rc_cent_t * len(cents))(*cents)
        fr = client.rc_clframe_t()
        fr.time, fr.maxclients, fr.viewentity = t, 4, 1
        assert L.RC_ClientAddEntities(fx, sc, arr, len(cents), C.byref(fr)) == len(cents)
        for m, c in zip(movers, arr):
            m["lerp"] = V(*c.lerp_origin)
        spot = [cam.origin[0] + fwd[0] * 150, cam.origin[1] + fwd[1] * 150, cam.origin[2]]
        if k % 7 == 2:
            L.RC_FxEffect(fx, client.FX_EXPLOSION, V(*spot), None, 0, 0, t)
            dl = L.RC_FxAllocDlight(fx, 0, t)
            dl.contents.origin = V(*spot); dl.contents.radius = 350.0; dl.contents.die = t + 0.5; dl.contents.decay = 300.0
        if k % 13 == 6:
            L.RC_FxEffect(fx, client.FX_TELEPORT_SPLASH, V(spot[0], spot[1] + 40, spot[2]), None, 0, 0, t)
        if k % 5 == 1:
            L.RC_FxBeam(fx, 1, models["progs/blob.mdl"], V(*rd.vieworg), V(spot[0], spot[1], spot[2] + 20), t)
        L.RC_FxAddParticles(fx, sc, t, old)
        L.RC_FxAddDlights(fx, sc, t, old)
        L.RC_FxAddCshifts(fx, sc, 0, dt)
        L.RC_ClientAddTempEntities(fx, sc, t, 1, vs.lerp_origin, None, bolts)
        assert L.RC_SceneEndFrame(sc, C.byref(rd), None) == k
    cnt = C.c_int()
    frames = L.RC_SceneFrames(sc, C.byref(cnt))
    L.RC_FxFree(fx)
    return sc, frames, cnt.value, {v: k for k, v in models.items()}


def to_specs(frames, n, nameof):
    out = []
    for i in range(n):
        rd = frames[i]
        ents = [views.EntitySpec(nameof[e.model], tuple(e.origin), tuple(e.angles), frame=e.frame, oldframe=e.oldframe, framelerp=e.framelerp,
                                 skinnum=e.skinnum, flags=e.flags, syncbase=e.syncbase) for e in (rd.entities[j] for j in range(rd.num_entities))]
        out.append(views.ViewSpec(tuple(rd.vieworg), tuple(rd.viewangles), time=rd.time, fov_x=rd.fov_x, fov_y=rd.fov_y, flags=rd.flags,
                                  rect=(rd.x, rd.y, rd.width, rd.height), entities=ents,
                                  dlights=[(*rd.dlights[j].origin, rd.dlights[j].radius) for j in range(rd.num_dlights)],
                                  particles=[(*rd.particles[j].org, rd.particles[j].color) for j in range(rd.num_particles)],
                                  cshifts=[(*rd.cshifts[j].destcolor, rd.cshifts[j].percent) for j in range(rd.num_cshifts)]))
    return out


