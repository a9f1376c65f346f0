"""End to end through the step in front of the path: a scripted play session -- a viewer walking
through the `full` map with stairs-like height changes, rockets and grenades flying (trails, their lights), explosions,
lightning beams, damage and pick-up flashes, a gun in hand -- is turned into frames by the producers of rc_client.c
(V_CalcRefdef, CL_AddEntities, CL_AddParticles, CL_AddDlights, CL_AddCshifts, CL_AddTempEntities through the scene lists)
and rendered by RC_TimeDemo; every frame must equal the reference's render of the same refdef_t.  (That the producers
fill the lists as the reference's do is tests/test_client_fx.py; this is the chain.)"""
import ctypes as C
import math
import multiprocessing as mp
import random

import numpy as np
import pytest

import rcutil
from tochris_b200 import api, client
from tochris_b200.refdef import default_lightstyles, refdef_t, vec3
from tochris_b200.scenes import views

W, H, NFRAMES = 320, 240, 48


def V(*x):
    return vec3(*[float(v) for v in x])


def _produce(r, L, names):
    """the session; returns the scene (keeps the frames alive), the frame array and its length"""
    rnd = random.Random(4)
    libc = C.CDLL(None)
    libc.srand(11)
    models = {n: r.model(n) for n in names}
    flagsof = {"progs/ball.mdl": client.MF_ROCKET, "progs/blob.mdl": client.MF_GRENADE}
    cams = rcutil.scene_views("full", NFRAMES, time=0.0)
    fx, sc = L.RC_FxNew(0), L.RC_SceneNew()
    bolts = (C.c_void_p * 3)(models["progs/blob.mdl"], 0, 0)
    movers = [dict(num=10 + i, model=["progs/ball.mdl", "progs/blob.mdl"][i % 2], off=[rnd.uniform(-60, 60), rnd.uniform(-60, 60), rnd.uniform(-10, 30)],
                   vel=[rnd.uniform(-200, 200), rnd.uniform(-200, 200), rnd.uniform(-20, 20)], lerp=None) for i in range(6)]
    t = 1.0
    styles = default_lightstyles()                        # cl_lightstyle: what the oracle's frames get as well
    for k in range(NFRAMES):
        dt = rnd.choice([0.013, 0.02, 0.033, 0.05])
        old, t = t, t + dt
        cam = cams[k // 4]                                # the viewer stays around a camera spot for four frames
        vs = client.default_viewstate()
        vs.time, vs.oldtime, vs.host_frametime = t, old, dt
        vs.lerp_origin = V(cam.origin[0] + 3.0 * (k % 4), cam.origin[1], cam.origin[2] - 22 + (8 if k % 8 >= 4 else 0))
        vs.velocity = V(150.0, 40.0 * math.sin(k), 0.0)
        vs.viewangles = V(cam.angles[0], cam.angles[1] + 2.0 * (k % 4), 0.0)
        vs.onground, vs.health = 1, 100
        vs.vrect[:] = [0, 0, W, H]
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


def _to_specs(frames, n, nameof):
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


def _worker(q, libpath, basedir):
    try:
        r = api.RefCuda(basedir, W, H, libpath=libpath)
        r.load_world("maps/full.bsp")
        L = client.bind(r.lib)
        sc, frames, n, nameof = _produce(r, L, ["progs/ball.mdl", "progs/blob.mdl"])
        specs = _to_specs(frames, n, nameof)
        fb = np.zeros((n, H, W), np.uint8)
        td = r.timedemo(frames, n, 16, fb)
        counts = dict(entities=[frames[i].num_entities for i in range(n)], particles=[frames[i].num_particles for i in range(n)],
                      dlights=[frames[i].num_dlights for i in range(n)], cshifts=[frames[i].num_cshifts for i in range(n)],
                      flags=[frames[i].flags for i in range(n)])
        r.scene_free(sc)
        r.shutdown()
        q.put(("ok", dict(fb=fb, td=td, specs=specs, counts=counts)))
    except Exception:  # noqa: BLE001
        import traceback
        q.put(("err", traceback.format_exc()))


def _check(libpath, assets_dir):
    from oracle import run_oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_worker, args=(q, libpath, assets_dir))
    p.start()
    kind, r = q.get(timeout=900)
    p.join(timeout=30)
    assert kind == "ok", r
    c = r["counts"]
    assert max(c["particles"]) > 1000 and max(c["entities"]) > 10 and max(c["dlights"]) >= 3 and max(c["cshifts"]) == 4, c
    assert r["td"]["frames"] == NFRAMES - 1 and r["td"]["status"] & ~rcutil.RC_STAT_ALIAS_RANGE == 0, api.status_names(r["td"]["status"])
    o = run_oracle.render(assets_dir, "maps/full.bsp", W, H, r["specs"])
    diff = [int((o["fb"][i] != r["fb"][i]).sum()) for i in range(NFRAMES)]
    if r["td"]["status"] & rcutil.RC_STAT_ALIAS_RANGE:
        assert max(diff) <= rcutil.TOLERANCE * W * H, diff          # (the reference read outside a skin: rcutil.assert_parity's rule)
    else:
        assert sum(diff) == 0, f"frames that differ from the reference: {[(i, d) for i, d in enumerate(diff) if d]}"
    return r


def test_replay_of_a_play_session_sim(assets_dir, sim_lib, oracle_available):
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    _check(sim_lib, assets_dir)


@pytest.mark.gpu
def test_replay_of_a_play_session_gpu(assets_dir, oracle_available):
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    r = _check(api.LIBPATH, assets_dir)
    print("timedemo over the produced frames:", r["td"]["text"].strip())
