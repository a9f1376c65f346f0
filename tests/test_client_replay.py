"""End to end through the step in front of the path: a scripted play session -- a viewer walking
through the `full` map with stairs-like height changes, rockets and grenades flying (trails, their lights), explosions,
lightning beams, damage and pick-up flashes, a gun in hand -- is turned into frames by the producers of rc_client.c
(V_CalcRefdef, CL_AddEntities, CL_AddParticles, CL_AddDlights, CL_AddCshifts, CL_AddTempEntities through the scene lists)
and rendered by RC_TimeDemo; every frame must equal the reference's render of the same refdef_t.  (That the producers
fill the lists as the reference's do is tests/test_client_fx.py; this is the chain.)"""
import multiprocessing as mp

import numpy as np
import pytest

import rcutil
from tochris_b200 import api, client
from tochris_b200.scenes import session
from tochris_b200.scenes.session import H, NFRAMES, W


def _worker(q, libpath, basedir):
    try:
        r = api.RefCuda(basedir, W, H, libpath=libpath)
        r.load_world("maps/full.bsp")
        L = client.bind(r.lib)
        sc, frames, n, nameof = session.produce(r, L, rcutil.scene_views("full", NFRAMES, time=0.0))
        specs = session.to_specs(frames, n, nameof)
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
