"""Device lifetimes of the alias models ("model cache eviction -> device upload lifetimes").  The
reference keeps alias models in the Cache_Alloc heap (ref_soft/r_model.c:95-130, common/zone.c): least recently used out
first, reloaded on demand, Mod_TouchModel (r_model.c:247) refreshes a model's place.  RC_SetModelCacheBytes gives the
device copy the same rules; pixels must never depend on it: every batch is compared with the same batch rendered with
everything resident, and those with the reference."""
import multiprocessing as mp

import numpy as np
import pytest

import rcutil
from tochris_b200 import api
from tochris_b200.scenes import mdlgen, views

W, H = 320, 200


def _batches():
    """three batches over the `full` map: A draws ball + blob, B draws m2 only, C draws ball only"""
    out = []
    for k, models in enumerate([("progs/ball.mdl", "progs/blob.mdl"), ("progs/m2.mdl",), ("progs/ball.mdl",)]):
        specs = rcutil.scene_views("full", 2, time=0.5 + k, entities=6)
        for sp in specs:
            ents = []
            for j, e in enumerate(sp.entities):
                m = models[j % len(models)]
                ents.append(views.EntitySpec(m, e.origin, e.angles, frame=e.frame % 2, oldframe=e.oldframe % 2, framelerp=e.framelerp, skinnum=0))
            sp.entities = ents
        out.append(specs)
    return out


def _worker(q, libpath, basedir):
    try:
        r = api.RefCuda(basedir, W, H, libpath=libpath)
        r.load_world("maps/full.bsp")
        batches = _batches()
        rds = [views.build_refdefs(sp, W, H, model_resolver=r.model) for sp in batches]   # loads ball, blob, m2
        res = {}
        # everything resident
        res["plain"] = [r.render(rd)[0] for rd in rds]
        res["info0"] = r.model_cache_info()
        per_model = res["info0"]["resident_bytes"]
        # room for ball + blob, not for a third model
        r.render(rds[0])
        r.set_model_cache_bytes(1)            # (so that the next call re-packs) ...
        r.render(rds[0])
        two = r.model_cache_info()["resident_bytes"]
        res["two_bytes"] = two
        r.set_model_cache_bytes(two)
        base = r.model_cache_info()
        fa = r.render(rds[0])[0]
        ia = r.model_cache_info()
        r.touch_model("progs/blob.mdl")       # blob is now used more recently than ball
        fb = r.render(rds[1])[0]              # needs m2: ball leaves, blob stays
        ib = r.model_cache_info()
        fc = r.render(rds[2])[0]              # needs ball: m2 (as big as ball) leaves, blob still fits
        ic = r.model_cache_info()
        fb2 = r.render(rds[1])[0]             # m2 again: reloaded from the host copy
        id_ = r.model_cache_info()
        res["lru"] = [fa, fb, fc, fb2]
        res["infos"] = [base, ia, ib, ic, id_]
        # a budget nothing fits: every batch still has its own models while it is drawn
        r.set_model_cache_bytes(1)
        res["tiny"] = [r.render(rd)[0] for rd in rds]
        res["info_tiny"] = r.model_cache_info()
        res["per_model"] = per_model
        r.shutdown()
        q.put(("ok", res))
    except Exception:  # noqa: BLE001
        import traceback
        q.put(("err", traceback.format_exc()))


def _run(libpath, basedir):
    views.write_file(basedir, "progs/m2.mdl", mdlgen.icosphere_mdl(seed=5))     # same shape as ball.mdl, another skin
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_worker, args=(q, libpath, basedir))
    p.start()
    kind, payload = q.get(timeout=600)
    p.join(timeout=30)
    if kind != "ok":
        raise RuntimeError(payload)
    return payload


def _check(libpath, assets_dir, oracle_available):
    r = _run(libpath, assets_dir)
    assert r["info0"]["loaded"] == 3 and r["info0"]["resident"] == 3 and r["info0"]["evictions"] == 0
    base, ia, ib, ic, id_ = r["infos"]
    assert ia["resident"] == 2 and ia["resident_bytes"] == r["two_bytes"] <= ia["budget_bytes"]
    # B: m2 in, ball (least recently used: blob was touched) out
    assert (ib["resident"], ib["uploads"] - ia["uploads"], ib["evictions"] - ia["evictions"]) == (2, 1, 1), (ia, ib)
    # C: ball in again, m2 out (ball + m2 do not fit, ball + blob do)
    assert (ic["resident"], ic["uploads"] - ib["uploads"], ic["evictions"] - ib["evictions"]) == (2, 1, 1), (ib, ic)
    assert ic["resident_bytes"] == r["two_bytes"]
    assert (id_["uploads"] - ic["uploads"], id_["evictions"] - ic["evictions"]) == (1, 1)
    assert id_["rebuilds"] - base["rebuilds"] == 4          # the new budget at A, then B, C and B again
    fa, fb, fc, fb2 = r["lru"]
    plain = r["plain"]
    for name, got, want in [("A", fa, plain[0]), ("B", fb, plain[1]), ("C", fc, plain[2]), ("B again", fb2, plain[1])]:
        assert np.array_equal(got, want), f"batch {name}: pixels changed with the model cache budget"
    for k in range(3):
        assert np.array_equal(r["tiny"][k], plain[k]), f"batch {k} with a one-byte budget"
    assert r["info_tiny"]["resident"] == 1 and r["info_tiny"]["resident_bytes"] > 1      # above the budget while drawn
    if oracle_available:
        from oracle import run_oracle
        for k, specs in enumerate(_batches()):
            o = run_oracle.render(assets_dir, "maps/full.bsp", W, H, specs)
            assert np.array_equal(o["fb"], plain[k]), f"batch {k} differs from the reference"
        assert (plain[0] != plain[1]).any()


def test_model_cache_lifetimes_sim(assets_dir, sim_lib, oracle_available):
    _check(sim_lib, assets_dir, oracle_available)


@pytest.mark.gpu
def test_model_cache_lifetimes_gpu(assets_dir, oracle_available):
    _check(api.LIBPATH, assets_dir, oracle_available)
