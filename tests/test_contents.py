"""SURVEY.md 8(f) rank 4, first piece: what the client asks before it fills refdef_t.flags (cl_view.c:786-791): the contents
of the leaf that holds the view origin, against the reference's Mod_PointInLeaf (r_model.c:83)."""
import numpy as np

import rcutil


import pytest


def test_point_contents_match_reference(assets_dir, sim_lib, oracle_available):
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    _contents_case(assets_dir, sim_lib)


@pytest.mark.gpu
def test_point_contents_match_reference_gpu(assets_dir, oracle_available):
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    from tochris_b200 import api
    _contents_case(assets_dir, api.LIBPATH)


def _contents_case(assets_dir, sim_lib):
    from oracle.refsoft import RefSoft
    from tochris_b200.api import RefCuda
    import multiprocessing as mp

    info = rcutil.scene_info("full")
    rng = np.random.default_rng(3)
    lo = np.array(info["mins"], dtype=np.float64) - 64 if "mins" in info else np.array([-1200.0, -1200.0, -300.0])
    hi = np.array(info["maxs"], dtype=np.float64) + 64 if "maxs" in info else np.array([1200.0, 1200.0, 400.0])
    pts = rng.uniform(lo, hi, size=(2000, 3)).astype(np.float32)

    def oracle(q):
        r = RefSoft(assets_dir, 320, 200, heap_mb=96)
        r.load_world("maps/full.bsp")
        q.put([r.point_contents(p) for p in pts])
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    p = ctx.Process(target=oracle, args=(q,))
    p.start()
    want = q.get(timeout=120)
    p.join(timeout=30)
    r = RefCuda(assets_dir, 320, 200, libpath=sim_lib)
    r.load_world("maps/full.bsp")
    got = [r.point_contents(p) for p in pts]
    assert got == want
    assert len(set(got)) >= 2                      # empty and solid at least (the generated maps have no water volumes)
    for pnt, c in zip(pts[:200], got[:200]):
        assert r.view_flags_for_origin(pnt) == (2 if c <= -3 else 0)      # RDF_UNDERWATER
    r.shutdown()
