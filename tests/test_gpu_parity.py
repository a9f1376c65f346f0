"""GPU parity tests proper: libref_cuda.so (CUDA, through the C-ABI) against the oracle
(the unmodified reference compiled into oracle/_ref) on the same seeded inputs, bit-exact
for the 8-bit framebuffer and the int16 z-buffer."""
import os

import numpy as np
import pytest

import rcutil
from tochris_b200 import api

pytestmark = pytest.mark.gpu

CASES = [
    # scene, w, h, views, time, time_step, extras
    ("box", 320, 200, 128, 1.0, 0.0, {}),          # BASELINE config C1: timerefresh sweep
    ("small", 320, 200, 16, 0.4, 0.31, {}),
    ("multi", 640, 480, 64, 1.0, 0.0, {}),         # BASELINE config C2
    ("full", 320, 240, 64, 0.37, 0.173, {}),       # sky + water + animated textures, time-varying lights
    ("multi", 160, 120, 512, 0.0, 0.01, {}),       # BASELINE config C4 (sample)
    ("full", 1280, 720, 6, 2.5, 0.5, dict(entities=64, particles=2048)),   # BASELINE config C3 (sample)
    ("multi", 320, 240, 48, 0.2, 0.05, dict(dlights=6)),                     # dynamic lights
    ("full", 640, 480, 12, 0.9, 0.11, dict(dlights=2, underwater_every=2, entities=12, particles=200)),  # warp + everything
    ("full", 320, 240, 64, 0.37, 0.173, dict(doors=2, dlights=3)),           # brush entities: clipped to the BSP, rotated, in one leaf
    ("full", 640, 480, 16, 1.1, 0.2, dict(doors=2, entities=10, particles=100, underwater_every=4)),
    ("full", 320, 240, 64, 0.37, 0.173, dict(skybox="testsky")),             # loadsky: the six sky-box faces
    ("full", 640, 480, 12, 1.1, 0.3, dict(skybox="testsky", doors=2, entities=6, dlights=2, underwater_every=5)),
    ("multi", 320, 240, 48, 0.3, 0.21, dict(sprites=12)),                     # sprite entities
    ("full", 640, 480, 12, 0.9, 0.17, dict(sprites=10, entities=8, particles=100, doors=2, underwater_every=5)),
    # BASELINE config C3 at full size: 256 alias entities + 8192 particles per view, 1280x720, batch 32
    ("full", 1280, 720, 32, 2.5, 0.5, dict(entities=256, particles=8192)),
    # BASELINE config C5 substitute (3840x2160 is not representable; 1920x1080 with the raised constants):
    # dynamic lights, under-water warp every third view, sky, water
    ("full", 1920, 1080, 6, 1.7, 0.4, dict(dlights=8, underwater_every=3, big=True)),
]


@pytest.mark.parametrize("scene,w,h,n,t0,dt,extra", CASES)
def test_matches_oracle(assets_dir, oracle_available, scene, w, h, n, t0, dt, extra):
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    from oracle import run_oracle
    kw = dict(extra)
    skybox = kw.pop("skybox", None)
    cvars = {"r_skyname": skybox} if skybox else {}
    if kw.pop("big", False):
        cvars.update(r_maxsurfs=4000.0, r_maxedges=12000.0)     # same pools on both sides (r_main.c:265-300)
    if scene != "box":
        kw.update(time=t0, time_step=dt)
    specs = rcutil.scene_views(scene, n, **kw)
    o = run_oracle.render(assets_dir, f"maps/{scene}.bsp", w, h, specs, cvars=cvars)
    r = rcutil.render_rc(api.LIBPATH, assets_dir, f"maps/{scene}.bsp", w, h, specs, cvars=cvars)
    assert r["stats"]["kernel_launches"] > 0
    rcutil.assert_parity(o, r, specs, w, h, allow_alias_range=False)


def test_golden_fixtures(assets_dir):
    """Committed golden vectors (made by tests/golden/make_golden.py from the oracle)."""
    gdir = os.path.join(rcutil.ROOT, "tests", "golden")
    for fn in sorted(os.listdir(gdir)):
        if not fn.endswith(".npz"):
            continue
        g = np.load(os.path.join(gdir, fn), allow_pickle=False)
        scene = str(g["scene"]); w, h, n = int(g["w"]), int(g["h"]), int(g["n"])
        extra = {k: int(g[k]) for k in ("entities", "particles") if k in g}
        specs = rcutil.scene_views(scene, n, time=float(g["t0"]), time_step=float(g["dt"]), **extra) if scene != "box" else rcutil.scene_views(scene, n)
        r = rcutil.render_rc(api.LIBPATH, assets_dir, f"maps/{scene}.bsp", w, h, specs)
        assert rcutil.sha(r["fb"]) == str(g["fb_sha"]), fn
        assert rcutil.sha(r["z"]) == str(g["z_sha"]), fn
        if "fb" in g:
            k = g["fb"].shape[0]
            assert (r["fb"][:k] == g["fb"]).all() and (r["z"][:k] == g["z"]).all(), fn


def test_batch_invariance(assets_dir):
    """A view's pixels do not depend on what else is in the batch, on the batch order, or on
    the surface cache being warm (purity of R_RenderView)."""
    specs = rcutil.scene_views("multi", 24, time=0.2, time_step=0.07)
    a = rcutil.render_rc(api.LIBPATH, assets_dir, "maps/multi.bsp", 320, 240, specs)
    b = rcutil.render_rc(api.LIBPATH, assets_dir, "maps/multi.bsp", 320, 240, specs[::-1])
    assert (a["fb"] == b["fb"][::-1]).all() and (a["z"] == b["z"][::-1]).all()
    c = rcutil.render_rc(api.LIBPATH, assets_dir, "maps/multi.bsp", 320, 240, specs[5:6])
    assert (c["fb"][0] == a["fb"][5]).all()


def _c4_worker(q, libpath, basedir, reps):
    try:
        import numpy as np
        from tochris_b200.api import RefCuda
        from tochris_b200.scenes import views
        base = rcutil.scene_views("multi", 512, time=0.0, time_step=0.01)
        r = RefCuda(basedir, 160, 120, libpath=libpath, max_views_in_flight=8192)      # several pipeline chunks
        r.load_world("maps/multi.bsp")
        rds = views.build_refdefs(base * reps, 160, 120, model_resolver=r.model)
        fb, z, st = r.render(rds)
        fb = fb.reshape(reps, 512, 120, 160); z = z.reshape(reps, 512, 120, 160)
        same = bool((fb == fb[0:1]).all() and (z == z[0:1]).all())
        q.put(("ok", dict(status=st, same=same, fb=fb[0].copy(), z=z[0].copy(), n=len(rds))))
        r.shutdown()
    except Exception:  # noqa: BLE001
        import traceback
        q.put(("err", traceback.format_exc()))


def test_c4_full_batch_properties(assets_dir, oracle_available):
    """BASELINE config C4 at full size: 65,536 independent 160x120 views in one call (eight pipeline chunks).
    The oracle cannot render that many in a test, so the batch is 128 repetitions of 512 distinct cameras:
    every repetition must be byte-identical to the first (a view's pixels do not depend on its position in
    the batch or on the chunk it lands in), and the first 512 must match the reference."""
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    import multiprocessing as mp
    from oracle import run_oracle
    base = rcutil.scene_views("multi", 512, time=0.0, time_step=0.01)
    o = run_oracle.render(assets_dir, "maps/multi.bsp", 160, 120, base)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_c4_worker, args=(q, api.LIBPATH, assets_dir, 128))
    p.start()
    kind, r = q.get(timeout=1800)
    p.join(timeout=60)
    assert kind == "ok", r
    assert r["n"] == 65536 and r["status"] == 0
    assert (r["fb"] == o["fb"]).all() and (r["z"] == o["z"]).all()
    assert r["same"]


def _parts_worker(q, libpath, basedir, reps):
    try:
        import torch
        from tochris_b200.api import RefCuda
        from tochris_b200.scenes import views
        base = rcutil.scene_views("full", 256, time=0.3, time_step=0.02, entities=3, particles=40, underwater_every=37)
        r = RefCuda(basedir, 160, 120, libpath=libpath)
        r.load_world("maps/full.bsp")
        rds = views.build_refdefs(base * reps, 160, 120, model_resolver=r.model)
        n = len(rds)
        fb = torch.zeros((n, 120, 160), dtype=torch.uint8, device="cuda")
        z = torch.zeros((n, 120, 160), dtype=torch.int16, device="cuda")
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            st = r.render_device(rds, n, fb.data_ptr(), z.data_ptr(), stream.cuda_stream)      # n >= 2048: drawn in parts
        torch.cuda.synchronize()
        launches_parts = r.stats()["kernel_launches"]
        fbh = fb.cpu().numpy().reshape(reps, 256, 120, 160); zh = z.cpu().numpy().reshape(reps, 256, 120, 160)
        same = bool((fbh == fbh[0:1]).all() and (zh == zh[0:1]).all())
        # once more: the first call found the alias span queue empty (its triangles drew themselves) and reported what it
        # would have asked of it; this one goes through the queue
        fb2 = torch.zeros_like(fb); z2 = torch.zeros_like(z)
        with torch.cuda.stream(stream):
            st |= r.render_device(rds, n, fb2.data_ptr(), z2.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        same = same and bool(torch.equal(fb, fb2) and torch.equal(z, z2))
        q.put(("ok", dict(status=st, same=same, fb=fbh[0].copy(), z=zh[0].copy(), n=n, launches=launches_parts)))
        r.shutdown()
    except Exception:  # noqa: BLE001
        import traceback
        q.put(("err", traceback.format_exc()))


def test_large_device_batches_drawn_in_parts(assets_dir, oracle_available):
    """RC_RenderViewsDevice draws a batch of 2048 views or more in parts (the host sets up part k+1 while the device draws
    part k; status words accumulate, one wait at the end): 9 x 256 cameras with entities, particles and
    under-water views -- every repetition byte-identical to the first, the first equal to the reference."""
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    import multiprocessing as mp
    from oracle import run_oracle
    base = rcutil.scene_views("full", 256, time=0.3, time_step=0.02, entities=3, particles=40, underwater_every=37)
    o = run_oracle.render(assets_dir, "maps/full.bsp", 160, 120, base)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_parts_worker, args=(q, api.LIBPATH, assets_dir, 9))
    p.start()
    kind, r = q.get(timeout=1800)
    p.join(timeout=60)
    assert kind == "ok", r
    assert r["n"] == 2304 and r["status"] & ~rcutil.RC_STAT_ALIAS_RANGE == 0, api.status_names(r["status"])
    rcutil.assert_parity(o, dict(fb=r["fb"], z=r["z"], status=r["status"]), base, 160, 120)
    assert r["same"]


@pytest.mark.gpu
def test_async_pipeline_matches_blocking_call(assets_dir):
    """RC_RenderViewsAsync / RC_WaitViews (up to three batches outstanding, copies overlapping the next batches' kernels)
    deliver the same bytes as the blocking RC_RenderViews, in order."""
    import numpy as np
    from tochris_b200.api import RefCuda
    from tochris_b200.scenes import views
    r = RefCuda(assets_dir, 320, 240)
    r.load_world("maps/multi.bsp")
    batches = []
    for b in range(7):
        specs = rcutil.scene_views("multi", 16, time=0.5 + b, time_step=0.07)
        batches.append(views.build_refdefs(specs, 320, 240, model_resolver=r.model))
    want = []
    for rds in batches:
        fb, z, st = r.render(rds)
        assert st == 0
        want.append((fb.copy(), z.copy()))
    bufs = [(np.zeros((16, 240, 320), np.uint8), np.zeros((16, 240, 320), np.int16)) for _ in range(3)]
    got = []
    for k, rds in enumerate(batches):
        if k >= 3:
            assert r.wait() == 0
            got.append((bufs[k % 3][0].copy(), bufs[k % 3][1].copy()))
        r.render_async(rds, 16, bufs[k % 3][0], bufs[k % 3][1])
        if k == 2:
            with pytest.raises(RuntimeError):      # a fourth outstanding call is refused, nothing is enqueued
                r.render_async(rds, 16, bufs[0][0], bufs[0][1])
    for k in range(4, 7):
        assert r.wait() == 0
        got.append((bufs[k % 3][0].copy(), bufs[k % 3][1].copy()))
    with pytest.raises(RuntimeError):
        r.wait()                       # nothing outstanding
    for (wf, wz), (gf, gz) in zip(want, got):
        assert np.array_equal(wf, gf) and np.array_equal(wz, gz)
    r.shutdown()


@pytest.mark.gpu
def test_sub_rectangles_match_oracle_gpu(assets_dir, oracle_available):
    """Sub-rectangle views on the device: view-rectangle edges inside 8-pixel cells (k_ends' end-of-row ownership),
    and a width that is not a multiple of 8 (k_draw hands every cell to the general code)."""
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    from oracle import run_oracle
    import test_hostsim_parity as hp
    specs = hp._rect_specs(20)
    o = run_oracle.render(assets_dir, "maps/full.bsp", 320, 200, specs)
    r = rcutil.render_rc(api.LIBPATH, assets_dir, "maps/full.bsp", 320, 200, specs)
    assert r["status"] == 0
    hp.check_rect_parity(o, r, specs)
    specs = rcutil.scene_views("multi", 6, time=0.3, time_step=0.2)          # 324 is not a multiple of 8
    o = run_oracle.render(assets_dir, "maps/multi.bsp", 324, 202, specs)
    r = rcutil.render_rc(api.LIBPATH, assets_dir, "maps/multi.bsp", 324, 202, specs)
    rcutil.assert_parity(o, r, specs, 324, 202, allow_alias_range=False)


FUZZ = [
    # scene, w, h, views, seed time, step -- shapes that stress the draw stage's cell logic: widths that are 8 but not
    # 16 aligned, odd heights, very narrow and very wide frames, many short spans (full scene from far away)
    ("full", 336, 243, 96, 0.11, 0.37),
    ("full", 200, 151, 128, 3.3, 0.29),
    ("multi", 1016, 571, 24, 0.7, 0.41),
    ("full", 72, 48, 256, 1.9, 0.13),
    ("multi", 1912, 300, 8, 2.2, 0.5),
    ("full", 644, 483, 24, 0.05, 0.77),          # 644 = 4 mod 8: every cell through the general code
]


@pytest.mark.parametrize("scene,w,h,n,t0,dt", FUZZ)
def test_fuzz_shapes_match_oracle(assets_dir, oracle_available, scene, w, h, n, t0, dt):
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    from oracle import run_oracle
    specs = rcutil.scene_views(scene, n, time=t0, time_step=dt)
    o = run_oracle.render(assets_dir, f"maps/{scene}.bsp", w, h, specs)
    r = rcutil.render_rc(api.LIBPATH, assets_dir, f"maps/{scene}.bsp", w, h, specs)
    rcutil.assert_parity(o, r, specs, w, h, allow_alias_range=False)


@pytest.mark.parametrize("scene,w,h,n,kw", [("multi", 320, 240, 6, {}), ("multi", 640, 480, 4, dict(dlights=3)), ("full", 400, 300, 6, {})])
def test_stage_dumps_match_oracle_gpu(assets_dir, oracle_available, scene, w, h, n, kw):
    """Stage-level differential test on the REAL kernels (k_front's edges and surfaces, k_scan's spans) through RC_DebugRead,
    against the dumps taken inside the reference -- not only on the host simulator."""
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    import test_hostsim_parity as hp
    hp.check_stage_dumps(api.LIBPATH, assets_dir, scene, w, h, n, **kw)


@pytest.mark.gpu
def test_async_stream_with_a_tiny_surface_cache_never_delivers_wrong_pixels(assets_dir):
    """A surface-cache pool that holds one batch's blocks but not two: in a pipelined stream (RC_RenderViewsAsync) the second
    batch runs out of pool.  Every batch must either come back clean and correct or carry RC_STAT_CACHE_POOL -- including
    the batches queued BEHIND the one that overflowed, which meet keys whose creator got no block --, and the blocking
    re-render the caller owes a flagged batch must be clean (the overflow schedules D_FlushCaches)."""
    import numpy as np
    from tochris_b200.api import RefCuda
    from tochris_b200.scenes import views
    W, H, n = 320, 240, 12
    sets = [rcutil.scene_views("multi", n, start=1 + 40 * k, time=0.5 + k) for k in range(2)]
    r = RefCuda(assets_dir, W, H)
    r.load_world("maps/multi.bsp")
    want, texels = [], []
    r.set_profiling(True)
    for sp in sets:
        r.flush_caches()
        fb, _, st = r.render(views.build_refdefs(sp, W, H, model_resolver=r.model), want_z=False)
        assert st == 0
        want.append(fb.copy()); texels.append(r.stats()["cache_texels"])
    r.set_profiling(False)
    r.shutdown()
    pool = 6 * 256 * 256 + int(max(texels) * 1.25) + 4096            # the sky-box pictures + one batch's blocks and a bit
    assert pool < 6 * 256 * 256 + sum(texels)
    r = RefCuda(assets_dir, W, H, surfcache_bytes=pool)
    r.load_world("maps/multi.bsp")
    rds = [views.build_refdefs(sp, W, H, model_resolver=r.model) for sp in sets]
    bufs = [np.zeros((n, H, W), np.uint8) for _ in range(3)]
    order = [0, 1, 0, 1, 0, 1, 1, 0]
    flagged = clean = 0
    pending = []
    for k, which in enumerate(order + [None] * 3):
        if len(pending) == 3 or which is None:
            if not pending:
                break
            slot, w_ = pending.pop(0)
            st = r.wait()
            if st == 0:
                assert np.array_equal(bufs[slot], want[w_]), "a batch with a clean status has wrong pixels"
                clean += 1
            else:
                assert st == 32, api.status_names(st)            # RC_STAT_CACHE_POOL and nothing else
                flagged += 1
        if which is not None:
            slot = k % 3
            r.render_async(rds[which], n, bufs[slot])
            pending.append((slot, which))
    assert flagged >= 1 and clean >= 1, (flagged, clean)
    for which in (1, 0):                                             # the blocking re-render: flushed first, then clean
        fb, _, st = r.render(rds[which], want_z=False)
        assert st == 0 and np.array_equal(fb, want[which])
    r.shutdown()
