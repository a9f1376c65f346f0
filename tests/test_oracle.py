"""The oracle (unmodified reference compiled into oracle/_ref) against the committed golden
vectors, and the struct layouts the C-ABI mirrors."""
import ctypes as C
import os

import numpy as np
import pytest

import rcutil
from tochris_b200 import refdef


def test_struct_sizes_match_reference(oracle_available, assets_dir):
    """client/ref.h structs as the reference compiler lays them out (136/80/16/16/68 B)."""
    assert C.sizeof(refdef.refdef_t) == 136
    assert C.sizeof(refdef.entity_t) == 80
    assert C.sizeof(refdef.particle_t) == 16
    assert C.sizeof(refdef.dlight_t) == 16
    assert C.sizeof(refdef.lightstyle_t) == 68
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    from oracle import run_oracle  # noqa: F401  (sizes are checked inside a fresh process)
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_sizes, args=(q, assets_dir))
    p.start(); sizes = q.get(timeout=120); p.join()
    assert sizes[:5] == [136, 80, 16, 16, 68]
    assert sizes[5:9] == [56, 88, 24, 32]          # edge_t, surf_t, espan_t, finalvert_t


def _sizes(q, base):
    from oracle.refsoft import RefSoft
    q.put(RefSoft(base, 320, 200).struct_sizes())


@pytest.mark.parametrize("name", ["smoke_small_320x200", "full_320x240"])
def test_oracle_reproduces_golden(oracle_available, assets_dir, name):
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    from oracle import run_oracle
    g = np.load(os.path.join(rcutil.ROOT, "tests", "golden", name + ".npz"))
    scene = str(g["scene"]); w, h, n = int(g["w"]), int(g["h"]), int(g["n"])
    bsp = np.frombuffer(open(os.path.join(assets_dir, "id1", "maps", scene + ".bsp"), "rb").read(), dtype=np.uint8)
    assert rcutil.sha(bsp) == str(g["bsp_sha"]), "the procedural BSP generator changed: regenerate the goldens"
    specs = rcutil.scene_views(scene, n, time=float(g["t0"]), time_step=float(g["dt"]))
    o = run_oracle.render(assets_dir, f"maps/{scene}.bsp", w, h, specs)
    assert rcutil.sha(o["fb"]) == str(g["fb_sha"])
    assert rcutil.sha(o["z"]) == str(g["z_sha"])
    k = g["fb"].shape[0]
    assert (o["fb"][:k] == g["fb"]).all() and (o["z"][:k] == g["z"]).all()


def test_alias_model_memory_image_matches_the_reference(assets_dir, sim_lib, oracle_available):
    """Clipped alias triangles step off the skin picture (d_polyse.c:385-447) and the reference then reads its own model
    memory around the skin; ref_cuda keeps a byte image of the block Mod_LoadAliasModel assembles (hunk headers, skin
    descriptors, st vertices, triangles, frames) so that those reads are exact.  The image must equal the reference's
    memory byte for byte (Mod_Extradata), for a model with frame groups and one with a skin group."""
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    import ctypes as C
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_image_worker, args=(q, assets_dir, sim_lib))
    p.start()
    res = q.get(timeout=300)
    p.join(timeout=30)
    assert res == [("progs/ball.mdl", 0), ("progs/blob.mdl", 0)], res


def _image_worker(q, base, sim_lib):
    try:
        import ctypes as C
        import numpy as np
        from tochris_b200.api import RefCuda
        from oracle.refsoft import RefSoft
        r = RefCuda(base, 320, 200, libpath=sim_lib)
        r.load_world("maps/small.bsp")
        o = RefSoft(base, 320, 200)
        o.load_world("maps/small.bsp")
        o.lib.Mod_Extradata.restype = C.c_void_p
        o.lib.Mod_Extradata.argtypes = [C.c_void_p]
        r.lib.RC_DebugAliasImage.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        out = []
        for name in ("progs/ball.mdl", "progs/blob.mdl"):
            m = r.model(name)
            n = r.lib.RC_DebugAliasImage(m, None, 0)
            buf = (C.c_ubyte * n)()
            r.lib.RC_DebugAliasImage(m, buf, n)
            ref = np.frombuffer((C.c_ubyte * n).from_address(o.lib.Mod_Extradata(o.model(name))), dtype=np.uint8)
            out.append((name, int((np.frombuffer(buf, dtype=np.uint8) != ref).sum()) if n > 1000 else -1))
        q.put(out)
    except Exception:  # noqa: BLE001
        import traceback
        q.put(traceback.format_exc())
