"""Link-level drop-in proof: ONE engine-side caller (tests/linkhost/engine_frames.c, compiled against the reference's own
client.h / ref.h / render.h / vid.h) linked against the reference renderer and against the product; both must leave the
same bytes in vid.buffer and d_pzbuffer for the same frame script (Mod_ForName -> R_NewMap -> R_BeginFrame ->
R_RenderView -> Draw_* -> R_EndFrame; client/cl_parse.c:271,294, client/cl_screen.c:1768-1807).  The binaries are
built where /root/reference exists (here) and travel with the snapshot; the product binary also runs the script through
GetRefAPI (1)."""
import os
import subprocess

import pytest

import rcutil

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "linkhost", "_build")
W, H, FRAMES = 320, 200, 12


def _ensure_built():
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "linkhost")])
    for b in ("linkhost_refsoft", "linkhost_refcuda", "linkhost_sim"):
        if not os.path.exists(os.path.join(BUILD, b)):
            pytest.skip(f"tests/linkhost/_build/{b} not built (needs /root/reference's headers)")


def _run(binary, assets_dir, out, env=None):
    info = rcutil.scene_info("small")
    x, y, z = info["spawn"][0]
    e = dict(os.environ)
    e.update(env or {})
    subprocess.run([os.path.join(BUILD, binary), assets_dir, "maps/small.bsp", str(W), str(H), str(FRAMES),
                    str(x), str(y), str(z + 30.0), out], check=True, env=e, timeout=600)
    with open(out, "rb") as f:
        return f.read()


def _imports(binary):
    """dynamic symbols the binary imports, and the shared objects it names"""
    nm = subprocess.run(["nm", "-D", "--undefined-only", os.path.join(BUILD, binary)], capture_output=True, text=True, check=True).stdout
    needed = subprocess.run(["readelf", "-d", os.path.join(BUILD, binary)], capture_output=True, text=True, check=True).stdout
    return set(l.split()[-1] for l in nm.splitlines() if l.strip()), needed


ENGINE_CALLS = {"Mod_ForName", "R_NewMap", "R_BeginFrame", "R_RenderView", "R_EndFrame", "Draw_Fill", "Draw_String",
                "Draw_Character", "Draw_TileClear", "Draw_Pic", "Draw_FadeScreen", "Draw_PicFromWad"}


def test_same_caller_links_against_both_renderers():
    _ensure_built()
    sa, na = _imports("linkhost_refsoft")
    sb, nb = _imports("linkhost_refcuda")
    assert ENGINE_CALLS <= sa and ENGINE_CALLS <= sb
    assert "librefsoft.so" in na and "libref_cuda.so" in nb and "librefsoft" not in nb


def test_linkhost_sim_matches_reference(assets_dir, tmp_path):
    """host logic of the render.h path (rectangles, rowbytes > width, recorded 2D commands) on the CPU"""
    _ensure_built()
    a = _run("linkhost_refsoft", assets_dir, str(tmp_path / "ref.bin"))
    b = _run("linkhost_sim", assets_dir, str(tmp_path / "sim.bin"))
    assert len(a) == len(b) and a == b


@pytest.mark.gpu
def test_linkhost_product_matches_reference(assets_dir, tmp_path):
    _ensure_built()
    a = _run("linkhost_refsoft", assets_dir, str(tmp_path / "ref.bin"))
    b = _run("linkhost_refcuda", assets_dir, str(tmp_path / "cuda.bin"))
    assert len(a) == len(b) and a == b
    c = _run("linkhost_refcuda", assets_dir, str(tmp_path / "cuda_refapi.bin"), env={"LINKHOST_REFAPI": "1"})
    assert c == a, "GetRefAPI (1) path differs"
