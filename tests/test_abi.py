"""The product library loads, exports every symbol include/ref_cuda.h declares, and fails
loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

import rcutil
from tochris_b200 import api


def _declared_symbols():
    src = open(os.path.join(rcutil.ROOT, "include", "ref_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set()
    for m in re.finditer(r"^\s*(?:[A-Za-z_][\w\s\*]*?)\b([A-Z][A-Za-z_]*_?[A-Za-z]\w*)\s*\(", src, flags=re.M):
        names.add(m.group(1))
    names = {n for n in names if re.match(r"^(RC_|R_|Mod_|GetRefAPI)", n)}
    return sorted(names)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(api.LIBPATH):
        subprocess.check_call([sys.executable, "-c", "import __graft_entry__ as g; g.build()"], cwd=rcutil.ROOT)
    lib = C.CDLL(api.LIBPATH)
    syms = _declared_symbols()
    assert len(syms) >= 20, syms
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ref_cuda.h but not exported"


def test_sim_library_exports_the_same_abi(sim_lib):
    lib = C.CDLL(sim_lib)
    for s in _declared_symbols():
        assert hasattr(lib, s), s


def test_no_cpu_fallback(assets_dir):
    """Without a GPU RC_Init must fail with a clear message instead of rendering on the CPU."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import torch\n"
        "from tochris_b200.api import RefCuda\n"
        "try:\n"
        "    RefCuda(%r, 320, 200)\n"
        "    print('INIT_OK')\n"
        "except RuntimeError as e:\n"
        "    print('INIT_FAILED', e)\n" % (rcutil.ROOT, assets_dir))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300).stdout
    import torch
    if torch.cuda.is_available():
        assert "INIT_OK" in out
    else:
        assert "INIT_FAILED" in out and "CUDA" in out


def test_loader_rejects_malformed_bsp(sim_lib, assets_dir, tmp_path):
    """Sys_Error cases of r_model.c become error codes."""
    import shutil
    base = str(tmp_path)
    shutil.copytree(os.path.join(assets_dir, "id1"), os.path.join(base, "id1"))
    good = open(os.path.join(base, "id1", "maps", "box.bsp"), "rb").read()
    bad = {
        "version": b"\x1e\x00\x00\x00" + good[4:],
        "truncated": good[: len(good) // 2],
        "lump": good[:8] + b"\xff\xff\xff\x7f" + good[12:],
    }
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from tochris_b200.api import RefCuda\n"
        "r = RefCuda(%r, 320, 200, libpath=%r)\n"
        "for n in %r:\n"
        "    try:\n"
        "        r.load_world('maps/bad_%%s.bsp' %% n); print(n, 'LOADED')\n"
        "    except RuntimeError as e:\n"
        "        print(n, 'REJECTED', e)\n" % (rcutil.ROOT, base, sim_lib, list(bad)))
    for n, data in bad.items():
        open(os.path.join(base, "id1", "maps", f"bad_{n}.bsp"), "wb").write(data)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300).stdout
    for n in bad:
        assert f"{n} REJECTED" in out, out
