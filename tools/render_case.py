"""Renders one synthetic case through libref_cuda.so in THIS process (so that compute-sanitizer / ncu can wrap it)
and compares it with the oracle when oracle/_ref is present.  GPU box only.

    python tools/render_case.py --scene full --w 320 --h 240 --n 4 --entities 6 --particles 100 --sprites 3 --doors 2
"""
import argparse
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rcutil  # noqa: E402
from tochris_b200.api import RefCuda, status_names  # noqa: E402
from tochris_b200.scenes import views  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="small")
    ap.add_argument("--w", type=int, default=320)
    ap.add_argument("--h", type=int, default=200)
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--async-calls", type=int, default=0, help="also run this many pipelined RC_RenderViewsAsync calls")
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--device", action="store_true", help="outputs stay on the device (RC_RenderViewsDevice on a torch stream): the plain path, one k_draw launch over all views")
    ap.add_argument("--flush", action="store_true", help="D_FlushCaches before every repetition (every render rebuilds its surface cache blocks)")
    for k in ("entities", "particles", "sprites", "doors", "dlights", "underwater_every"):
        ap.add_argument("--" + k, type=int, default=0)
    a = ap.parse_args()

    base = tempfile.mkdtemp(prefix="rc_case_")
    rcutil.make_assets(base, names=[a.scene])
    extra = {k: getattr(a, k) for k in ("entities", "particles", "sprites", "doors", "dlights", "underwater_every") if getattr(a, k)}
    specs = rcutil.scene_views(a.scene, a.n, time=0.4, time_step=0.31, **extra) if a.scene != "box" else rcutil.scene_views("box", a.n)
    r = RefCuda(base, a.w, a.h)
    r.load_world(f"maps/{a.scene}.bsp")
    rds = views.build_refdefs(specs, a.w, a.h, model_resolver=r.model)
    if a.device:
        import torch
        dfb = torch.empty((a.n, a.h, a.w), dtype=torch.uint8, device="cuda")
        dz = torch.empty((a.n, a.h, a.w), dtype=torch.int16, device="cuda")
        stream = torch.cuda.Stream()
    for _ in range(a.reps):
        if a.flush:
            r.flush_caches()
        if a.device:
            with torch.cuda.stream(stream):
                st = r.render_device(rds, a.n, dfb.data_ptr(), dz.data_ptr(), stream.cuda_stream)
            torch.cuda.synchronize()
            fb, z = dfb.cpu().numpy(), dz.cpu().numpy()
        else:
            fb, z, st = r.render(rds)
    print("render_case: status", status_names(st), "launches", r.stats()["kernel_launches"])
    if a.async_calls:
        import numpy as np
        bufs = [np.zeros_like(fb) for _ in range(3)]
        for k in range(a.async_calls):
            if k >= 3:
                assert r.wait() == 0
                assert (bufs[k % 3] == fb).all()
            r.render_async(rds, a.n, bufs[k % 3])
        for k in range(min(3, a.async_calls)):
            assert r.wait() == 0
        print("render_case: pipelined calls match the blocking call")
    from oracle import refsoft, run_oracle  # noqa: E402
    if refsoft.available() and not a.no_oracle:
        o = run_oracle.render(base, f"maps/{a.scene}.bsp", a.w, a.h, specs)
        rcutil.assert_parity(o, dict(fb=fb, z=z, status=st), specs, a.w, a.h)
        print("render_case: bit-exact against the oracle (colour and z)")
    r.shutdown()


if __name__ == "__main__":
    main()
