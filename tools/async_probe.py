"""Where a pipelined e2e step spends its wall time (CPU enqueue vs waits).  GPU box only."""
import os, sys, time, tempfile
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rcutil
from tochris_b200.api import RefCuda
from tochris_b200.scenes import views
base = tempfile.mkdtemp(prefix="rc_probe_")
rcutil.make_assets(base, names=["multi"])
W, H, n = 640, 480, 64
specs = rcutil.scene_views("multi", n, time=1.0)
r = RefCuda(base, W, H)
r.load_world("maps/multi.bsp")
rds = views.build_refdefs(specs, W, H, model_resolver=r.model)
bufs = [torch.empty((n, H, W), dtype=torch.uint8, pin_memory=True).numpy() for _ in range(3)]
l2flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
side = torch.cuda.Stream()
for flush in (False, True, "l2", "l2side"):
    for rep in range(2):
        tq, tw = 0.0, 0.0
        torch.cuda.synchronize(); t00 = time.perf_counter()
        K = 20
        for k in range(K):
            if k >= 3:
                t0 = time.perf_counter(); r.wait(); tw += time.perf_counter() - t0
            if flush: r.flush_caches()
            if flush == "l2": l2flush.zero_()
            if flush == "l2side":
                with torch.cuda.stream(side): l2flush.zero_()
            t0 = time.perf_counter(); r.render_async(rds, n, bufs[k % 3]); tq += time.perf_counter() - t0
        for k in range(3):
            t0 = time.perf_counter(); r.wait(); tw += time.perf_counter() - t0
        tot = time.perf_counter() - t00
        print(f"flush={flush} rep={rep}: total {tot/K*1e3:.3f} ms/step, enqueue {tq/K*1e3:.3f}, wait {tw/K*1e3:.3f}")
t0 = time.perf_counter()
for k in range(20): r.render(rds, n, want_z=False, fb=bufs[0])
print(f"blocking: {(time.perf_counter()-t0)/20*1e3:.3f} ms/step")
# plain D2H copies of one step's frame buffers while device-output renders run on another stream: how much the
# render's memory traffic slows the PCIe copy (and the other way round)
fbd = torch.empty((n, H, W), dtype=torch.uint8, device="cuda"); zd = torch.empty((n, H, W), dtype=torch.int16, device="cuda")
hb = torch.empty((n, H, W), dtype=torch.uint8, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for load in (False, True):
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record(s1)
    if load:
        for k in range(40): r.render_device(rds, n, fbd.data_ptr(), zd.data_ptr(), s1.cuda_stream)
    r1.record(s1)
    with torch.cuda.stream(s2):
        for a, b in evs:
            a.record(s2); hb.copy_(fbd, non_blocking=True); b.record(s2)
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    print(f"copy alone={not load}: median {ts[10]:.3f} ms, max {ts[-1]:.3f}; 40 renders beside it {r0.elapsed_time(r1)/40:.3f} ms each")
