"""The N>1 path on the CPU: world_size-2 gloo processes each render their shard of a batch
(host-simulation build of the C-ABI, test infrastructure) and gather the framebuffers to rank 0;
the result must equal the single-process render of the whole batch byte for byte."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import rcutil
from tochris_b200 import shard

W, H, N = 160, 120, 9          # odd batch: the shards differ in size


def test_shard_bounds_cover_every_view_once():
    for n in (1, 7, 64, 65536):
        for world in (1, 2, 3, 8):
            b = [shard.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, base, sim_lib, outdir):
    sys.path.insert(0, os.path.join(rcutil.ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tochris_b200.api import RefCuda
    from tochris_b200.scenes import views
    specs = rcutil.scene_views("small", N, time=0.3, time_step=0.2)
    lo, hi = shard.shard_bounds(N, rank, world)
    r = RefCuda(base, W, H, libpath=sim_lib)
    r.load_world("maps/small.bsp")
    rds = views.build_refdefs(specs[lo:hi], W, H, model_resolver=r.model)
    fb, _, st = r.render(rds, want_z=False)
    assert st == 0
    out = shard.gather_frames(torch.from_numpy(fb), N, rank, world)
    if rank == 0:
        np.save(os.path.join(outdir, "gathered.npy"), out.numpy())
    else:
        assert out is None
    r.shutdown()
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one(assets_dir, sim_lib, tmp_path):
    specs = rcutil.scene_views("small", N, time=0.3, time_step=0.2)
    whole = rcutil.render_rc(sim_lib, assets_dir, "maps/small.bsp", W, H, specs)
    mp.spawn(_worker, args=(2, _free_port(), assets_dir, sim_lib, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered.npy"))
    assert got.shape == whole["fb"].shape
    assert (got == whole["fb"]).all()


# ---- the shipped N>1 data plane on real GPUs: CUDA-IPC mirror into rank 0's gather buffer (bench.py --gpus N path) ----
GW, GH, GN = 320, 240, 24          # views per rank


def _gpu_worker(rank, world, base, outdir, qs):
    try:
        sys.path.insert(0, os.path.join(rcutil.ROOT, "tests"))
        torch.cuda.set_device(rank)
        from tochris_b200.api import RefCuda
        from tochris_b200.scenes import views
        specs = rcutil.scene_views("multi", GN * world, time=0.3, time_step=0.2)
        lo, hi = shard.shard_bounds(GN * world, rank, world)
        r = RefCuda(base, GW, GH, device=rank)
        r.load_world("maps/multi.bsp")
        rds = views.build_refdefs(specs[lo:hi], GW, GH, model_resolver=r.model)
        n, px = hi - lo, GW * GH
        dev = torch.device("cuda", rank)
        fb = torch.zeros((n, GH, GW), dtype=torch.uint8, device=dev)
        zb = torch.zeros((n, GH, GW), dtype=torch.int16, device=dev)
        if rank == 0:
            root = r.device_alloc(world * GN * px)
            h = r.ipc_export(root)
            for q in qs[1:]:
                q.put(h)
        else:
            root = r.ipc_import(qs[rank].get(timeout=120))
        stream = torch.cuda.Stream(dev)
        for groups in (2, 5):                       # RC_SetMirrorGroups: the copies start after each finished group
            r.set_mirror_groups(groups)
            r.set_mirror_output(root + lo * px)
            for _ in range(3):                      # repeated calls: the CUDA-graph replay path as well as the first capture
                st = r.render_device(rds, n, fb.data_ptr(), zb.data_ptr(), stream.cuda_stream)
                assert st == 0
            stream.synchronize()
        # deferred mirror (RC_SetMirrorDeferred / RC_MirrorFence): the copy of a batch overlaps the next batch; two frame
        # buffers in rotation, then one more call into a buffer whose copy is still in flight (it must wait for it)
        fb2 = torch.zeros_like(fb)
        r.set_mirror_deferred(True)
        for k in range(5):
            st = r.render_device(rds, n, (fb2 if k & 1 else fb).data_ptr(), zb.data_ptr(), stream.cuda_stream)
            assert st == 0
        st = r.render_device(rds, n, fb.data_ptr(), zb.data_ptr(), stream.cuda_stream)
        assert st == 0
        r.mirror_fence(stream.cuda_stream)
        stream.synchronize()
        assert torch.equal(fb, fb2)
        r.set_mirror_deferred(False)
        r.set_mirror_output(0)
        np.save(os.path.join(outdir, f"local{rank}.npy"), fb.cpu().numpy())
        qs[0].put(("done", rank)) if rank else None
        if rank == 0:
            for _ in range(world - 1):
                assert qs[0].get(timeout=300)[0] == "done"
            import ctypes as C
            arrived = torch.empty(world * GN * px, dtype=torch.uint8, device=dev)
            assert C.CDLL("libcudart.so").cudaMemcpy(C.c_void_p(arrived.data_ptr()), C.c_void_p(root), C.c_size_t(world * GN * px), 3) == 0
            np.save(os.path.join(outdir, "gathered.npy"), arrived.cpu().numpy().reshape(world * GN, GH, GW))
            for q in qs[1:]:
                q.put("bye")
            r.device_free(root)
        else:
            assert qs[rank].get(timeout=300) == "bye"   # keep the imported mapping alive until rank 0 has read the buffer
            r.ipc_close(root)
        r.shutdown()
    except Exception:  # noqa: BLE001
        import traceback
        with open(os.path.join(outdir, f"error{rank}.txt"), "w") as f:
            f.write(traceback.format_exc())
        raise


@pytest.mark.gpu
def test_two_gpus_mirror_into_rank0_matches_oracle(assets_dir, oracle_available, tmp_path):
    """RC_DeviceAlloc / RC_IpcExport / RC_IpcImport / RC_SetMirrorOutput / RC_SetMirrorGroups on two GPUs: what arrives in
    rank 0's gather buffer over NVLink equals the reference's rendering of the whole batch."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    qs = [ctx.Queue() for _ in range(2)]
    procs = [ctx.Process(target=_gpu_worker, args=(rk, 2, assets_dir, str(tmp_path), qs)) for rk in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=900)
    errs = [open(os.path.join(str(tmp_path), f)).read() for f in os.listdir(str(tmp_path)) if f.startswith("error")]
    assert not errs and all(p.exitcode == 0 for p in procs), "\n".join(errs)
    got = np.load(os.path.join(str(tmp_path), "gathered.npy"))
    for rk in range(2):
        assert np.array_equal(got[rk * GN:(rk + 1) * GN], np.load(os.path.join(str(tmp_path), f"local{rk}.npy")))
    if oracle_available:
        from oracle import run_oracle
        specs = rcutil.scene_views("multi", GN * 2, time=0.3, time_step=0.2)
        o = run_oracle.render(assets_dir, "maps/multi.bsp", GW, GH, specs, want_z=False)
        assert np.array_equal(got, o["fb"])
