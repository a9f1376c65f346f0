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
