"""How much device->host traffic the box takes when N ranks copy at once (the floor under e2e at N GPUs).  Launch with
torchrun --nproc-per-node N.  Each rank copies `mb` MB from its GPU to its own pinned buffer `reps` times; rank 0 prints
per-rank and aggregate GB/s for: all ranks at once / ranks one after the other / write-combined pinned memory."""
import os, sys, time, ctypes as C
import torch, torch.distributed as dist

def main():
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mb = 19.66; n = int(mb * 1e6); reps = 30
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    rt = C.CDLL("libcudart.so")
    out = {}
    for kind in ("pinned", "write_combined"):
        p = C.c_void_p()
        assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), 0 if kind == "pinned" else 4) == 0
        s = torch.cuda.Stream()
        def run(sync_each=False):
            torch.cuda.synchronize()
            if world > 1: dist.barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                rt.cudaMemcpyAsync(p, C.c_void_p(src.data_ptr()), C.c_size_t(n), 2, C.c_void_p(s.cuda_stream))
            s.synchronize()
            return time.perf_counter() - t0
        run(); dt = run()
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[kind + "_all_at_once"] = dict(per_rank_gbs=round(n * reps / float(t[0]) / 1e9, 1), aggregate_gbs=round(world * n * reps / float(t[0]) / 1e9, 1), ms_per_copy=round(float(t[0]) / reps * 1e3, 3))
        # one rank at a time
        alone = []
        for q in range(world):
            if world > 1: dist.barrier()
            if q == rank:
                torch.cuda.synchronize(); t0 = time.perf_counter()
                for _ in range(reps):
                    rt.cudaMemcpyAsync(p, C.c_void_p(src.data_ptr()), C.c_size_t(n), 2, C.c_void_p(s.cuda_stream))
                s.synchronize(); alone.append(n * reps / (time.perf_counter() - t0) / 1e9)
            if world > 1: dist.barrier()
        t = torch.tensor([alone[0]], device="cuda", dtype=torch.float64)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MIN)
        out[kind + "_alone_min_gbs"] = round(float(t[0]), 1)
        rt.cudaFreeHost(p)
    if rank == 0:
        import json
        print(json.dumps(dict(world=world, mb_per_copy=mb, **out)))
    if world > 1: dist.destroy_process_group()

if __name__ == "__main__":
    main()
