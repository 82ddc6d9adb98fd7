"""Host link probe for the multi-GPU end-to-end numbers: every rank copies pinned host <-> device buffers at the same time
(torchrun, one rank per GPU) and reports its own bandwidth, alone and with all ranks active.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py"""
import os
import time

import torch
import torch.distributed as dist


def bw(dst, src, reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    return reps * src.numel() / (time.perf_counter() - t0) / 1e9


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << 30
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    bw(host, dev, 2)
    res = {}
    for name, (d, s) in {"d2h": (host, dev), "h2d": (dev, host)}.items():
        if world > 1:
            dist.barrier()
        res[name + "_all_ranks_busy"] = bw(d, s, 8)
    if world > 1:
        dist.barrier()
        for r in range(world):          # one rank at a time
            if r == rank:
                res["d2h_alone"] = bw(host, dev, 4)
            dist.barrier()
    out = [None] * world
    if world > 1:
        dist.all_gather_object(out, res)
    else:
        out = [res]
    if rank == 0:
        for r, o in enumerate(out):
            print("rank %d: " % r + ", ".join("%s %.1f GB/s" % (k, v) for k, v in o.items()))
        print("sum d2h with all ranks busy: %.1f GB/s" % sum(o["d2h_all_ranks_busy"] for o in out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
