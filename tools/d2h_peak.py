#!/usr/bin/env python
"""What the host side can take from N GPUs at once: concurrent device-to-host copies into
page-locked memory, one process per GPU (torchrun), 1 / 2 / 4 / ... / N ranks copying at a time.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29511 tools/d2h_peak.py [--gb 2] [--reps 4]

For every subset size k the first k ranks copy `gb` GB `reps` times (plain cudaMemcpyAsync, one
per repetition) while the others wait; the aggregate GB/s of the slowest rank's wall time is the
ceiling bench.py's `e2e` is judged against (`e2e.ceiling`).  Variants: page-locked memory as
libb200ode allocates it (cuMemHostAlloc, portable), write-combined page-locked memory
(cudaHostAllocWriteCombined: no snooping on the way in, but slow for the host to read), each with
the process bound to its GPU's NUMA node before allocating, or not.  Prints one JSON line."""
import argparse
import ctypes as ct
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=2.0)
    ap.add_argument("--reps", type=int, default=4)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(a.gb * (1 << 30)) // 8
    src = torch.empty(n, dtype=torch.float64, device="cuda")
    src.fill_(1.0)
    cudart = ct.CDLL("/usr/local/cuda/lib64/libcudart.so.12")  # (same primary context as torch's)
    cudart.cudaSetDevice(local)
    out = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxtime(t):
        if world == 1:
            return t
        x = torch.tensor([t], dtype=torch.float64, device="cuda")
        dist.all_reduce(x, op=dist.ReduceOp.MAX)
        return float(x.item())

    from numbalsoda_b200._dist import Ranks
    for numa in (False, True):
        if numa:
            r = Ranks.__new__(Ranks)
            r.local = local
            cpus = Ranks.bind_to_gpu_numa_node(r)
            out["numa_cpus_rank0"] = cpus[:4] + ["..."] + cpus[-2:] if (rank == 0 and cpus) else out.get("numa_cpus_rank0")
        for kind in ("pinned", "write_combined"):
            p = ct.c_void_p()
            # cudaHostAllocPortable = 1, | cudaHostAllocWriteCombined = 4
            rc = cudart.cudaHostAlloc(ct.byref(p), ct.c_size_t(n * 8), ct.c_uint(1 if kind == "pinned" else 5))
            if rc != 0:
                out["%s_alloc_error" % kind] = rc
                continue
            hptr = p.value
            ct.memset(hptr, 0, min(n * 8, 1 << 20))
            k = 1
            sizes = []
            while k < world:
                sizes.append(k)
                k *= 2
            sizes.append(world)
            for k in sizes:
                barrier()
                t0 = time.perf_counter()
                if rank < k:
                    for _ in range(a.reps):
                        cudart.cudaMemcpyAsync(ct.c_void_p(hptr), ct.c_void_p(src.data_ptr()), ct.c_size_t(n * 8),
                                               ct.c_int(2), ct.c_void_p(0))
                    cudart.cudaDeviceSynchronize()
                t = maxtime(time.perf_counter() - t0)
                out["%s%s_%dgpu_gbs" % (kind, "_numa" if numa else "", k)] = round(k * a.reps * n * 8 / t / 1e9, 1)
            cudart.cudaFreeHost(p)
    if rank == 0:
        out.update({"what": "aggregate device-to-host GB/s, k ranks copying concurrently", "gb_per_copy": a.gb,
                    "reps": a.reps, "world": world})
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
