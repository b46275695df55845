"""Multi-GPU plumbing: one process per GPU, the ensemble sharded by trajectory.

Trajectories are independent, so there is no data-path
collective: torch.distributed is used only to rendezvous, to barrier around timed
regions, to reduce timings (max over ranks) and counts (sum), and -- optionally -- to
gather per-rank results on rank 0.  NCCL on GPUs, gloo in the CPU tests."""
import os

import numpy as np

from ._solve import shard_bounds


class Ranks:
    """Rank context from the torchrun environment (RANK / WORLD_SIZE / LOCAL_RANK)."""

    def __init__(self, backend=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.backend = backend
        self.device = None
        self._dist = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            self.backend = backend
            if backend == "nccl":
                self.device = torch.device("cuda", self.local)
                torch.cuda.set_device(self.device)
                dist.init_process_group("nccl", device_id=self.device)
            else:
                self.device = torch.device("cpu")
                dist.init_process_group(backend)
            self._dist = dist

    def bind_to_gpu_numa_node(self):
        """Pin this process to the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI
        device), so that the page-locked result buffers it allocates afterwards are on that
        NUMA node: with 8 ranks streaming rows to the host, buffers on the far socket halve
        the device-to-host rate.  Returns the CPU list, or None when it cannot be determined."""
        try:
            import subprocess
            bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader",
                                  "-i", str(self.local)], capture_output=True, text=True,
                                 timeout=20).stdout.strip().lower()
            if not bus:
                return None
            bus = bus[-12:] if len(bus) > 12 else bus  # 00000000:1b:00.0 -> 0000:1b:00.0
            path = "/sys/bus/pci/devices/%s/local_cpulist" % bus
            cpus = set()
            for part in open(path).read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            allowed = os.sched_getaffinity(0)
            cpus &= allowed
            if not cpus:
                return None
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
        except Exception:
            return None

    def shard(self, B):
        """[b0, b1) of a global ensemble of B trajectories owned by this rank."""
        b = shard_bounds(B, self.world)
        return b[self.rank], b[self.rank + 1]

    def barrier(self):
        if self._dist is not None:
            self._dist.barrier()

    def _reduce(self, x, op):
        if self._dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device)
        self._dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        """Max over ranks: every multi-GPU time is reported as the slowest rank's."""
        return self._reduce(x, self._dist.ReduceOp.MAX if self._dist else None)

    def sum(self, x):
        return self._reduce(x, self._dist.ReduceOp.SUM if self._dist else None)

    def gather_rows(self, local, B):
        """Concatenate per-rank blocks (sharded by `shard`) on rank 0; other ranks get None."""
        if self._dist is None:
            return local
        import torch
        bounds = shard_bounds(B, self.world)
        counts = [bounds[r + 1] - bounds[r] for r in range(self.world)]
        cmax = max(counts)
        # (gather wants equal shapes: shards differ by at most one row, pad to the largest)
        t = torch.zeros((cmax,) + tuple(local.shape[1:]), dtype=torch.from_numpy(local).dtype)
        t[:local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local))
        if self.backend == "nccl":
            t = t.to(self.device)
        outs = [torch.empty_like(t) for _ in range(self.world)] if self.rank == 0 else None
        self._dist.gather(t, outs, dst=0)
        if self.rank != 0:
            return None
        return torch.cat([o[:c] for o, c in zip(outs, counts)]).cpu().numpy()

    def close(self):
        if self._dist is not None:
            self._dist.destroy_process_group()
            self._dist = None
