"""One-process-per-GPU plumbing for bench.py: rank discovery, barrier, max-over-ranks.

The exact-GP evaluation at the headline size (N=16384: a 2 GiB factor) fits one B200 many
times over, so the metric scales by running INDEPENDENT REPLICAS - different hyper-parameter
vectors per GPU, exactly what the random-restart loop of the reference's optimizers evaluates
(/root/reference/pyGPs/Core/opt.py:305-318).  There is no data-path collective; torch.distributed
(NCCL on GPUs, gloo in the CPU tests) is used only for the barrier and the max-over-ranks timing.
"""
import math
import os


class DistCtx(object):
    def __init__(self, backend=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self._dist = None
        self._torch = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            kw = {}
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                kw["device_id"] = torch.device("cuda", self.local_rank)
            dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world, **kw)
            self._dist, self._torch, self.backend = dist, torch, backend

    def _tensor(self, v):
        t = self._torch.tensor([float(v)], dtype=self._torch.float64)
        return t.cuda(self.local_rank) if self.backend == "nccl" else t

    def barrier(self):
        if self._dist is not None:
            self._dist.barrier()

    def max(self, v):
        if self._dist is None:
            return float(v)
        t = self._tensor(v)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, v):
        if self._dist is None:
            return float(v)
        t = self._tensor(v)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM)
        return float(t.item())

    def broadcast_bytes(self, payload, nbytes, src=0):
        """Broadcast a short byte string (the NCCL unique id) from `src` to all ranks."""
        if self._dist is None:
            return payload
        t = self._torch.zeros(nbytes, dtype=self._torch.uint8)
        if self.rank == src:
            t = self._torch.frombuffer(bytearray(payload), dtype=self._torch.uint8).clone()
        if self.backend == "nccl":
            t = t.cuda(self.local_rank)
        self._dist.broadcast(t, src=src)
        return bytes(t.cpu().numpy().tobytes())

    def shard_engine(self, engine):
        """Bind a libgpk engine to this process group: block-cyclic column sharding over NCCL/NVLink."""
        uid = None
        if self.world > 1:
            uid = self.broadcast_bytes(engine.dist_unique_id() if self.rank == 0 else None, 128)
        engine.dist_init(self.rank, self.world, uid)
        return engine

    def close(self):
        if self._dist is not None:
            self._dist.destroy_process_group()
            self._dist = None


def replica_hyp(step, rank, base_ell=math.log(2.0), base_sf=0.0, base_sn=math.log(0.1)):
    """Hyper-parameters of evaluation `step` on replica `rank`: every (step, rank) differs, so nothing
    but X and y can be cached between evaluations (SURVEY 8(d) definition of one eval)."""
    t = 0.37 * step + 1.3 * rank
    return ([base_ell + 0.02 * math.sin(t), base_sf + 0.02 * math.cos(1.7 * t)], base_sn + 0.01 * math.sin(0.9 * t))


def aggregate_rate(units_per_rank, world, max_seconds):
    """Whole-job throughput: units all ranks processed / max-over-ranks time."""
    return units_per_rank * world / max_seconds
