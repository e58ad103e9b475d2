"""Data parallelism: one process per GPU, gradient all-reduce over NCCL (NVLink 5 / NVSwitch).

Replaces `tf.distribute.MirroredStrategy().scope()` of the reference snippets (README.md:39-49,
tutorial/2_multi_scale_NIF.ipynb:493): the global batch is split evenly by rows, every rank runs the
fused forward / reverse kernels on its rows with the loss already divided by the GLOBAL batch, the
flat fp32 gradient buffer is summed across ranks (one all-reduce per step), and every rank applies
the identical Adam update to its replica.  Rows are independent, so there is no other exchange.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


class DataParallel:
    def __init__(self, backend: str | None = None):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if self.world > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            if self.backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                dist.init_process_group("nccl", rank=self.rank, world_size=self.world,
                                        device_id=torch.device("cuda", self.local_rank))
            else:
                dist.init_process_group(self.backend, rank=self.rank, world_size=self.world)

    @property
    def device(self) -> torch.device:
        return torch.device("cuda", self.local_rank) if torch.cuda.is_available() else torch.device("cpu")

    def allreduce_(self, t: torch.Tensor) -> torch.Tensor:
        """In-place sum over ranks (no-op for a single process)."""
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def allreduce_start(self, t: torch.Tensor):
        """Begin an in-place sum over ranks on NCCL's own stream (ordered after the work already enqueued on the
        current stream); returns a handle for allreduce_finish, or None for a single process.  Kernels enqueued on
        the current stream afterwards run concurrently with the transfer."""
        if self.world > 1:
            return dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True)
        return None

    @staticmethod
    def allreduce_finish(handle) -> None:
        """Make the current stream wait for an allreduce_start."""
        if handle is not None:
            handle.wait()

    def max_(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t

    def broadcast_(self, t: torch.Tensor, src: int = 0) -> torch.Tensor:
        if self.world > 1:
            dist.broadcast(t, src)
        return t

    def barrier(self):
        if self.world > 1:
            dist.barrier()

    def attach(self, model, symmetric=None):
        """Make `model` (a nif_b200 Model) data parallel: replicas start from rank 0's parameters.

        symmetric (default: env NIF_B200_SYMM, on): move the flat parameter and gradient buffers into NVLink symmetric
        memory with an NVSwitch multicast mapping, so that the Adam update runs as ONE kernel fused with its collective
        (nif_adam_step_multimem: reduce-scatter by multimem.ld_reduce, update of this rank's slice, all-gather by
        multimem.st) instead of NCCL all-reduce + a full update on every rank.  Falls back to the NCCL path when the
        fabric has no multicast support."""
        self.broadcast_(model.net.theta)
        model.dist = self
        model._symm = None
        if symmetric is None:
            symmetric = os.environ.get("NIF_B200_SYMM", "1") != "0"
        if symmetric and self.world > 1 and self.backend == "nccl":
            model._symm = self._make_symmetric(model.net)
        return model

    def _make_symmetric(self, net):
        try:
            import torch.distributed._symmetric_memory as symm
            dev = net.theta.device
            th = symm.empty(net.n_flat, dtype=torch.float32, device=dev)
            gr = symm.empty(net.n_flat, dtype=torch.float32, device=dev)
            hp = symm.rendezvous(th, dist.group.WORLD)
            hg = symm.rendezvous(gr, dist.group.WORLD)
            ok = torch.tensor([1 if (net.n_flat % 4 == 0 and hp.multicast_ptr and hg.multicast_ptr) else 0], device=dev)
        except Exception as e:  # no symmetric-memory support in this build / on this fabric
            ok = torch.tensor([0], device=net.theta.device)
            self._symm_error = repr(e)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # every rank takes the same path
        if int(ok) == 0:
            return None
        th.copy_(net.theta)
        gr.zero_()
        net._rebind(th, gr)
        return {"hp": hp, "hg": hg, "p_mc": int(hp.multicast_ptr), "g_mc": int(hg.multicast_ptr)}

    def shutdown(self):
        if self.world > 1 and dist.is_initialized():
            dist.destroy_process_group()
