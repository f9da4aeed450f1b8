"""Multi-GPU plumbing: one process per GPU, envs / replay shards partitioned, learner replicated.

SURVEY 8(e): env e of a job with ``total_envs`` lives on rank ``e * world // total_envs`` (contiguous blocks);
its frame ring and replay shard live on the same GPU; each rank samples ``batch // world`` transitions from its
own shard; Q-network, target network and Adam state are replicated; the only collective is one sum
all-reduce of the flat fp32 gradient vector per update (NCCL over NVLink on GPUs, gloo in the CPU tests).
Mean losses divide by the GLOBAL batch inside the loss kernel, so the reduced vector is the gradient of the
global loss and every rank applies the identical Adam step.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init(backend: str | None = None):
    """Initialise torch.distributed from the torchrun environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 or dist.is_initialized():
        return int(os.environ.get("RANK", "0")), world, int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group(backend)
    return dist.get_rank(), dist.get_world_size(), local_rank


def shard_envs(total_envs: int, rank: int, world: int):
    """(first_env_id, num_envs) of this rank: contiguous blocks, remainder spread over the first ranks."""
    base, rem = divmod(total_envs, world)
    n = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n


def owner_of_env(env: int, total_envs: int, world: int) -> int:
    base, rem = divmod(total_envs, world)
    cut = rem * (base + 1)
    return env // (base + 1) if env < cut else rem + (env - cut) // base


def local_batch(global_batch: int, world: int) -> int:
    if global_batch % world:
        raise ValueError("the minibatch must divide evenly over the ranks")
    return global_batch // world


class _DevicePtr:
    """zero-copy torch view of a raw device pointer (an exchange buffer owned by libflappy_b200.so)"""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (int(ptr), False), "version": 3, "strides": None}


class PeerGradExchange:
    """Gradient exchange fused with Adam over NVLink peer memory (csrc/fb_dist.cu): every rank's gradient vector lives in
    an exchange buffer the peers map through CUDA IPC; one kernel per step publishes / waits, sums the ranks' gradients
    in rank order from peer memory and applies TF-1 Adam -- no separate all-reduce.  Needs one process per GPU of one
    node and an initialised torch.distributed group (used once, to swap the 192-byte IPC handles)."""

    def __init__(self, n_floats: int, device, rank: int | None = None, world: int | None = None, connect: bool = True):
        import ctypes as C
        from . import _lib
        self._C, self._lib = C, _lib
        self._L = _lib.lib()
        self.device = torch.device(device)
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.n = int(n_floats)
        h = C.c_void_p()
        _lib.check(self._L.fb_dist_create(self.rank, self.world, self.n, C.byref(h)), "fb_dist_create")
        self._h = h
        self._views = []
        for par in (0, 1):
            p = C.c_void_p()
            _lib.check(self._L.fb_dist_grads(self._h, par, C.byref(p)), "fb_dist_grads")
            self._views.append(torch.as_tensor(_DevicePtr(p.value, self.n), device=self.device))
        if connect and self.world > 1:
            nb = self._L.fb_dist_handle_bytes()
            mine = C.create_string_buffer(nb)
            _lib.check(self._L.fb_dist_handles(self._h, mine), "fb_dist_handles")
            t = torch.frombuffer(bytearray(mine.raw), dtype=torch.uint8).to(self.device)
            out = [torch.empty_like(t) for _ in range(self.world)]
            dist.all_gather(out, t)
            allh = b"".join(bytes(o.cpu().numpy().tobytes()) for o in out)
            _lib.check(self._L.fb_dist_connect(self._h, allh), "fb_dist_connect")
            dist.barrier()

    def set_two_shot(self, on: bool):
        """force the one-shot / two-shot form of the exchange (default: two-shot from four ranks up)"""
        self._lib.check(self._L.fb_dist_set_two_shot(self._h, int(on)), "fb_dist_set_two_shot")

    def connect_local(self, q: int, peer: "PeerGradExchange"):
        """test hook: several ranks inside one process"""
        self._lib.check(self._L.fb_dist_connect_local(self._h, q, peer._h), "fb_dist_connect_local")

    @property
    def grads(self) -> torch.Tensor:
        """the exchange buffer this step's gradients go into"""
        return self._views[self._L.fb_dist_parity(self._h)]

    def adam(self, net, alpha: float, grad_scale: float = 1.0, reduced_out: torch.Tensor | None = None, wait: bool = True):
        self._lib.check(self._L.fb_dist_adam(self._h, net._h, net.params.data_ptr(), net.adam_m.data_ptr(), net.adam_v.data_ptr(),
                                             float(alpha), float(net.beta1), float(net.beta2), float(net.adam_eps), float(grad_scale),
                                             reduced_out.data_ptr() if reduced_out is not None else None, int(wait),
                                             torch.cuda.current_stream(self.device).cuda_stream), "fb_dist_adam")

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.fb_dist_destroy(self._h)
                self._h = None
        except Exception:
            pass


def allreduce_gradients(flat_grads: torch.Tensor):
    """sum all-reduce of the flat gradient vector (898,722 floats = 3.6 MB; dueling 899,235)"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return flat_grads
