"""Tree-batch data parallelism: contiguous shards of trees per rank and the one collective on the path,
the gather of per-tree result rows.  Plumbing only (torch.distributed; NCCL on GPUs,
gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_trees, rank, world):
    """Contiguous block of ceil(T/G) trees per rank (last ranks may be short or empty)."""
    per = (n_trees + world - 1) // world
    lo = min(n_trees, rank * per)
    hi = min(n_trees, lo + per)
    return lo, hi


def gather_rows(local_rows, n_trees_total=None, group=None):
    """All-gather [T_local, K] result rows into [T_total, K] on every rank.  Equal shards use one
    all_gather_into_tensor; ragged shards are padded to the largest shard and trimmed."""
    world = dist.get_world_size(group)
    if world == 1:
        return local_rows
    t_local = torch.tensor([local_rows.shape[0]], dtype=torch.int64, device=local_rows.device)
    counts = [torch.zeros_like(t_local) for _ in range(world)]
    dist.all_gather(counts, t_local, group=group)
    counts = [int(c.item()) for c in counts]
    k = local_rows.shape[1]
    tmax = max(counts)
    if all(c == tmax for c in counts) and dist.get_backend(group) == "nccl":
        out = torch.empty((world * tmax, k), dtype=local_rows.dtype, device=local_rows.device)
        dist.all_gather_into_tensor(out, local_rows.contiguous(), group=group)
        return out
    pad = torch.zeros((tmax, k), dtype=local_rows.dtype, device=local_rows.device)
    pad[: local_rows.shape[0]] = local_rows
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
    if n_trees_total is not None:
        assert out.shape[0] == n_trees_total
    return out


class PeerGather:
    """Gather of per-tree result rows to rank 0 by the copy engines over NVLink / NVSwitch (no SM is taken from the kernels
    the gather overlaps with): rank 0 owns `nbuf` receive buffers of world x rows x row_bytes, exported as CUDA IPC handles
    (gsm_ipc_export) that every rank opens (gsm_ipc_open); push() is one device-to-device cudaMemcpyAsync into the peer's
    memory on a side stream (gsm_copy_async).  The process group is used for the handle exchange only."""

    def __init__(self, rank, world, device_index, rows, row_bytes, nbuf=3, group=None):
        import ctypes as C
        from . import _lib
        self.C, self.lib = C, _lib.load()
        self.rank, self.world, self.dev = rank, world, device_index
        self.rows, self.row_bytes, self.nbuf = rows, row_bytes, nbuf
        self.bytes_per_rank = rows * row_bytes
        self.local = []      # rank 0: its own allocations
        self.base = []       # every rank: device pointers of rank 0's buffers as seen from this process
        handles = [None] * nbuf
        if rank == 0:
            for _ in range(nbuf):
                p = C.c_void_p()
                rc = self.lib.gsm_device_alloc(C.byref(p), device_index, world * self.bytes_per_rank)
                if rc != 0:
                    raise RuntimeError("gsm_device_alloc failed: %d" % rc)
                self.local.append(p)
            for k, p in enumerate(self.local):
                h = C.create_string_buffer(64)
                rc = self.lib.gsm_ipc_export(p, h)
                if rc != 0:
                    raise RuntimeError("gsm_ipc_export failed: %d" % rc)
                handles[k] = bytes(h.raw)
        if world > 1:
            dist.broadcast_object_list(handles, src=0, group=group)
        for k in range(nbuf):
            if rank == 0:
                self.base.append(self.local[k].value)
            else:
                p = C.c_void_p()
                rc = self.lib.gsm_ipc_open(device_index, handles[k], C.byref(p))
                if rc != 0:
                    raise RuntimeError("gsm_ipc_open failed: %d" % rc)
                self.base.append(p.value)

    def push(self, k, src_ptr, cuda_stream):
        """rows of this rank -> buffer k of rank 0, asynchronously on `cuda_stream`"""
        C = self.C
        rc = self.lib.gsm_copy_async(C.c_void_p(self.base[k] + self.rank * self.bytes_per_rank), C.c_void_p(src_ptr),
                                     self.bytes_per_rank, C.c_void_p(cuda_stream))
        if rc != 0:
            raise RuntimeError("gsm_copy_async failed: %d" % rc)

    def read_root(self, k):
        """rank 0, after the pushes have completed on every rank (barrier): buffer k as a host numpy array [world*rows, row_bytes/8]"""
        import numpy as np
        C = self.C
        out = np.empty((self.world * self.rows, self.row_bytes // 8), np.float64)
        rc = self.lib.gsm_copy_async(out.ctypes.data_as(C.c_void_p), C.c_void_p(self.base[k]), out.nbytes, None)
        if rc != 0:
            raise RuntimeError("gsm_copy_async (to host) failed: %d" % rc)
        torch.cuda.synchronize()
        return out

    def close(self):
        if self.rank == 0:
            for p in self.local:
                self.lib.gsm_device_free(p, self.dev)
        else:
            for p in self.base:
                self.lib.gsm_ipc_close(self.C.c_void_p(p))
        self.local, self.base = [], []
