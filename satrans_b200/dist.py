"""Multi-GPU plumbing: one process per GPU (torchrun), `torch.distributed` over NCCL/NVLink.

Partitioning (the reference has no counterpart — its only multi-device hook is the
single-process nn.DataParallel of meta_basemodel.py:272-275, which replicates the tables):

  * embedding tables: ROW-SHARDED.  Global row g of the concatenated table lives on rank g % R at local row
    g // R.  Every rank maps every other rank's shard (and its touched-row bitmap / ranks / gradient
    accumulators) through CUDA IPC, so the fused kernels gather rows and scatter-add row gradients DIRECTLY
    over NVLink peer memory — no id all-to-all, no row all-to-all, no host-side split sizes.
  * dense parameters (148,578 fp32 for EN): replicated; their flat gradient buffer is all-reduced (SUM) with
    one NCCL call per step.  The loss is a SUM over samples (reference meta_basemodel.py:317), so per-rank
    gradients add without re-weighting and R ranks x B samples == one step on the R*B global batch.
  * L2 + Adam over untouched rows: local to each shard (no communication).

Cross-rank ordering inside a step: `sat_peer_barrier`, one tiny kernel that exchanges epoch flags through the
peer-mapped workspaces (stream-ordered, no host synchronisation, no NCCL call).
"""
import torch
import torch.distributed as dist


class DistContext:
    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun)")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > 8:
            raise NotImplementedError("row sharding spans the GPUs of ONE node (<= 8)")
        self._flag = None
        self._keep = []            # exported storages must stay alive
        self._opened = {}          # (rank, ipc handle) -> base pointer mapped into this device's address space

    # ---- stream-ordered barrier ---------------------------------------------------------------------
    def barrier(self, device):
        if self._flag is None:
            self._flag = torch.zeros(1, dtype=torch.float32, device=device)
        dist.all_reduce(self._flag, group=self.group)

    def host_barrier(self):
        """Blocking barrier of the processes (set-up time only; steps use Engine._barrier = sat_peer_barrier)."""
        dist.barrier(group=self.group)

    def all_reduce_sum(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    # ---- peer mapping of a CUDA tensor ---------------------------------------------------------------
    def exchange_pointers(self, tensor):
        """All ranks call this with their own tensor; returns the list of R device pointers (own entry = the
        local pointer, the others = CUDA-IPC mappings of the peers' tensors)."""
        from . import _lib
        if tensor.device.type != 'cuda':
            # host-side logic test path (gloo): pointers are only meaningful per process
            objs = [None] * self.world
            dist.all_gather_object(objs, ('cpu', tensor.data_ptr()), group=self.group)
            return [tensor.data_ptr() if r == self.rank else 0 for r in range(self.world)]
        import ctypes
        st = tensor.untyped_storage()
        # (device, cudaIpcMemHandle bytes, storage bytes, offset of the storage inside the exporter's allocation, ...)
        share = st._share_cuda_()
        self._keep.append(st)
        off = int(share[3]) + tensor.storage_offset() * tensor.element_size()
        handle = bytes(share[1])
        # newer torch prefixes [version byte][type byte] (type 'c' = plain cudaMalloc block, 'e' = expandable segment)
        if len(handle) == 66 and handle[1:2] == b'c':
            handle = handle[2:]
        elif len(handle) == 65 and handle[:1] == b'c':
            handle = handle[1:]
        if len(handle) != 64:
            raise RuntimeError('unsupported CUDA IPC handle (%d bytes, type %r): disable expandable_segments'
                               % (len(handle), handle[:2]))
        objs = [None] * self.world
        dist.all_gather_object(objs, (int(share[0]), handle, off), group=self.group)
        ptrs = []
        for r, (peer_dev, handle, o) in enumerate(objs):
            if r == self.rank:
                ptrs.append(tensor.data_ptr())
                continue
            base = self._opened.get((r, handle))
            if base is None:       # one cudaIpcOpenMemHandle per exporter allocation, with THIS device current
                _lib.check(_lib.lib().sat_enable_peer_access(peer_dev), 'sat_enable_peer_access')
                out = ctypes.c_void_p()
                _lib.check(_lib.lib().sat_ipc_open(handle, ctypes.byref(out)), 'sat_ipc_open')
                base = out.value
                self._opened[(r, handle)] = base
            ptrs.append(base + o)
        return ptrs


def local_rows(total_rows, rank, world):
    """Rows of the concatenated table held by `rank` under row % world sharding."""
    return (total_rows - rank + world - 1) // world


def shard_of(row, world):
    return row % world, row // world
