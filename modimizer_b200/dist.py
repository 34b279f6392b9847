"""Hash-sharded modset over the GPUs of one node: one process per GPU,
torch.distributed (NCCL over NVLink/NVSwitch) for the plumbing.

The reference has no distributed mode; its offline recipe is one modset per
input merged with modsetMerge (reference modset.c:106-128, modutils.c:101-103).
Here reads are sharded by input chunk, the table by an independent hash of the
k-mer (mg_owner in csrc/mg_common.cuh), and every batch has exactly one
exchange step: the selected modimizers go to their owner GPU with a variable
all-to-all (a G x G count exchange, then the payload), where the owner inserts
and counts them.  Counting is a commutative sum, so the union of the shards is
bit-identical to the single-GPU modset whatever the number of GPUs.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import ModgpuError, check
from .modset import Modset


def exchange(send, send_counts, group=None):
    """variable all-to-all of a 1-D tensor laid out as contiguous per-destination
    segments.  Returns (recv, recv_counts).  Works on CUDA tensors with NCCL and
    on CPU tensors with gloo (the world_size-2 CPU tests)."""
    world = dist.get_world_size(group)
    sc = torch.as_tensor(send_counts, dtype=torch.int64, device=send.device)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc, group=group)
    hs = [int(x) for x in sc.cpu().tolist()]
    hr = [int(x) for x in rc.cpu().tolist()]
    recv = torch.empty(sum(hr), dtype=send.dtype, device=send.device)
    if world == 1:
        recv.copy_(send[:hs[0]])
    else:
        dist.all_to_all_single(recv, send[:sum(hs)], output_split_sizes=hr, input_split_sizes=hs, group=group)
    return recv, hr


class ShardedModset:
    """A modset whose table is sharded over the ranks of `group` by k-mer hash.

    Each rank calls add()/add_device() with ITS OWN chunk of the input; the
    k-mers it selects are routed to their owners.  `bits` is the per-GPU table
    size (reference tableBits semantics, capacity 2^(bits-2) entries per GPU)."""

    def __init__(self, bits, k=19, w=31, seed=17, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.local = Modset(bits, k, w, seed)
        self._lib = _lib.load()
        self.dev = torch.device("cuda", torch.cuda.current_device())
        # all kernels and NCCL calls are ordered on torch's current stream
        self.local.set_stream(torch.cuda.current_stream().cuda_stream)
        self.total_selected = 0

    def close(self):
        self.local.close()

    def clear(self):
        check(self._lib.modgpuModsetClear(self.local._p), "modsetClear")

    def _route_and_insert(self, kptr, n):
        st = torch.cuda.current_stream().cuda_stream
        self.total_selected += n
        if self.world == 1:
            check(self._lib.modgpuModsetInsertDevice(self.local._p, C.c_void_p(kptr), n), "insert")
            return n
        counts = torch.zeros(self.world, dtype=torch.int64, device=self.dev)
        check(self._lib.modgpuOwnerCount(C.c_void_p(kptr), n, self.world, counts.data_ptr(), st), "ownerCount")
        cursors = torch.cumsum(counts, 0) - counts
        send = torch.empty(max(n, 1), dtype=torch.int64, device=self.dev)
        check(self._lib.modgpuOwnerScatter(C.c_void_p(kptr), n, self.world, cursors.data_ptr(), send.data_ptr(), st), "ownerScatter")
        recv, _ = exchange(send, counts, self.group)
        if recv.numel():
            check(self._lib.modgpuModsetInsertDevice(self.local._p, C.c_void_p(recv.data_ptr()), recv.numel()), "insert")
        self._keep = (send, recv)            # keep alive until the stream has consumed them
        return n

    def add_device(self, d_bases, d_offsets, nseq, nbases, is_ascii=0):
        """this rank's chunk, resident in device memory (< 2^32 bases per call)"""
        if self.world == 1:                      # nothing to route: the single-GPU pipeline (fused select -> table)
            n = self.local.add_device(d_bases, d_offsets, nseq, nbases, is_ascii)
            self.total_selected += n
            return n
        kptr = C.c_void_p()
        n = C.c_uint64()
        check(self._lib.modgpuModsetSelectDevice(self.local._p, C.c_void_p(d_bases), C.c_void_p(d_offsets), nseq, nbases,
                                                 is_ascii, C.byref(kptr), C.byref(n)), "select")
        return self._route_and_insert(kptr.value or 0, n.value)

    def add(self, data, offsets, is_ascii=0):
        """this rank's chunk in host memory (< 2^32 bases per call)"""
        data = np.ascontiguousarray(data, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        return self.add_pointers(data.ctypes.data, offsets.ctypes.data, len(offsets) - 1, is_ascii)

    def add_pointers(self, host_ptr, offsets_ptr, nseq, is_ascii=0):
        if self.world == 1:
            n = self.local.add_pointers(host_ptr, offsets_ptr, nseq, is_ascii)
            self.total_selected += n
            return n
        kptr = C.c_void_p()
        n = C.c_uint64()
        check(self._lib.modgpuModsetSelectHost(self.local._p, C.c_void_p(host_ptr), C.c_void_p(offsets_ptr), nseq,
                                               is_ascii, C.byref(kptr), C.byref(n)), "select")
        return self._route_and_insert(kptr.value or 0, n.value)

    # ---- whole-set results ---------------------------------------------
    def local_max(self):
        return self.local.max

    def global_max(self):
        """total distinct modimizers over all shards"""
        t = torch.tensor([self.local.max], dtype=torch.int64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
        return int(t.item())

    def histogram(self):
        """depth histogram of the whole set: sum of the shard histograms"""
        t = torch.from_numpy(self.local.histogram().astype(np.int64)).to(self.dev)
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
        return t.cpu().numpy().astype(np.uint32)

    def gather_sorted_dump(self):
        """(kmer, depth, info) of the whole set sorted by k-mer, on every rank"""
        v, d, i = self.local.sorted_dump()
        if self.world == 1:
            return v, d, i
        parts = [None] * self.world
        dist.all_gather_object(parts, (v, d, i), group=self.group)
        v = np.concatenate([p[0] for p in parts]); d = np.concatenate([p[1] for p in parts]); i = np.concatenate([p[2] for p in parts])
        o = np.argsort(v, kind="stable")
        return v[o], d[o], i[o]
