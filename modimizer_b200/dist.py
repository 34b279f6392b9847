"""Hash-sharded modset over the GPUs of one node: one process per GPU,
torch.distributed (NCCL over NVLink/NVSwitch) for the plumbing.

The sharded build itself is C (modimizer_b200/csrc/sharded.cu, modgpuSharded* in include/modgpu.h): this class is its
ctypes mirror, binding the three communicator callbacks (equal-split all-to-all on a stream, host all-gather,
barrier) to torch.distributed.  The older exchange flavours (buckets through an NCCL all-to-all, per-owner segments,
list exchange) stay here in Python as fallbacks and A/B references.

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


class _Comm(C.Structure):
    """ModgpuComm (include/modgpu.h)"""
    A2A = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
    AG = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t)
    BAR = C.CFUNCTYPE(C.c_int, C.c_void_p)
    _fields_ = [("ctx", C.c_void_p), ("rank", C.c_int), ("world", C.c_int), ("alltoall", A2A), ("allgather", AG), ("barrier", BAR)]


class _DevBytes:
    """a raw device pointer as a CUDA-array-interface object (torch.as_tensor wraps it without a copy)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def torch_comm(group, dev):
    """the three ModgpuComm callbacks on torch.distributed; returns (struct, keepalive)"""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0

    def a2a(ctx, d_send, d_recv, per_peer, stream):
        try:
            n = int(per_peer) * world
            send = torch.as_tensor(_DevBytes(d_send, n), device=dev)
            recv = torch.as_tensor(_DevBytes(d_recv, n), device=dev)
            dist.all_to_all_single(recv, send, group=group)         # ordered on torch's current stream = the modset's stream
            return 0
        except Exception:                                            # never let an exception cross the C frame
            import traceback; traceback.print_exc()
            return -1

    def ag(ctx, h_in, h_out, nbytes):
        try:
            nbytes = int(nbytes)
            mine = torch.frombuffer(bytearray(C.string_at(h_in, nbytes)), dtype=torch.uint8).to(dev)
            out = torch.empty(world * nbytes, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(out, mine, group=group)
            C.memmove(h_out, out.cpu().numpy().tobytes(), world * nbytes)
            return 0
        except Exception:
            import traceback; traceback.print_exc()
            return -1

    def bar(ctx):
        try:
            dist.barrier(group=group)
            return 0
        except Exception:
            import traceback; traceback.print_exc()
            return -1

    cbs = (_Comm.A2A(a2a), _Comm.AG(ag), _Comm.BAR(bar))
    comm = _Comm(None, rank, world, *cbs)
    return comm, cbs


class ShardedModset:
    """A modset whose table is sharded over the ranks of `group` by k-mer hash.

    Each rank calls add()/add_device() with ITS OWN chunk of the input; the
    k-mers it selects are routed to their owners.  `bits` is the per-GPU table
    size (reference tableBits semantics, capacity 2^(bits-2) entries per GPU)."""

    def __init__(self, bits, k=19, w=31, seed=17, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._lib = _lib.load()
        self.dev = torch.device("cuda", torch.cuda.current_device())
        # the C sharded modset owns this rank's shard; self.local is the reference-style view of it
        self._comm, self._comm_keep = torch_comm(group, self.dev)
        self._sh = self._lib.modgpuShardedCreate(bits, k, w, seed, C.byref(self._comm))
        if not self._sh:
            raise ModgpuError("modgpuShardedCreate: " + _lib.last_error())
        self.local = Modset(bits, k, w, seed, _handle=self._lib.modgpuShardedLocal(self._sh), _owner=self)
        # all kernels and NCCL calls are ordered on torch's current stream
        self.local.set_stream(torch.cuda.current_stream().cuda_stream)
        self.total_selected = 0
        self.w = w
        # fused, sync-free exchange (default for world > 1): per-owner segments straight out of hash_select,
        # equal-split all-to-all, bulk insert of the received segments with device-side counts
        self.fused = True
        # fused flavour: "p2p"  = per-(owner, region) buckets stay in the selecting rank's memory; the owner's region
        #                         build reads them through peer-mapped pointers over NVLink (no payload collective,
        #                         no host sync: the fill-count all-to-all doubles as the cross-GPU barrier)
        #                "peer" = the same buckets moved with an equal-split NCCL all-to-all, then a local build
        #                "segments" = per-owner segments + a scatter pass at the receiver
        self.fused_mode = "p2p"
        # deferred peer build (set_accumulate): the k-mers of up to `accumulate` batches wait in the peer buckets and one
        # exchange + build applies them all - a populated per-rank table is rewritten once per group, not per batch
        self.accumulate = 1
        self._peer_cap = 0
        self._seg_cap = 0
        self._sel_pending = 0
        self._sel_acc = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self._ovf_acc = torch.zeros(1, dtype=torch.int32, device=self.dev)

    def close(self):
        if self._sh:
            self._lib.modgpuShardedDestroy(self._sh)            # COLLECTIVE (barrier before the peer buffers are unmapped)
            self._sh = None
            self.local._p = None

    # ---- peer-memory exchange (fused_mode "p2p"): modgpuSharded* (csrc/sharded.cu) ---------------------------
    def reserve(self, max_bases_per_batch):
        """COLLECTIVE: size and map the peer buckets for batches of up to max_bases_per_batch bases per rank.
        Called implicitly by the first add.  False when a rank could not map its peers (everybody then uses NCCL)."""
        if self._lib.modgpuShardedReserve(self._sh, int(max_bases_per_batch)) != 0:
            self.fused_mode = "peer"
            return False
        self._reserved = True
        return True

    def set_accumulate(self, n_batches):
        """COLLECTIVE: up to n_batches batches share one count exchange and one peer build (peer-memory mode).  What
        is waiting reaches the tables at synchronize() - call it before reading the local sets."""
        self.accumulate = max(1, int(n_batches))
        if self.world > 1:
            check(self._lib.modgpuShardedSetAccumulate(self._sh, self.accumulate), "shardedSetAccumulate")
            self._reserved = False

    def set_robust(self, on=True):
        """COLLECTIVE: overflow segments sized for the worst case (after a ModgpuError about a skipped group)"""
        check(self._lib.modgpuShardedSetRobust(self._sh, 1 if on else 0), "shardedSetRobust")
        self._reserved = False

    def clear(self):
        check(self._lib.modgpuShardedClear(self._sh), "shardedClear")

    def _route_and_insert(self, kptr, n):
        st = torch.cuda.current_stream().cuda_stream
        self.total_selected += n
        if self.world > 1:
            self._sel_pending += n
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

    # ---- fused exchange ---------------------------------------------------
    def _ensure_segments(self, nbases):
        expected = nbases // max(self.w, 1) + 1
        cap = int(expected / self.world * 1.15) + 8192
        if cap > self._seg_cap:
            self._seg_cap = cap
            self._send = torch.empty(self.world * cap, dtype=torch.int64, device=self.dev)
            self._recv = torch.empty(self.world * cap, dtype=torch.int64, device=self.dev)
            self._scount = torch.zeros(self.world, dtype=torch.int32, device=self.dev)
            self._rcount = torch.zeros(self.world, dtype=torch.int32, device=self.dev)
        return expected

    def _exchange_and_insert(self, expected):
        cap = self._seg_cap
        dist.all_to_all_single(self._rcount, self._scount, group=self.group)
        dist.all_to_all_single(self._recv, self._send, group=self.group)          # equal splits of cap k-mers
        check(self._lib.modgpuModsetInsertSegments(self.local._p, C.c_void_p(self._recv.data_ptr()), self.world, cap,
                                                   C.c_void_p(self._rcount.data_ptr()), expected), "insertSegments")
        self._sel_acc += self._scount.sum()
        self._ovf_acc = torch.maximum(self._ovf_acc, (self._scount.max() > cap).to(torch.int32).reshape(1))

    def _ensure_peer(self, nbases):
        import math
        R = int(self._lib.modgpuModsetRegions(self.local._p))
        expected = nbases // max(self.w, 1) + 1
        mean = expected / float(self.world * R)
        # bucket capacity: Poisson mean + 10 % + 4 sigma; what does not fit (repeated k-mers pile up in
        # their bucket) travels in the per-owner overflow segments, exchanged with their true sizes
        cap = (int(1.1 * mean + 4.0 * math.sqrt(mean) + 8) + 1) & ~1
        if cap > self._peer_cap:
            self._peer_cap, self._R = cap, R
            self._ovf_cap = max(65536, expected // 4)
            n = self.world * R
            self._sb = torch.empty(n * cap, dtype=torch.int64, device=self.dev)
            self._rb = torch.empty(n * cap, dtype=torch.int64, device=self.dev)
            self._sc = torch.zeros(n, dtype=torch.int32, device=self.dev)
            self._rc = torch.zeros(n, dtype=torch.int32, device=self.dev)
            self._so = torch.empty(self.world * self._ovf_cap, dtype=torch.int64, device=self.dev)
            self._soc = torch.zeros(self.world, dtype=torch.int32, device=self.dev)
            self._roc = torch.zeros(self.world, dtype=torch.int32, device=self.dev)
            self._cnt = torch.zeros(1, dtype=torch.int64, device=self.dev)

    def _peer_exchange_and_build(self):
        """returns False when some rank's overflow segment overflowed: nothing was exchanged or inserted, every
        rank takes the list path for this batch (the decision is collective)"""
        g, G, oc = self.group, self.world, self._ovf_cap
        dist.all_to_all_single(self._roc, self._soc, group=g)                        # overflow counts
        flag = (self._soc.max() > oc).to(torch.int64).reshape(1)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=g)
        host = torch.cat([flag, self._soc.to(torch.int64), self._roc.to(torch.int64), self._cnt]).cpu().tolist()   # the one sync
        if host[0]:
            return False
        ssz, rsz = [int(x) for x in host[1:1 + G]], [int(x) for x in host[1 + G:1 + 2 * G]]
        dist.all_to_all_single(self._rc, self._sc, group=g)                          # bucket fill counts
        dist.all_to_all_single(self._rb, self._sb, group=g)                          # the buckets, equal splits
        check(self._lib.modgpuModsetBuildFromBuckets(self.local._p, C.c_void_p(self._rb.data_ptr()), C.c_void_p(self._rc.data_ptr()),
                                                     self._peer_cap, G, None, 0, None), "buildFromBuckets")
        if sum(ssz) or sum(rsz) or True:                                             # collective: every rank calls it
            so = torch.cat([self._so[o * oc:o * oc + ssz[o]] for o in range(G)]) if sum(ssz) else self._so[:0]
            ro = torch.empty(sum(rsz), dtype=torch.int64, device=self.dev)
            dist.all_to_all_single(ro, so, output_split_sizes=rsz, input_split_sizes=ssz, group=g)
            if ro.numel():
                check(self._lib.modgpuModsetInsertDevice(self.local._p, C.c_void_p(ro.data_ptr()), ro.numel()), "insertOverflow")
            self._keep = (so, ro)
        self._sel_pending += int(host[-1])
        return True

    def synchronize(self):
        """COLLECTIVE: finish the outstanding batches; returns the number of k-mers this rank selected since the last
        call.  Raises when a group of batches was skipped for skew: NOTHING of it was applied on any rank (the C layer
        decides from flags every rank receives) - set_robust() or fused = False, then add those batches again."""
        if self.world == 1:
            n, self.total_selected = self.total_selected, 0
            return n
        nsel = C.c_uint64(0)
        rc = self._lib.modgpuShardedSynchronize(self._sh, C.byref(nsel))
        vals = torch.cat([self._sel_acc, self._ovf_acc.to(torch.int64)]).cpu().tolist()
        self._sel_acc.zero_(); self._ovf_acc.zero_()
        n, self._sel_pending = int(nsel.value) + int(vals[0]) + self._sel_pending, 0
        if rc != 0:
            raise ModgpuError("sharded synchronize (%d): %s" % (rc, _lib.last_error()))
        if vals[1]:
            # the NCCL fallback flavours ("peer" checks before it builds; "segments" does not): state which
            raise ModgpuError("owner segment overflow in fused_mode '%s' (skewed batch): clear() and rebuild with "
                              "ShardedModset.fused = False" % self.fused_mode)
        return n

    def add_device(self, d_bases, d_offsets, nseq, nbases, is_ascii=0):
        """this rank's chunk, resident in device memory (< 2^32 bases per call)"""
        if self.world == 1:                      # nothing to route: the single-GPU pipeline (fused select -> table)
            n = self.local.add_device(d_bases, d_offsets, nseq, nbases, is_ascii)
            self.total_selected += n
            return n
        if self.fused and self.fused_mode == "p2p":
            if getattr(self, "_reserved", False) or self.reserve(nbases):
                check(self._lib.modgpuShardedAddDevice(self._sh, C.c_void_p(d_bases), C.c_void_p(d_offsets), nseq, nbases, is_ascii),
                      "shardedAddDevice")
                return 0                         # the count comes from synchronize()
        if self.fused and self.fused_mode == "peer":
            self._ensure_peer(nbases)
            check(self._lib.modgpuModsetSelectBucketsDevice(self.local._p, C.c_void_p(d_bases), C.c_void_p(d_offsets), nseq, nbases,
                                                            is_ascii, self.world, C.c_void_p(self._sb.data_ptr()), self._peer_cap,
                                                            C.c_void_p(self._sc.data_ptr()), C.c_void_p(self._so.data_ptr()),
                                                            self._ovf_cap, C.c_void_p(self._soc.data_ptr()),
                                                            C.c_void_p(self._cnt.data_ptr())), "selectBuckets")
            if self._peer_exchange_and_build():
                return 0                         # the count comes from synchronize()
            return self._add_device_list(d_bases, d_offsets, nseq, nbases, is_ascii)
        if self.fused:
            expected = self._ensure_segments(nbases)
            check(self._lib.modgpuModsetSelectOwnersDevice(self.local._p, C.c_void_p(d_bases), C.c_void_p(d_offsets), nseq, nbases,
                                                           is_ascii, self.world, C.c_void_p(self._send.data_ptr()), self._seg_cap,
                                                           C.c_void_p(self._scount.data_ptr())), "selectOwners")
            self._exchange_and_insert(expected)
            return 0                             # asynchronous: the count comes from synchronize()
        return self._add_device_list(d_bases, d_offsets, nseq, nbases, is_ascii)

    def _add_device_list(self, d_bases, d_offsets, nseq, nbases, is_ascii=0):
        """list-based exchange: robust for any skew (variable all-to-all of the selected list)"""
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

    def add_pointers(self, host_ptr, offsets_ptr, nseq, is_ascii=0, nbases=None):
        if self.world == 1:
            n = self.local.add_pointers(host_ptr, offsets_ptr, nseq, is_ascii)
            self.total_selected += n
            return n
        if self.fused and nbases is None:
            nbases = int((C.c_uint64 * (nseq + 1)).from_address(offsets_ptr)[nseq])
        if self.fused and self.fused_mode == "p2p":
            if getattr(self, "_reserved", False) or self.reserve(nbases):
                check(self._lib.modgpuShardedAdd(self._sh, C.c_void_p(host_ptr), C.c_void_p(offsets_ptr), nseq, is_ascii), "shardedAdd")
                return 0
        if self.fused and self.fused_mode == "peer":
            self._ensure_peer(nbases)
            check(self._lib.modgpuModsetSelectBucketsHost(self.local._p, C.c_void_p(host_ptr), C.c_void_p(offsets_ptr), nseq,
                                                          is_ascii, self.world, C.c_void_p(self._sb.data_ptr()), self._peer_cap,
                                                          C.c_void_p(self._sc.data_ptr()), C.c_void_p(self._so.data_ptr()),
                                                          self._ovf_cap, C.c_void_p(self._soc.data_ptr()),
                                                          C.c_void_p(self._cnt.data_ptr())), "selectBuckets")
            if self._peer_exchange_and_build():
                return 0
            return self._add_pointers_list(host_ptr, offsets_ptr, nseq, is_ascii)
        if self.fused:
            expected = self._ensure_segments(nbases)
            check(self._lib.modgpuModsetSelectOwnersHost(self.local._p, C.c_void_p(host_ptr), C.c_void_p(offsets_ptr), nseq,
                                                         is_ascii, self.world, C.c_void_p(self._send.data_ptr()), self._seg_cap,
                                                         C.c_void_p(self._scount.data_ptr())), "selectOwners")
            self._exchange_and_insert(expected)
            return 0
        return self._add_pointers_list(host_ptr, offsets_ptr, nseq, is_ascii)

    def _add_pointers_list(self, host_ptr, offsets_ptr, nseq, is_ascii=0):
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
