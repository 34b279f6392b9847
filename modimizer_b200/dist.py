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
        self._p2p = None
        self._peer_cap = 0
        self._seg_cap = 0
        self._sel_pending = 0
        self._sel_acc = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self._ovf_acc = torch.zeros(1, dtype=torch.int32, device=self.dev)

    def close(self):
        self._p2p_release()
        self.local.close()

    # ---- peer-memory exchange (fused_mode "p2p") -----------------------------
    def _p2p_release(self):
        st = self._p2p
        self._p2p = None
        if not st:
            return
        torch.cuda.synchronize()
        if dist.is_initialized() and self.world > 1:
            dist.barrier(group=self.group)              # nobody still reads my buffers
        for p in st["opened"]:
            self._lib.modgpuPeerClose(C.c_void_p(p))
        for p in st["mine"]:
            self._lib.modgpuPeerFree(C.c_void_p(p))

    def reserve(self, max_bases_per_batch):
        """COLLECTIVE: size and map the peer buckets for batches of up to max_bases_per_batch bases per rank.
        Called implicitly by the first add; call it again (on every rank) before feeding larger batches -
        a batch larger than reserved still works, its surplus travels in the overflow segments."""
        import math
        lib, G = self._lib, self.world
        R = int(lib.modgpuModsetRegions(self.local._p))
        t = torch.tensor([int(max_bases_per_batch)], dtype=torch.int64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        nb = int(t.item())
        expected = (nb // max(self.w, 1) + 1) * self.accumulate     # a group of batches shares the buckets
        mean = expected / float(G * R)
        cap = (int(1.1 * mean + 4.0 * math.sqrt(mean) + 8) + 1) & ~1
        ovf_cap = max(65536, expected // 4)
        self._p2p_release()
        mine, opened, ok = [], [], 1
        sb, so = [], []
        for b in range(2):                               # double buffered: one barrier per batch suffices
            p1 = lib.modgpuPeerAlloc(G * R * cap * 8)
            p2 = lib.modgpuPeerAlloc(G * ovf_cap * 8)
            if not p1 or not p2:
                ok = 0
            sb.append(p1 or 0); so.append(p2 or 0)
            mine += [x for x in (p1, p2) if x]
        handles = []
        for ptr in sb + so:
            h = (C.c_ubyte * 64)()
            if not ptr or lib.modgpuPeerExport(C.c_void_p(ptr), h) != 0:
                ok = 0
            handles.append(bytes(h))
        allh = [None] * G
        dist.all_gather_object(allh, (ok, handles), group=self.group)
        ok = min(x[0] for x in allh)
        # peer pointers, already offset to THIS owner's part of every source's arrays
        bptr = [(C.c_void_p * G)() for _ in range(2)]
        optr = [(C.c_void_p * G)() for _ in range(2)]
        if ok:
            for s_rank in range(G):
                ptrs = []
                for i, hb in enumerate(allh[s_rank][1]):
                    if s_rank == self.rank:
                        ptrs.append((sb + so)[i])
                    else:
                        q = lib.modgpuPeerOpen(hb)
                        if not q:
                            ok = 0
                            q = 0
                        else:
                            opened.append(q)
                        ptrs.append(q)
                for b in range(2):
                    bptr[b][s_rank] = ptrs[b] + self.rank * R * cap * 8
                    optr[b][s_rank] = ptrs[2 + b] + self.rank * ovf_cap * 8
        t = torch.tensor([ok], dtype=torch.int64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        self._p2p = {"mine": mine, "opened": opened}
        if not int(t.item()):                            # some rank could not map a peer: everybody uses NCCL
            self._p2p_release()
            self.fused_mode = "peer"
            return False
        self._p2p.update(cap=cap, ovf_cap=ovf_cap, R=R, sb=sb, so=so, bptr=bptr, optr=optr, batch=0, pending=0,
                         sc=torch.zeros(G * R, dtype=torch.int32, device=self.dev),
                         rc=torch.zeros(G * R, dtype=torch.int32, device=self.dev),
                         soc=torch.zeros(G, dtype=torch.int32, device=self.dev),
                         roc=torch.zeros(G, dtype=torch.int32, device=self.dev),
                         cnt=torch.zeros(1, dtype=torch.int64, device=self.dev))
        return True

    def _p2p_add(self, select, nbases):
        """select(sb, cap, sc, so, ovf_cap, soc, cnt) launches this rank's hash/select into bucket set `sb`"""
        if self._p2p is None and not self.reserve(nbases):
            return False
        st = self._p2p
        b = st["batch"] & 1                      # the bucket set of this group of batches
        if st["pending"]:                        # joins the batches already waiting: fill counts are kept (MODGPU_SEL_APPEND)
            flags = getattr(self.local, "_flags", 0)
            check(self._lib.modgpuModsetSetFlags(self.local._p, flags | 128), "set_flags")
            try:
                select(st["sb"][b], st["cap"], st["sc"], st["so"][b], st["ovf_cap"], st["soc"], st["cnt"])
            finally:
                check(self._lib.modgpuModsetSetFlags(self.local._p, flags), "set_flags")
        else:
            select(st["sb"][b], st["cap"], st["sc"], st["so"][b], st["ovf_cap"], st["soc"], st["cnt"])
        st["pending"] += 1
        if st["pending"] >= self.accumulate:
            self._p2p_flush()
        return True

    def _p2p_flush(self):
        """COLLECTIVE: exchange the fill counts of the waiting batches and build them into the owners' tables"""
        st, g, G = self._p2p, self.group, self.world
        if not st or not st["pending"]:
            return
        b = st["batch"] & 1
        st["batch"] += 1
        st["pending"] = 0
        # fill counts to the owners; completing these two small collectives also means every rank's select is done
        dist.all_to_all_single(st["rc"], st["sc"], group=g)
        dist.all_to_all_single(st["roc"], st["soc"], group=g)
        check(self._lib.modgpuModsetBuildFromPeers(self.local._p, st["bptr"][b], C.c_void_p(st["rc"].data_ptr()), st["cap"], G,
                                                   st["optr"][b], st["ovf_cap"], C.c_void_p(st["roc"].data_ptr())), "buildFromPeers")
        self._sel_acc += st["cnt"]
        self._ovf_acc = torch.maximum(self._ovf_acc, (st["soc"].max() > st["ovf_cap"]).to(torch.int32).reshape(1))

    def set_accumulate(self, n_batches):
        """COLLECTIVE: up to n_batches batches share one count exchange and one peer build (peer-memory mode).  What
        is waiting reaches the tables at synchronize() - call it before reading the local sets."""
        if self.world > 1:
            self._p2p_flush()
            self._p2p_release()                  # the buckets are sized for a group: mapped again by the next add
        self.accumulate = max(1, int(n_batches))

    def clear(self):
        if self._p2p and self._p2p["pending"]:   # batches waiting in the peer buckets are dropped with the rest; the next
            self._p2p["pending"] = 0             # group reuses the same bucket set (no peer has been told to read it)
        check(self._lib.modgpuModsetClear(self.local._p), "modsetClear")

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
        """finish the outstanding fused batches; returns the number of k-mers this rank selected since the
        last call.  Raises when a segment overflowed (heavily skewed batch): repeat it with fused = False."""
        if self.world == 1:
            n, self.total_selected = self.total_selected, 0
            return n
        self._p2p_flush()
        vals = torch.cat([self._sel_acc, self._ovf_acc.to(torch.int64)]).cpu().tolist()
        self._sel_acc.zero_(); self._ovf_acc.zero_()
        if vals[1]:
            raise ModgpuError("owner segment overflow (skewed batch): repeat with ShardedModset.fused = False")
        n, self._sel_pending = int(vals[0]) + self._sel_pending, 0
        return n

    def add_device(self, d_bases, d_offsets, nseq, nbases, is_ascii=0):
        """this rank's chunk, resident in device memory (< 2^32 bases per call)"""
        if self.world == 1:                      # nothing to route: the single-GPU pipeline (fused select -> table)
            n = self.local.add_device(d_bases, d_offsets, nseq, nbases, is_ascii)
            self.total_selected += n
            return n
        if self.fused and self.fused_mode == "p2p":
            def sel(sb, cap, sc, so, oc, soc, cnt):
                check(self._lib.modgpuModsetSelectBucketsDevice(self.local._p, C.c_void_p(d_bases), C.c_void_p(d_offsets), nseq, nbases,
                                                                is_ascii, self.world, C.c_void_p(sb), cap, C.c_void_p(sc.data_ptr()),
                                                                C.c_void_p(so), oc, C.c_void_p(soc.data_ptr()),
                                                                C.c_void_p(cnt.data_ptr())), "selectBuckets")
            if self._p2p_add(sel, nbases):
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
            def sel(sb, cap, sc, so, oc, soc, cnt):
                check(self._lib.modgpuModsetSelectBucketsHost(self.local._p, C.c_void_p(host_ptr), C.c_void_p(offsets_ptr), nseq,
                                                              is_ascii, self.world, C.c_void_p(sb), cap, C.c_void_p(sc.data_ptr()),
                                                              C.c_void_p(so), oc, C.c_void_p(soc.data_ptr()),
                                                              C.c_void_p(cnt.data_ptr())), "selectBuckets")
            if self._p2p_add(sel, nbases):
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
