// table.cu - K3..K6: the modset as an open-addressing table in HBM.
//
// replaces: Modset / modsetCreate / modsetIndexFind (reference modset.h:17-28,
// modset.c:15-62), the saturating ++depth of addSequence (modutils.c:26),
// depthHistogram (modutils.c:53-63), the -s / -sM classification
// (modutils.c:205-219), modmap's multiplicity classes (modmap.c:125-129) and
// the dense arrays modsetPack leaves behind (modset.c:36-43).
//
// Layout: 2^(bits-1) slots of 16 bytes {kmer u64, count u32, aux u32}; two
// slots per 32-byte sector so find-or-insert + count touches ONE sector.
// Insert = read key, atomicCAS on EMPTY, RED.ADD on the count (+ RED.MIN of the
// input ordinal when the reference's first-occurrence numbering is wanted).
// Probing is linear but confined to an aligned REGION of 2048 slots (32 KiB), so
// that a region is a self-contained sub-table: small batches insert straight
// into HBM (one random sector per probe, bound by random-access throughput,
// ~14 G inserts/s measured, profiles/atomic_roofline_r01.txt); large batches are
// scattered into per-region buckets and every region is built in SHARED MEMORY
// by one block and written back with coalesced 16-byte stores, which turns the
// table traffic into two streaming passes (region_build_kernel).
// 20 algorithmic bytes per selected k-mer.
#include <stdlib.h>
#include "mg_device.cuh"
#include "mg_scan.cuh"
#include "mg_table.cuh"

struct MgBulk { uint32_t slotBits, regionBits, nRegions, cap; uint32_t *cursors; uint64_t *buckets, *overflow; uint64_t overflowCap; uint64_t expected; };

struct ModgpuTable {
  MgSlot *slots;
  int bits;                  // reference tableBits
  uint32_t slotBits;         // bits - 1
  uint64_t nSlots;
  uint64_t maxEntries;       // reference: max must stay < size = 2^(bits-2) - 1   (modset.c:24-26,58)
  unsigned long long *dEntries;   // device: distinct entries
  uint32_t *dError;          // device: probe overflow flag
  uint32_t *dScratch;        // block counts for the ordered compactions
  uint64_t scratchWords;
  uint64_t numbered;         // host: highest dense index handed out
  unsigned long long *hPinned;    // pinned readback word(s)
  bool clearPending;         // the table is logically empty but the slots were not rewritten yet
  uint64_t *dBuckets;        // bulk insert: per-region buckets of k-mers
  uint64_t bucketBytes;
  uint32_t *dCursors;        // bulk insert: per-region fill counts (+ overflow count at [nRegions])
  uint64_t *dOverflow;       // k-mers that did not fit their bucket
  uint64_t overflowCap;
  // deferred build: the buckets stay open over several scattered chunks and the regions are built once
  // (mg_table_bulk_open / _commit / _rollback / _close); every reader of the table closes them first
  bool bulkOpen;
  MgBulk bulk;
  uint64_t bulkRoom, bulkUsed;    // expected k-mers the open buckets were sized for / scattered so far
  uint64_t bulkOvfSeen;           // overflow-list entries in use after the last committed chunk (host copy)
  uint32_t *dCursorSnap;          // the fill counts before the chunk in flight (rollback of a skewed chunk)
  // Modset.info beyond the two copy bits (MS_MINOR, MS_REPEAT, MS_INTERNAL, MS_RDNA: modset.h:49-52, set by modasm):
  // info >> 2 per dense index, kept beside the table once an import carried an info array; aux holds only the copy bits
  uint8_t *dInfoHi;
  uint64_t infoHiCap;
  bool trackInfo;
};

// ------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(256) table_clear_kernel(MgSlot *slots, uint64_t nSlots)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint4 e;
  e.x = 0xFFFFFFFFu; e.y = 0xFFFFFFFFu; e.z = 0u; e.w = MG_AUX_FRESH;
  uint4 *p = reinterpret_cast<uint4 *>(slots);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nSlots; i += stride) p[i] = e;
}

// n is read from device memory (the count hash_select just produced) so that no
// host round trip sits between select and insert; nHost bounds it (the cap).
template <bool EXACT, bool STRICT = false, bool COUNT = true>   // STRICT: a device count beyond the capacity means "incomplete list": do nothing
__global__ void __launch_bounds__(256) table_insert_kernel(MgSlot *slots, uint32_t slotBits,
                                                           const uint64_t *__restrict__ kmers,
                                                           const unsigned long long *__restrict__ nDev, uint64_t nHost,
                                                           uint32_t *__restrict__ slotOut,
                                                           unsigned long long *entries, uint32_t *error)
{
  uint64_t n = nDev ? (uint64_t)*nDev : nHost;
  if (STRICT && n > nHost) return;
  if (n > nHost) n = nHost;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint32_t fresh = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    { uint64_t key = kmers[i] & 0x3FFFFFFFFFFFFFFFull;       // drop the optional strand bit
      bool isNew;
      uint64_t s = probe_insert(slots, slotBits, key, &isNew);
      if (s == 0xFFFFFFFFFFFFFFFFull) { atomicExch(error, 1u); if (slotOut) slotOut[i] = 0xFFFFFFFFu; continue; }
      fresh += isNew ? 1u : 0u;
      if (COUNT) atomicAdd(&slots[s].count, 1u);        // COUNT false: modsetIndexFind (.., true) alone, the caller owns the depth
      if (EXACT) atomicMin(&slots[s].aux, MG_AUX_ORD + (uint32_t)i);
      if (slotOut) slotOut[i] = (uint32_t)s;
    }
  fresh = mg_warp_sum(fresh);
  if (mg_lane() == 0 && fresh) atomicAdd(entries, (unsigned long long)fresh);
}

__global__ void __launch_bounds__(256) table_lookup_kernel(const MgSlot *slots, uint32_t slotBits,
                                                           const uint64_t *__restrict__ kmers,
                                                           const unsigned long long *__restrict__ nDev, uint64_t nHost,
                                                           uint32_t *__restrict__ out)
{
  uint64_t n = nDev ? (uint64_t)*nDev : nHost;
  if (n > nHost) n = nHost;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    { uint64_t key = kmers[i] & 0x3FFFFFFFFFFFFFFFull;
      uint64_t s = probe_find(slots, slotBits, key);
      out[i] = (s == 0xFFFFFFFFFFFFFFFFull) ? 0u : __ldcg(&slots[s].aux);
    }
}

// ---- bulk insert: scatter into per-region buckets, build regions in smem ----
// cursors[region] counts the k-mers aimed at the region; the first `cap` of them
// land in its bucket, the rest in the overflow list (cursors[nRegions] counts it).
__global__ void __launch_bounds__(256) bucket_scatter_kernel(const uint64_t *__restrict__ kmers, uint64_t n, uint32_t slotBits,
                                                             uint32_t nRegions, uint32_t cap, uint32_t *cursors,
                                                             uint64_t *__restrict__ buckets, uint64_t *__restrict__ overflow,
                                                             uint64_t overflowCap, uint32_t *error)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    { const uint64_t key = kmers[i] & 0x3FFFFFFFFFFFFFFFull;
      const uint32_t region = (uint32_t)(mg_slot_hash(key, slotBits) >> MG_REGION_BITS);
      const uint32_t pos = atomicAdd(&cursors[region], 1u);
      if (pos < cap) buckets[(uint64_t)region * cap + pos] = key;
      else
        { const uint32_t o = atomicAdd(&cursors[nRegions], 1u);
          if (o < overflowCap) overflow[o] = key; else atomicExch(error, 1u);
        }
    }
}

// the same scatter reading nSegs segments of segCap k-mers whose fill counts sit in device memory
// (what the multi-GPU exchange delivers: one segment per source rank)
__global__ void __launch_bounds__(256) bucket_scatter_seg_kernel(const uint64_t *__restrict__ segs, uint32_t nSegs, uint64_t segCap,
                                                                 const uint32_t *__restrict__ segCounts, uint32_t slotBits,
                                                                 uint32_t nRegions, uint32_t cap, uint32_t *cursors,
                                                                 uint64_t *__restrict__ buckets, uint64_t *__restrict__ overflow,
                                                                 uint64_t overflowCap, uint32_t *error)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint32_t sg = 0; sg < nSegs; ++sg)
    { uint64_t n = segCounts[sg];
      if (n > segCap) { n = segCap; atomicExch(error, 3u); }          // the sender overflowed its segment
      const uint64_t *kmers = segs + (uint64_t)sg * segCap;
      for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        { const uint64_t key = kmers[i] & 0x3FFFFFFFFFFFFFFFull;
          const uint32_t region = (uint32_t)(mg_slot_hash(key, slotBits) >> MG_REGION_BITS);
          const uint32_t pos = atomicAdd(&cursors[region], 1u);
          if (pos < cap) buckets[(uint64_t)region * cap + pos] = key;
          else
            { const uint32_t o = atomicAdd(&cursors[nRegions], 1u);
              if (o < overflowCap) overflow[o] = key; else atomicExch(error, 1u);
            }
        }
    }
}

// one block builds one region: load (or, for a logically empty table, create)
// its 2048 slots in shared memory, insert + count the bucket with shared-memory
// atomics, store the region back.  Streaming, coalesced, 16 bytes per thread.
#define MG_BUILD_PRELOAD 4
// One block builds one region from nSrc buckets (one per source rank in the multi-GPU exchange):
// source s has its buckets at buckets + s * srcStride * cap and its cursors at cursors + s * srcStride.
template <bool FRESH>
__global__ void __launch_bounds__(256) region_build_multi_kernel(MgSlot *slots, uint32_t slotBits, const uint64_t *__restrict__ buckets,
                                                                 const uint32_t *__restrict__ cursors, uint32_t cap, uint32_t nSrc,
                                                                 uint64_t srcStride, unsigned long long *entries, uint32_t *error)
{
  __shared__ uint4 sR[MG_REGION_SLOTS];
  const uint32_t region = blockIdx.x;
  uint4 *g = reinterpret_cast<uint4 *>(slots) + (uint64_t)region * MG_REGION_SLOTS;
  uint4 e;
  e.x = 0xFFFFFFFFu; e.y = 0xFFFFFFFFu; e.z = 0u; e.w = MG_AUX_FRESH;
#pragma unroll
  for (int i = 0; i < MG_REGION_SLOTS / 256; ++i)
    sR[i * 256 + threadIdx.x] = FRESH ? e : __ldcs(g + i * 256 + threadIdx.x);
  __syncthreads();
  MgSlot *sS = reinterpret_cast<MgSlot *>(sR);
  uint32_t fresh = 0;
  for (uint32_t src = 0; src < nSrc; ++src)
    { uint32_t cnt = cursors[src * srcStride + region];
      if (cnt > cap) cnt = cap;
      const uint64_t *b = buckets + ((uint64_t)src * srcStride + region) * cap;
      for (uint32_t j = threadIdx.x; j < cnt; j += 256)
        { const unsigned long long key = b[j] & 0x3FFFFFFFFFFFFFFFull;
          uint32_t s = (uint32_t)mg_slot_hash(key, slotBits) & (MG_REGION_SLOTS - 1);
          uint32_t probes = 0;
#pragma unroll 1
          for (; probes < MG_REGION_SLOTS; ++probes, s = (s + 1) & (MG_REGION_SLOTS - 1))
            { unsigned long long *kp = reinterpret_cast<unsigned long long *>(&sS[s].key);
              unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(kp);
              if (cur == key) break;
              if (cur == MG_EMPTY)
                { unsigned long long old = atomicCAS(kp, MG_EMPTY, key);
                  if (old == MG_EMPTY) { ++fresh; break; }
                  if (old == key) break;
                }
            }
          if (probes == MG_REGION_SLOTS) { atomicExch(error, 1u); continue; }
          atomicAdd(&sS[s].count, 1u);
        }
    }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < MG_REGION_SLOTS / 256; ++i) __stcs(g + i * 256 + threadIdx.x, sR[i * 256 + threadIdx.x]);
  fresh = mg_warp_sum(fresh);
  if (mg_lane() == 0 && fresh) atomicAdd(entries, (unsigned long long)fresh);
}

// The same with the buckets still in the SOURCE ranks' memory (peer-mapped over NVLink): the transfer of
// the exchange happens here, bucket by bucket, overlapped with the creation of the region in shared
// memory, and only filled entries cross the links.  cursors = local copies of the fill counts
// [nSrc][srcStride]; src.p[s] = source s's bucket array for this owner (nRegions x cap).
struct MgPeerSrc { const uint64_t *p[MODGPU_MAX_PEERS]; };

__device__ __forceinline__ unsigned long long mg_ld_peer(const uint64_t *p)
{ // the data was produced by another GPU's kernel: never served from this SM's L1
  unsigned long long v;
  asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

// Persistent and software-pipelined: a block builds regions b, b + grid, ...; while region i is being
// built in shared memory, the k-mers of region i+1 are already in flight into registers and the fill
// counts of region i+2 are being fetched, so neither the NVLink round trip (peer sources) nor the DRAM
// latency (local source) sits on the critical path.  Work split: a warp owns (source, part) units - with
// nSrc <= 8 sources every bucket is read by 8 / nSrc warps, lanes on consecutive k-mers (coalesced
// 256-byte requests over NVLink), no search for the source of an element.  PEER: loads bypass this SM's L1.
// Geometry: NW warps per block, PRE k-mers per lane and unit kept in registers.  A bucket part that holds more than 32 PRE
// k-mers makes its warp fetch the rest with the latency exposed, and the block barrier makes the whole block wait for it:
// with 8 warps a part holds 740 / 8 = 92 k-mers on average (the genome build at load 0.36), so three preloads (96) miss
// for every third part; 16 warps halve the parts (46 +- 7 against 64) - measured (r02): 0.63 ms against 0.53 ms with 8
// warps and 5 blocks per SM on the local genome build, and four preloads spill: the 8-warp form stays the default.
template <bool FRESH, bool PEER, int MG_PIPE_UNITS, int NW = 8, int MG_PIPE_PRELOAD = 3>   // units per warp: 1 for nSrc <= NW, 2 up to 2 NW
__global__ void __launch_bounds__(NW * 32, NW == 16 ? 3 : 5) region_build_pipe_kernel(MgSlot *slots, uint32_t slotBits, const MgPeerSrc src,
                                                                const uint32_t *__restrict__ cursors, uint32_t cap, uint32_t nSrc,
                                                                uint64_t srcStride, uint32_t nRegions,
                                                                unsigned long long *entries, uint32_t *error,
                                                                const uint32_t *__restrict__ guard = nullptr, uint32_t guardLimit = 0,
                                                                int emptyOnSkip = 0)
{
  __shared__ uint4 sR[MG_REGION_SLOTS];
  // the scatter that filled the buckets ran out of overflow space: the batch is incomplete, touch nothing - or, when the
  // caller cannot undo its bookkeeping (the sharded build: every rank takes the same decision from the same flags), create
  // the regions of a fresh table empty and add nothing
  const bool skip = guard && __ldg(guard) > guardLimit;
  if (skip && !(FRESH && emptyOnSkip)) return;
  __shared__ const uint64_t *sSrc[MODGPU_MAX_PEERS];            // (a dynamically indexed kernel parameter would live in local memory)
  MgSlot *sS = reinterpret_cast<MgSlot *>(sR);
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t fresh = 0;
  if (tid == 0)
    {
#pragma unroll
      for (int i = 0; i < MODGPU_MAX_PEERS; ++i) sSrc[i] = src.p[i];
    }
  __syncthreads();
  const uint32_t parts = nSrc >= NW ? 1u : (uint32_t)NW / nSrc;   // warps per bucket
  const uint32_t nUnits = nSrc * parts;
  // this warp's units: u = warp, warp + 8
  const uint64_t *uBase[MG_PIPE_UNITS];
  const uint32_t *uCur[MG_PIPE_UNITS];
  uint32_t uFirst[MG_PIPE_UNITS];
  bool uOn[MG_PIPE_UNITS];
#pragma unroll
  for (int q = 0; q < MG_PIPE_UNITS; ++q)
    { const uint32_t u = warp + NW * q;
      uOn[q] = u < nUnits;
      const uint32_t sIdx = uOn[q] ? u % nSrc : 0u;
      uBase[q] = sSrc[sIdx];
      uCur[q] = cursors + sIdx * srcStride;
      uFirst[q] = (u / nSrc) * 32 + lane;                        // first element of this lane inside the bucket
    }
  const uint32_t step = parts * 32;

  auto count_of = [&](int q, uint32_t region) -> uint32_t {
    uint32_t c = 0;
    if (uOn[q] && region < nRegions && !skip) { c = __ldg(uCur[q] + region); if (c > cap) c = cap; }
    return c;
  };
  auto load_key = [&](int q, uint32_t region, uint32_t j) -> unsigned long long {
    const uint64_t *p = uBase[q] + (uint64_t)region * cap + j;
    return (PEER ? mg_ld_peer(p) : __ldcs(reinterpret_cast<const unsigned long long *>(p))) & 0x3FFFFFFFFFFFFFFFull;
  };
  auto insert = [&](unsigned long long key) {
    uint32_t sl = (uint32_t)mg_slot_hash(key, slotBits) & (MG_REGION_SLOTS - 1);
    uint32_t probes = 0;
#pragma unroll 1
    for (; probes < MG_REGION_SLOTS; ++probes, sl = (sl + 1) & (MG_REGION_SLOTS - 1))
      { unsigned long long *kp = reinterpret_cast<unsigned long long *>(&sS[sl].key);
        if (FRESH)
          { const unsigned long long old = atomicCAS(kp, MG_EMPTY, key);     // see region_build_kernel
            if (old == MG_EMPTY) { ++fresh; break; }
            if (old == key) break;
            continue;
          }
        unsigned long long cu = *reinterpret_cast<volatile unsigned long long *>(kp);
        if (cu == key) break;
        if (cu == MG_EMPTY)
          { unsigned long long old = atomicCAS(kp, MG_EMPTY, key);
            if (old == MG_EMPTY) { ++fresh; break; }
            if (old == key) break;
          }
      }
    if (probes == MG_REGION_SLOTS) { atomicExch(error, 1u); return; }
    atomicAdd(&sS[sl].count, 1u);
  };

  uint32_t region = blockIdx.x;
  if (region >= nRegions) return;
  // pipeline state: cnt = fills of the current region, pre = its first k-mers; cntN / cntA = one and two regions ahead
  uint32_t cnt[MG_PIPE_UNITS], cntN[MG_PIPE_UNITS], cntA[MG_PIPE_UNITS];
  unsigned long long pre[MG_PIPE_UNITS][MG_PIPE_PRELOAD];
#pragma unroll
  for (int q = 0; q < MG_PIPE_UNITS; ++q)
    { cnt[q] = count_of(q, region);
      cntN[q] = count_of(q, region + gridDim.x);
      cntA[q] = count_of(q, region + 2 * gridDim.x);
#pragma unroll
      for (int v = 0; v < MG_PIPE_PRELOAD; ++v)
        { const uint32_t j = uFirst[q] + v * step;
          pre[q][v] = (j < cnt[q]) ? load_key(q, region, j) : MG_EMPTY;
        }
    }

  for (; region < nRegions; region += gridDim.x)
    { const uint32_t next = region + gridDim.x;
      uint4 *g = reinterpret_cast<uint4 *>(slots) + (uint64_t)region * MG_REGION_SLOTS;
      bool any = FRESH;
      if (!FRESH)
        { uint32_t c = 0;
#pragma unroll
          for (int q = 0; q < MG_PIPE_UNITS; ++q) c |= cnt[q];
          any = __syncthreads_or(c != 0);                     // nothing to add: leave the region alone (block-uniform)
        }
      if (any)
        { uint4 e;
          e.x = 0xFFFFFFFFu; e.y = 0xFFFFFFFFu; e.z = 0u; e.w = MG_AUX_FRESH;
#pragma unroll
          for (int i = 0; i < MG_REGION_SLOTS / (NW * 32); ++i)
            sR[i * (NW * 32) + tid] = FRESH ? e : __ldcs(g + i * (NW * 32) + tid);
        }
      // the next region's k-mers: in flight while this one is built
      unsigned long long nxt[MG_PIPE_UNITS][MG_PIPE_PRELOAD];
#pragma unroll
      for (int q = 0; q < MG_PIPE_UNITS; ++q)
#pragma unroll
        for (int v = 0; v < MG_PIPE_PRELOAD; ++v)
          { const uint32_t j = uFirst[q] + v * step;
            nxt[q][v] = (j < cntN[q]) ? load_key(q, next, j) : MG_EMPTY;
          }
      uint32_t cntB[MG_PIPE_UNITS];
#pragma unroll
      for (int q = 0; q < MG_PIPE_UNITS; ++q) cntB[q] = count_of(q, next + 2 * gridDim.x);     // three regions ahead
      __syncthreads();                                         // the region is initialised
      if (any)
        {
#pragma unroll
          for (int q = 0; q < MG_PIPE_UNITS; ++q)
            {
#pragma unroll
              for (int v = 0; v < MG_PIPE_PRELOAD; ++v)
                if (uFirst[q] + v * step < cnt[q]) insert(pre[q][v]);
              for (uint32_t j = uFirst[q] + MG_PIPE_PRELOAD * step; j < cnt[q]; j += step) insert(load_key(q, region, j));
            }
        }
      __syncthreads();                                         // the region is complete
      if (any)
        {
#pragma unroll
          for (int i = 0; i < MG_REGION_SLOTS / (NW * 32); ++i) __stcs(g + i * (NW * 32) + tid, sR[i * (NW * 32) + tid]);
        }
#pragma unroll
      for (int q = 0; q < MG_PIPE_UNITS; ++q)
        { cnt[q] = cntN[q]; cntN[q] = cntA[q]; cntA[q] = cntB[q];
#pragma unroll
          for (int v = 0; v < MG_PIPE_PRELOAD; ++v) pre[q][v] = nxt[q][v];
        }
    }
  fresh = mg_warp_sum(fresh);
  if (mg_lane() == 0 && fresh) atomicAdd(entries, (unsigned long long)fresh);
}

// The peer build with the transfer on the TMA engine: the filled part of every source bucket of a region comes over
// NVLink as ONE bulk copy (cp.async.bulk from the peer-mapped pointer into a shared-memory ring, two regions ahead,
// completion on an mbarrier) instead of 8-byte loads of the lanes - a few large NVLink reads per region instead of many
// 256-byte ones, no registers tied up by k-mers in flight, and two regions of prefetch distance.  Warp 0 issues the
// copies for region i+2 when region i has been consumed (it reads the fill counts then: local memory, behind the
// write-back of the other threads); a warp inserts one (source, part) unit from the ring.
template <bool FRESH>
__global__ void __launch_bounds__(256, 4) region_build_tma_kernel(MgSlot *slots, uint32_t slotBits, const MgPeerSrc src,
                                                                  const uint32_t *__restrict__ cursors, uint32_t cap, uint32_t nSrc,
                                                                  uint64_t srcStride, uint32_t nRegions, unsigned long long *entries,
                                                                  uint32_t *error, const uint32_t *__restrict__ guard, const uint32_t hint)
{
  __shared__ uint4 sR[MG_REGION_SLOTS];
  __shared__ __align__(8) uint64_t sBar[2];
  __shared__ uint32_t sCnt[2][MODGPU_MAX_PEERS];
  extern __shared__ __align__(128) uint8_t sDynRing[];                 // 2 stages x nSrc x cap k-mers
  const bool skip = guard && __ldg(guard) > 0u;
  if (skip && !FRESH) return;
  MgSlot *sS = reinterpret_cast<MgSlot *>(sR);
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint64_t *ring = reinterpret_cast<uint64_t *>(sDynRing);
  const uint32_t stageWords = nSrc * cap;
  if (blockIdx.x >= nRegions) return;

  auto issue = [&](uint32_t region, uint32_t stage) {                  // warp 0: lane s serves source s
    if (lane < nSrc)
      { uint32_t c = 0;
        if (region < nRegions && !skip) { c = __ldg(cursors + lane * srcStride + region); if (c > cap) c = cap; }
        sCnt[stage][lane] = c;
        const uint32_t bytes = (c * 8u + 15u) & ~15u;                  // (cap is even: a whole bucket is a multiple of 16 bytes)
        if (bytes)
          { // the buckets are read once: evict-first releases their lines (the select pass stored them evict-last)
            if (hint) mg_tma_load_1d_hint(ring + (size_t)stage * stageWords + (size_t)lane * cap, src.p[lane] + (uint64_t)region * cap, bytes, &sBar[stage], MG_L2_EVICT_FIRST);
            else mg_tma_load_1d(ring + (size_t)stage * stageWords + (size_t)lane * cap, src.p[lane] + (uint64_t)region * cap, bytes, &sBar[stage]);
          }
        mg_mbar_expect_tx(&sBar[stage], bytes);                        // one arrival per source, with its bytes
      }
  };

  if (tid == 0)
    { mg_mbar_init(&sBar[0], nSrc); mg_mbar_init(&sBar[1], nSrc);
      mg_fence_barrier_init();
      mg_fence_proxy_async();
    }
  __syncthreads();
  if (warp == 0) { issue(blockIdx.x, 0); issue(blockIdx.x + gridDim.x, 1); }
  const uint32_t parts = nSrc >= 8 ? 1u : 8u / nSrc;                    // warps per bucket
  const uint32_t nUnits = nSrc * parts;
  uint32_t fresh = 0;
  uint32_t it = 0;
  for (uint32_t region = blockIdx.x; region < nRegions; region += gridDim.x, ++it)
    { const uint32_t stage = it & 1;
      uint4 *g = reinterpret_cast<uint4 *>(slots) + (uint64_t)region * MG_REGION_SLOTS;
      mg_mbar_wait(&sBar[stage], (it >> 1) & 1);                       // this region's k-mers have landed (and sCnt is two barriers old)
      bool any = FRESH;
      if (!FRESH)
        { uint32_t c = 0;
          for (uint32_t s = 0; s < nSrc; ++s) c |= sCnt[stage][s];
          any = c != 0;                                                // block-uniform: nothing to add leaves the region alone
        }
      if (any)
        { uint4 e;
          e.x = 0xFFFFFFFFu; e.y = 0xFFFFFFFFu; e.z = 0u; e.w = MG_AUX_FRESH;
#pragma unroll
          for (int i = 0; i < MG_REGION_SLOTS / 256; ++i)
            sR[i * 256 + tid] = FRESH ? e : __ldcs(g + i * 256 + tid);
        }
      __syncthreads();                                                 // the region is initialised
      if (any)
        for (uint32_t u = warp; u < nUnits; u += 8)
          { const uint32_t s = u % nSrc, part = u / nSrc, cnt = sCnt[stage][s];
            const uint64_t *b = ring + (size_t)stage * stageWords + (size_t)s * cap;
            for (uint32_t j = part * 32 + lane; j < cnt; j += parts * 32)
              { const unsigned long long key = b[j] & 0x3FFFFFFFFFFFFFFFull;
                uint32_t sl = (uint32_t)mg_slot_hash(key, slotBits) & (MG_REGION_SLOTS - 1);
                uint32_t probes = 0;
#pragma unroll 1
                for (; probes < MG_REGION_SLOTS; ++probes, sl = (sl + 1) & (MG_REGION_SLOTS - 1))
                  { unsigned long long *kp = reinterpret_cast<unsigned long long *>(&sS[sl].key);
                    if (FRESH)
                      { const unsigned long long old = atomicCAS(kp, MG_EMPTY, key);
                        if (old == MG_EMPTY) { ++fresh; break; }
                        if (old == key) break;
                        continue;
                      }
                    unsigned long long cu = *reinterpret_cast<volatile unsigned long long *>(kp);
                    if (cu == key) break;
                    if (cu == MG_EMPTY)
                      { unsigned long long old = atomicCAS(kp, MG_EMPTY, key);
                        if (old == MG_EMPTY) { ++fresh; break; }
                        if (old == key) break;
                      }
                  }
                if (probes == MG_REGION_SLOTS) { atomicExch(error, 1u); continue; }
                atomicAdd(&sS[sl].count, 1u);
              }
          }
      __syncthreads();                                                 // the region is complete, the ring stage consumed
      if (warp == 0) issue(region + 2 * gridDim.x, stage);             // two regions ahead, behind the write-back below
      if (any)
        {
#pragma unroll
          for (int i = 0; i < MG_REGION_SLOTS / 256; ++i) __stcs(g + i * 256 + tid, sR[i * 256 + tid]);
        }
    }
  fresh = mg_warp_sum(fresh);
  if (mg_lane() == 0 && fresh) atomicAdd(entries, (unsigned long long)fresh);
}

template <bool FRESH, int PRELOAD>
__global__ void __launch_bounds__(256) region_build_kernel(MgSlot *slots, uint32_t slotBits, const uint64_t *__restrict__ buckets,
                                                              const uint32_t *__restrict__ cursors, uint32_t cap,
                                                              unsigned long long *entries, uint32_t *error,
                                                              const uint32_t *__restrict__ guard, uint32_t guardLimit)
{
  __shared__ uint4 sR[MG_REGION_SLOTS];
  // the scatter that filled the buckets ran out of overflow space: the batch is incomplete, touch nothing
  // (the host learns it from the same counter after the launch and repeats the batch through the list path)
  if (guard && __ldg(guard) > guardLimit) return;
  const uint32_t region = blockIdx.x;
  uint4 *g = reinterpret_cast<uint4 *>(slots) + (uint64_t)region * MG_REGION_SLOTS;
  uint32_t cnt = cursors[region];
  if (cnt > cap) cnt = cap;
  if (!FRESH && cnt == 0) return;                              // nothing to add: leave the region alone
  // the bucket's k-mers first: their DRAM latency overlaps the creation / load of the region
  const uint64_t *b = buckets + (uint64_t)region * cap;
  unsigned long long pre[PRELOAD > 0 ? PRELOAD : 1];
#pragma unroll
  for (int u = 0; u < PRELOAD; ++u)
    { const uint32_t j = u * 256 + threadIdx.x;
      pre[u] = (j < cnt) ? __ldcs(reinterpret_cast<const unsigned long long *>(b) + j) : MG_EMPTY;
    }
  uint4 e;
  e.x = 0xFFFFFFFFu; e.y = 0xFFFFFFFFu; e.z = 0u; e.w = MG_AUX_FRESH;
#pragma unroll
  for (int i = 0; i < MG_REGION_SLOTS / 256; ++i)
    sR[i * 256 + threadIdx.x] = FRESH ? e : __ldcs(g + i * 256 + threadIdx.x);
  __syncthreads();
  MgSlot *sS = reinterpret_cast<MgSlot *>(sR);
  uint32_t fresh = 0;
  for (uint32_t j = threadIdx.x, u = 0; j < cnt; j += 256, ++u)
    { unsigned long long key;
      if (u < (uint32_t)PRELOAD)
        { // select from the preloaded registers without dynamic indexing
          key = pre[0];
#pragma unroll
          for (int v = 1; v < PRELOAD; ++v) if (u == (uint32_t)v) key = pre[v];
        }
      else key = b[j];
      uint32_t s = (uint32_t)mg_slot_hash(key, slotBits) & (MG_REGION_SLOTS - 1);
      uint32_t probes = 0;
#pragma unroll 1
      for (; probes < MG_REGION_SLOTS; ++probes, s = (s + 1) & (MG_REGION_SLOTS - 1))
        { unsigned long long *kp = reinterpret_cast<unsigned long long *>(&sS[s].key);
          if (FRESH)
            { // a region that starts empty: the slot is probably free, claim it without reading it first (one
              // shared-memory operation instead of two for every new k-mer; measured 0.64 -> 0.58 ms on the
              // genome build, +3 % on a 30x readset where most k-mers repeat)
              const unsigned long long old = atomicCAS(kp, MG_EMPTY, key);
              if (old == MG_EMPTY) { ++fresh; break; }
              if (old == key) break;
              continue;
            }
          unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(kp);
          if (cur == key) break;
          if (cur == MG_EMPTY)
            { unsigned long long old = atomicCAS(kp, MG_EMPTY, key);
              if (old == MG_EMPTY) { ++fresh; break; }
              if (old == key) break;
            }
        }
      if (probes == MG_REGION_SLOTS) { atomicExch(error, 1u); continue; }
      atomicAdd(&sS[s].count, 1u);
    }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < MG_REGION_SLOTS / 256; ++i) __stcs(g + i * 256 + threadIdx.x, sR[i * 256 + threadIdx.x]);
  fresh = mg_warp_sum(fresh);
  if (mg_lane() == 0 && fresh) atomicAdd(entries, (unsigned long long)fresh);
}

// ---- numbering: rank the not-yet-numbered entries (mg_scan.cuh) ------------
struct FlagNewSlot {         // entries inserted but not numbered yet, in slot order
  MgSlot *slots;
  uint64_t base;             // indices already handed out
  __device__ uint32_t value(uint64_t i) const { return (slots[i].key != MG_EMPTY && slots[i].aux >= MG_AUX_ORD) ? 1u : 0u; }
  __device__ void emit(uint64_t i, uint32_t rank, uint32_t v) const { if (v) slots[i].aux = (uint32_t)((base + rank + 1) << 2); }
};

struct FlagFirstOccurrence { // list elements that were the first occurrence of a new entry
  MgSlot *slots;
  const uint32_t *slotOf;
  uint64_t base;
  __device__ uint32_t value(uint64_t i) const
  { uint32_t s = slotOf[i]; return (s != 0xFFFFFFFFu && slots[s].aux == MG_AUX_ORD + (uint32_t)i) ? 1u : 0u; }
  __device__ void emit(uint64_t i, uint32_t rank, uint32_t v) const
  { if (v) slots[slotOf[i]].aux = (uint32_t)((base + rank + 1) << 2); }
};

__global__ void __launch_bounds__(256) gather_index_kernel(const MgSlot *slots, const uint32_t *__restrict__ slotOf,
                                                           uint64_t n, uint32_t *__restrict__ index)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    { uint32_t s = slotOf[i];
      index[i] = (s == 0xFFFFFFFFu) ? 0u : (__ldcg(&slots[s].aux) >> 2);
    }
}

// depth = count clamped to 65535: the reference's ++ saturates (modutils.c:26)
__device__ __forceinline__ uint32_t clamp16(uint32_t c) { return c > 65535u ? 65535u : c; }

#define MG_HIST_SMEM 4096
__global__ void __launch_bounds__(256) table_hist_kernel(const MgSlot *slots, uint64_t nSlots, uint32_t *bins)
{
  __shared__ uint32_t sBins[MG_HIST_SMEM];
  for (int i = threadIdx.x; i < MG_HIST_SMEM; i += 256) sBins[i] = 0;
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nSlots; i += stride)
    { uint4 v = __ldcs(reinterpret_cast<const uint4 *>(slots) + i);
      if ((v.x & v.y) == 0xFFFFFFFFu) continue;          // EMPTY key
      uint32_t d = clamp16(v.z);
      if (d < MG_HIST_SMEM) atomicAdd(&sBins[d], 1u); else atomicAdd(&bins[d], 1u);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < MG_HIST_SMEM; i += 256)
    { uint32_t c = sBins[i]; if (c) atomicAdd(&bins[i], c); }
}

// mode 0: -s   depth<c1 -> 0, <c2 -> 1, <cM -> 2, else 3      (modutils.c:205-214)
// mode 1: -sM  depth>=cM -> 3, others unchanged                (modutils.c:215-219)
// mode 2: exact multiplicity 1 -> 1, 2 -> 2, else 3            (modmap.c:125-129)
// mode 3: tally the current classes only                      (modsetSummary, modset.c:149-150)
__global__ void __launch_bounds__(256) table_classify_kernel(MgSlot *slots, uint64_t nSlots, int mode,
                                                             int c1, int c2, int cM, int zeroDepth, uint32_t *classCounts)
{
  uint32_t tally[4] = { 0, 0, 0, 0 };
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nSlots; i += stride)
    { uint4 v = reinterpret_cast<const uint4 *>(slots)[i];
      if ((v.x & v.y) == 0xFFFFFFFFu) continue;
      if (v.w >= MG_AUX_ORD) continue;                   // not numbered: not part of the set yet
      int d = zeroDepth ? 0 : (int)clamp16(v.z);            // modmap-built sets keep ms->depth at 0
      uint32_t cls = v.w & 3u;
      if (mode == 0) cls = d < c1 ? 0u : d < c2 ? 1u : d < cM ? 2u : 3u;
      else if (mode == 1) { if (d >= cM) cls = 3u; }
      else if (mode == 2) cls = (v.z == 1u) ? 1u : (v.z == 2u) ? 2u : 3u;
      if (mode != 3) slots[i].aux = (v.w & ~3u) | cls;
      ++tally[cls];
    }
  if (classCounts)
    {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        { uint32_t t = mg_warp_sum(tally[c]);
          if (mg_lane() == 0 && t) atomicAdd(&classCounts[c], t);
        }
    }
}

__global__ void __launch_bounds__(256) table_export_kernel(const MgSlot *slots, uint64_t nSlots,
                                                           uint64_t *value, uint16_t *depth, uint8_t *info, uint32_t *count32,
                                                           const uint8_t *__restrict__ infoHi)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nSlots; i += stride)
    { uint4 v = __ldcs(reinterpret_cast<const uint4 *>(slots) + i);
      if ((v.x & v.y) == 0xFFFFFFFFu || v.w >= MG_AUX_ORD) continue;
      uint32_t ix = (v.w >> 2) - 1;                      // reference indices are 1-based
      if (value) value[ix] = ((uint64_t)v.y << 32) | v.x;
      if (depth) depth[ix] = (uint16_t)clamp16(v.z);
      if (info) info[ix] = (uint8_t)((v.w & 3u) | (infoHi ? ((uint32_t)infoHi[ix] << 2) : 0u));   // the whole byte (modset.c:73,87)
      if (count32) count32[ix] = v.z;
    }
}

__global__ void __launch_bounds__(256) table_import_kernel(MgSlot *slots, uint32_t slotBits,
                                                           const uint64_t *__restrict__ value,
                                                           const uint16_t *__restrict__ depth,
                                                           const uint8_t *__restrict__ info, uint64_t n, uint64_t indexBase,
                                                           unsigned long long *entries, uint32_t *error, uint8_t *__restrict__ infoHi)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint32_t fresh = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    { bool isNew;
      if (infoHi) infoHi[indexBase + i] = info ? (uint8_t)(info[i] >> 2) : (uint8_t)0;
      uint64_t s = probe_insert(slots, slotBits, value[i], &isNew);
      if (s == 0xFFFFFFFFFFFFFFFFull || !isNew) { atomicExch(error, s == 0xFFFFFFFFFFFFFFFFull ? 1u : 2u); continue; }
      ++fresh;
      slots[s].count = depth ? (uint32_t)depth[i] : 0u;
      slots[s].aux = (uint32_t)((indexBase + i + 1) << 2) | (info ? (info[i] & 3u) : 0u);
    }
  fresh = mg_warp_sum(fresh);
  if (mg_lane() == 0 && fresh) atomicAdd(entries, (unsigned long long)fresh);
}

// ------------------------------------------------------------------- host
static unsigned grid_for(uint64_t n, int perSm)
{
  uint64_t blocks = (n + 255) / 256;
  uint64_t maxBlocks = (uint64_t)mg_num_sms() * perSm;
  if (blocks > maxBlocks) blocks = maxBlocks;
  if (!blocks) blocks = 1;
  return (unsigned)blocks;
}

extern "C" ModgpuTable *modgpuTableCreate(int bits, void *stream)
{
  if (bits < 20 || bits > 34)                            // modset.c:17
    { mg_set_error("table bits %d must be between 20 and 34", bits); return nullptr; }
  // slot ids travel as uint32 (slotOf / slotOut arrays of the exact-order, merge and readset paths; 0xFFFFFFFF = none),
  // and the dense index has 30 bits beside the two copy bits: one GPU table stops at bits 32 (2^31 slots = 32 GiB,
  // 805 M entries).  The reference's 33 and 34 are reached by sharding the table over GPUs (modgpuSharded*).
  if (bits > 32)
    { mg_set_error("table bits %d: one GPU table supports 20..32 (2^31 slots, 805 M entries); shard larger sets over GPUs", bits);
      return nullptr;
    }
  // random 16-byte probes: ask L2 to fetch 32-byte sectors instead of whole lines
  // (MODGPU_L2_FETCH=32|64|128 overrides; measured in profiles/)
  { const char *g = getenv("MODGPU_L2_FETCH");
    size_t gran = g ? (size_t)atoi(g) : 0;
    if (gran == 32 || gran == 64 || gran == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
  }
  ModgpuTable *t = new ModgpuTable();
  t->bits = bits;
  t->slotBits = (uint32_t)(bits - 1);
  t->nSlots = 1ull << t->slotBits;
  t->maxEntries = (1ull << (bits - 2)) - 2;              // reference dies when max >= size = 2^(bits-2)-1
  if (t->maxEntries > MG_MAX_INDEX) t->maxEntries = MG_MAX_INDEX;
  t->numbered = 0;
  t->slots = nullptr; t->dEntries = nullptr; t->dError = nullptr; t->dScratch = nullptr; t->hPinned = nullptr;
  t->clearPending = false; t->dBuckets = nullptr; t->bucketBytes = 0; t->dCursors = nullptr; t->dOverflow = nullptr; t->overflowCap = 0;
  t->bulkOpen = false; t->bulkRoom = t->bulkUsed = t->bulkOvfSeen = 0; t->dCursorSnap = nullptr;
  t->dInfoHi = nullptr; t->infoHiCap = 0; t->trackInfo = false;
  t->scratchWords = t->nSlots / MG_CP_CHUNK + 1024;
  if (mg_check_cuda(cudaMalloc(&t->slots, t->nSlots * sizeof(MgSlot)), "cudaMalloc(table)", __FILE__, __LINE__) ||
      mg_check_cuda(cudaMalloc(&t->dEntries, 64), "cudaMalloc", __FILE__, __LINE__) ||
      mg_check_cuda(cudaMalloc(&t->dScratch, t->scratchWords * sizeof(uint32_t)), "cudaMalloc", __FILE__, __LINE__) ||
      mg_check_cuda(cudaMallocHost(&t->hPinned, 64), "cudaMallocHost", __FILE__, __LINE__))
    { modgpuTableDestroy(t); return nullptr; }
  t->dError = reinterpret_cast<uint32_t *>(t->dEntries + 1);
  if (modgpuTableClear(t, stream)) { modgpuTableDestroy(t); return nullptr; }
  return t;
}

extern "C" void modgpuTableDestroy(ModgpuTable *t)
{
  if (!t) return;
  if (t->slots) cudaFree(t->slots);
  if (t->dEntries) cudaFree(t->dEntries);
  if (t->dScratch) cudaFree(t->dScratch);
  if (t->hPinned) cudaFreeHost(t->hPinned);
  if (t->dBuckets) cudaFree(t->dBuckets);
  if (t->dCursors) cudaFree(t->dCursors);
  if (t->dOverflow) cudaFree(t->dOverflow);
  if (t->dCursorSnap) cudaFree(t->dCursorSnap);
  if (t->dInfoHi) cudaFree(t->dInfoHi);
  delete t;
}

// The clear is lazy: a bulk insert into a logically empty table creates every
// region itself (region_build_kernel<true>), so the separate 16 B/slot clearing
// pass only runs when something else touches the table first.
int mg_table_bulk_close(ModgpuTable *t, cudaStream_t st);

// ... and k-mers still waiting in open buckets (deferred build) are applied before anything reads the table
static int ensure_cleared(ModgpuTable *t, cudaStream_t st)
{
  if (t->bulkOpen) { int rc = mg_table_bulk_close(t, st); if (rc) return rc; }
  if (!t->clearPending) return MODGPU_OK;
  table_clear_kernel<<<grid_for(t->nSlots, 16), 256, 0, st>>>(t->slots, t->nSlots);
  MG_LAUNCH_CHECK("table_clear");
  t->clearPending = false;
  return MODGPU_OK;
}

extern "C" int modgpuTableClear(ModgpuTable *t, void *stream)
{
  cudaStream_t st = (cudaStream_t)stream;
  MG_CUDA(cudaMemsetAsync(t->dEntries, 0, 64, st));
  t->numbered = 0;
  t->clearPending = true;
  t->bulkOpen = false;                                   // k-mers waiting in open buckets are dropped with the rest
  t->trackInfo = false;                                  // an empty set has no info bytes
  return MODGPU_OK;
}

// the info side array covers dense indices 1..n (entry i at [i-1]); new entries start with no flags
static int ensure_info_hi(ModgpuTable *t, uint64_t n, uint64_t have, cudaStream_t st)
{
  if (n > t->infoHiCap)
    { uint8_t *q = nullptr;
      const uint64_t want = n + n / 4 + 4096;
      MG_CUDA(cudaMalloc(&q, want));
      if (t->dInfoHi && have) MG_CUDA(cudaMemcpyAsync(q, t->dInfoHi, have, cudaMemcpyDeviceToDevice, st));
      if (t->dInfoHi) { MG_CUDA(cudaStreamSynchronize(st)); cudaFree(t->dInfoHi); }
      t->dInfoHi = q; t->infoHiCap = want;
    }
  if (n > have) MG_CUDA(cudaMemsetAsync(t->dInfoHi + have, 0, n - have, st));
  return MODGPU_OK;
}

uint8_t *mg_table_info_hi(ModgpuTable *t) { return t->trackInfo ? t->dInfoHi : nullptr; }
// start keeping whole info bytes for the entries numbered so far (no flags yet)
int mg_table_ensure_info(ModgpuTable *t, cudaStream_t st)
{
  if (t->trackInfo) return MODGPU_OK;
  int rc = ensure_info_hi(t, t->numbered, 0, st);
  if (rc) return rc;
  t->trackInfo = true;
  return MODGPU_OK;
}
void mg_table_set_track_info(ModgpuTable *t, bool on) { t->trackInfo = on; }

// ---- bulk insert in three steps, so that hash_select can do the scatter itself:
//   mg_table_bulk_begin   size + zero the per-region buckets for ~expectedN k-mers
//   (scatter)             bucket_scatter_kernel on a list, or hash_select<SCATTER>
//   mg_table_bulk_finish  build every region in shared memory, then the overflow
int mg_table_bulk_close(ModgpuTable *t, cudaStream_t st);

int mg_table_bulk_begin(ModgpuTable *t, uint64_t expectedN, uint64_t maxN, MgBulk *b, cudaStream_t st)
{
  if (t->bulkOpen) { int rc = mg_table_bulk_close(t, st); if (rc) return rc; }    // the buckets are about to be reused
  const uint32_t nRegions = (uint32_t)(t->nSlots >> MG_REGION_BITS);
  uint64_t cap64 = expectedN / nRegions + expectedN / (4ull * nRegions) + 64;
  cap64 = (cap64 + 1) & ~1ull;
  if (cap64 > 0x7FFFFFFFull) { mg_set_error("bulk insert: bucket capacity overflow"); return MODGPU_EINVAL; }
  const uint64_t needBuckets = (uint64_t)nRegions * cap64 * sizeof(uint64_t);
  if (needBuckets > t->bucketBytes)
    { if (t->dBuckets) cudaFree(t->dBuckets);
      t->dBuckets = nullptr; t->bucketBytes = 0;
      MG_CUDA(cudaMalloc(&t->dBuckets, needBuckets + needBuckets / 8));
      t->bucketBytes = needBuckets + needBuckets / 8;
    }
  if (!t->dCursors) MG_CUDA(cudaMalloc(&t->dCursors, ((size_t)nRegions + 16) * sizeof(uint32_t)));
  const uint64_t needOvf = maxN + 1024;                  // a list of maxN can overflow entirely (one hot k-mer)
  if (needOvf > t->overflowCap)
    { if (t->dOverflow) cudaFree(t->dOverflow);
      t->dOverflow = nullptr; t->overflowCap = 0;
      MG_CUDA(cudaMalloc(&t->dOverflow, needOvf * sizeof(uint64_t)));
      t->overflowCap = needOvf;
    }
  MG_CUDA(cudaMemsetAsync(t->dCursors, 0, ((size_t)nRegions + 16) * sizeof(uint32_t), st));
  b->slotBits = t->slotBits; b->regionBits = MG_REGION_BITS; b->nRegions = nRegions; b->cap = (uint32_t)cap64;
  b->expected = expectedN;
  b->cursors = t->dCursors; b->buckets = t->dBuckets; b->overflow = t->dOverflow; b->overflowCap = t->overflowCap;
  return MODGPU_OK;
}

// number of k-mers that missed their bucket in the last scatter (device word, uint32)
const uint32_t *mg_table_bulk_overflow_count(const ModgpuTable *t) { return t->dCursors + (t->nSlots >> MG_REGION_BITS); }

int mg_table_bulk_finish(ModgpuTable *t, const MgBulk *b, cudaStream_t st)
{
  static int variant = -1;
  if (variant < 0) { const char *v = getenv("MODGPU_BUILD_VARIANT"); variant = v ? atoi(v) : 0; }
  const uint32_t guardLimit = b->overflowCap > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)b->overflowCap;
  // a fresh table goes through the persistent pipelined kernel (one streaming write of the table); regions of a
  // populated table are read-modify-written by the block-per-region kernel, whose many short blocks hide the region
  // loads better (93 Gbases in 3-Gbase chunks into an 8 GiB table: 144 ms against 178 ms with the persistent kernel;
  // probing the populated table in place, bucket by bucket, was slower still: 190 ms)
  if (variant == 0 && t->clearPending)
    { // default: the persistent, software-pipelined build (the kernel of the peer-memory exchange) on the one local
      // source.  Same speed as the one-block-per-region kernel when that one is at its best (0.53-0.55 ms), but it
      // stays there: the block-per-region kernel was measured at 1.1 ms on some boxes / days with nothing else changed
      static int blocksPerSm = 0, nw = 0;
      if (!blocksPerSm)
        { const char *v = getenv("MODGPU_BUILD_NW");
          nw = (v && atoi(v) == 16) ? 16 : 8;     // 16 warps measured slower (0.63 against 0.53 ms on the local genome build): kept for A/B
          if (nw == 16) MG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, region_build_pipe_kernel<true, false, 1, 16, 2>, 512, 0));
          else MG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, region_build_pipe_kernel<true, false, 1>, 256, 0));
          if (blocksPerSm < 1) blocksPerSm = 1;
        }
      uint32_t grid = (uint32_t)mg_num_sms() * (uint32_t)blocksPerSm;
      if (grid > b->nRegions) grid = b->nRegions;
      MgPeerSrc src;
      for (uint32_t s2 = 0; s2 < MODGPU_MAX_PEERS; ++s2) src.p[s2] = b->buckets;
#define MG_LOCAL_LAUNCH(FR) do { if (nw == 16) region_build_pipe_kernel<FR, false, 1, 16, 2><<<grid, 512, 0, st>>>(t->slots, t->slotBits, src, b->cursors, b->cap, 1, b->nRegions, b->nRegions, t->dEntries, t->dError, b->cursors + b->nRegions, guardLimit); \
                                 else region_build_pipe_kernel<FR, false, 1><<<grid, 256, 0, st>>>(t->slots, t->slotBits, src, b->cursors, b->cap, 1, b->nRegions, b->nRegions, t->dEntries, t->dError, b->cursors + b->nRegions, guardLimit); } while (0)
      if (t->clearPending) { MG_LOCAL_LAUNCH(true); t->clearPending = false; }
      else MG_LOCAL_LAUNCH(false);
#undef MG_LOCAL_LAUNCH
    }
  else
    {
#define MG_BUILD_LAUNCH(FR, PL) region_build_kernel<FR, PL><<<b->nRegions, 256, 0, st>>>(t->slots, t->slotBits, b->buckets, b->cursors, b->cap, t->dEntries, t->dError, b->cursors + b->nRegions, guardLimit)
  if (t->clearPending)
    { if (variant == 1) MG_BUILD_LAUNCH(true, 4); else if (variant == 2) MG_BUILD_LAUNCH(true, 2); else MG_BUILD_LAUNCH(true, 0);
      t->clearPending = false;
    }
  else
    { if (variant == 1) MG_BUILD_LAUNCH(false, 4); else if (variant == 2) MG_BUILD_LAUNCH(false, 2); else MG_BUILD_LAUNCH(false, 0); }
#undef MG_BUILD_LAUNCH
    }
  MG_LAUNCH_CHECK("region_build");
  // stragglers of over-full buckets go straight into HBM; their number is only known on the device:
  // the insert kernel reads it as the 64-bit word {cursors[nRegions], cursors[nRegions+1] == 0}
  table_insert_kernel<false, true><<<grid_for(65536, 4), 256, 0, st>>>(
      t->slots, t->slotBits, b->overflow, reinterpret_cast<const unsigned long long *>(b->cursors + b->nRegions),
      b->overflowCap, nullptr, t->dEntries, t->dError);
  MG_LAUNCH_CHECK("overflow_insert");
  return MODGPU_OK;
}

// ---- deferred build (modgpuModsetSetAccumulate): the buckets stay open over up to `accum` chunks of about
// `expected` k-mers each, so that a populated table is read-modify-written once per `accum` chunks instead of once per
// chunk (93 Gbases into an 8 GiB table: the per-chunk build moved 16 GiB for 0.4 GB of new k-mers).
//   open      reuse the open buckets when the chunk fits (bucket room and worst-case overflow room), else build what
//             is waiting and start new buckets; snapshots the fill counts so that the chunk can be taken back
//   commit    the chunk is in (the host has seen its overflow counter)
//   rollback  the chunk overflowed the overflow list (pathological skew): restore the fill counts, build what was
//             there before; the caller repeats the chunk through the list path
//   close     build the regions from whatever is waiting
static uint64_t bulk_ovf_need(uint64_t expected) { return 2 * expected + 65536; }

int mg_table_bulk_open(ModgpuTable *t, uint64_t expected, uint32_t accum, MgBulk *b, cudaStream_t st)
{
  const size_t curBytes = ((size_t)(t->nSlots >> MG_REGION_BITS) + 16) * sizeof(uint32_t);
  if (t->bulkOpen && t->bulkUsed + expected <= t->bulkRoom && t->bulkOvfSeen + bulk_ovf_need(expected) <= t->bulk.overflowCap)
    { if (!t->dCursorSnap) MG_CUDA(cudaMalloc(&t->dCursorSnap, curBytes));
      MG_CUDA(cudaMemcpyAsync(t->dCursorSnap, t->dCursors, curBytes, cudaMemcpyDeviceToDevice, st));
      *b = t->bulk;
      return MODGPU_OK;
    }
  if (t->bulkOpen) { int rc = mg_table_bulk_close(t, st); if (rc) return rc; }
  if (accum < 1) accum = 1;
  const uint64_t room = expected * accum;
  int rc = mg_table_bulk_begin(t, room, bulk_ovf_need(expected) + room / 8, b, st);
  if (rc) return rc;
  t->bulk = *b; t->bulkOpen = true; t->bulkRoom = room; t->bulkUsed = 0; t->bulkOvfSeen = 0;
  return MODGPU_OK;
}

void mg_table_bulk_commit(ModgpuTable *t, uint64_t expected, uint64_t ovfSeen) { t->bulkUsed += expected; t->bulkOvfSeen = ovfSeen; }

int mg_table_bulk_rollback(ModgpuTable *t, cudaStream_t st)
{
  if (!t->bulkOpen) return MODGPU_OK;
  if (!t->bulkUsed) { t->bulkOpen = false; return MODGPU_OK; }          // nothing was waiting before this chunk
  const size_t curBytes = ((size_t)(t->nSlots >> MG_REGION_BITS) + 16) * sizeof(uint32_t);
  MG_CUDA(cudaMemcpyAsync(t->dCursors, t->dCursorSnap, curBytes, cudaMemcpyDeviceToDevice, st));
  return mg_table_bulk_close(t, st);
}

int mg_table_bulk_close(ModgpuTable *t, cudaStream_t st)
{
  if (!t->bulkOpen) return MODGPU_OK;
  t->bulkOpen = false;
  if (!t->bulkUsed) return MODGPU_OK;
  return mg_table_bulk_finish(t, &t->bulk, st);
}

bool mg_table_bulk_is_open(const ModgpuTable *t) { return t->bulkOpen; }

// Bulk find-or-insert + count of a long list (count mode, any order): 8 B read +
// 8 B written per k-mer for the scatter, then one streaming pass over the table.
int mg_table_insert_bulk(ModgpuTable *t, const uint64_t *d_kmers, uint64_t n, cudaStream_t st)
{
  if (!n) return MODGPU_OK;
  MgBulk b;
  int rc = mg_table_bulk_begin(t, n, n, &b, st);
  if (rc) return rc;
  bucket_scatter_kernel<<<grid_for(n, 16), 256, 0, st>>>(d_kmers, n, t->slotBits, b.nRegions, b.cap, b.cursors, b.buckets,
                                                         b.overflow, b.overflowCap, t->dError);
  MG_LAUNCH_CHECK("bucket_scatter");
  return mg_table_bulk_finish(t, &b, st);
}

// bulk insert of nSegs device segments with device-side counts (no host round trip)
int mg_table_insert_segments(ModgpuTable *t, const uint64_t *d_segs, uint32_t nSegs, uint64_t segCap,
                             const uint32_t *d_counts, uint64_t expectedN, cudaStream_t st)
{
  if (!nSegs || !segCap) return MODGPU_OK;
  MgBulk b;
  int rc = mg_table_bulk_begin(t, expectedN, (uint64_t)nSegs * segCap, &b, st);
  if (rc) return rc;
  bucket_scatter_seg_kernel<<<grid_for(segCap, 16), 256, 0, st>>>(d_segs, nSegs, segCap, d_counts, t->slotBits, b.nRegions, b.cap,
                                                                  b.cursors, b.buckets, b.overflow, b.overflowCap, t->dError);
  MG_LAUNCH_CHECK("bucket_scatter_seg");
  return mg_table_bulk_finish(t, &b, st);
}

// build every region from nSrc received bucket arrays (each nRegions x cap) + per-source overflow segments
int mg_table_build_from_buckets(ModgpuTable *t, const uint64_t *d_buckets, const uint32_t *d_cursors, uint32_t cap, uint32_t nSrc,
                                const uint64_t *d_overflow, uint64_t overflowCap, const uint32_t *d_ovfCounts, cudaStream_t st)
{
  const uint32_t nRegions = (uint32_t)(t->nSlots >> MG_REGION_BITS);
  if (t->bulkOpen) { int rc = mg_table_bulk_close(t, st); if (rc) return rc; }
  if (t->clearPending)
    { region_build_multi_kernel<true><<<nRegions, 256, 0, st>>>(t->slots, t->slotBits, d_buckets, d_cursors, cap, nSrc, nRegions, t->dEntries, t->dError);
      t->clearPending = false;
    }
  else
    region_build_multi_kernel<false><<<nRegions, 256, 0, st>>>(t->slots, t->slotBits, d_buckets, d_cursors, cap, nSrc, nRegions, t->dEntries, t->dError);
  MG_LAUNCH_CHECK("region_build_multi");
  // the (rare) k-mers that did not fit their bucket at the sender: direct inserts, counts on the device.
  // d_ovfCounts are uint32; widen into the table's scratch so that the insert kernel can read 64-bit counts.
  if (!t->dCursors) MG_CUDA(cudaMalloc(&t->dCursors, ((size_t)nRegions + 16) * sizeof(uint32_t)));
  for (uint32_t s = 0; d_overflow && s < nSrc; ++s)
    { unsigned long long *wide = t->dEntries + 4 + (s & 3);
      MG_CUDA(cudaMemsetAsync(wide, 0, 8, st));
      MG_CUDA(cudaMemcpyAsync(wide, d_ovfCounts + s, 4, cudaMemcpyDeviceToDevice, st));
      table_insert_kernel<false><<<grid_for(65536, 4), 256, 0, st>>>(t->slots, t->slotBits, d_overflow + (uint64_t)s * overflowCap, wide,
                                                                     overflowCap, nullptr, t->dEntries, t->dError);
      MG_LAUNCH_CHECK("overflow_insert");
    }
  return MODGPU_OK;
}

// build every region from the buckets in nSrc ranks' memory (peer-mapped) + their overflow segments.
// d_cursors: [nSrc][cursorStride] fill counts (cursorStride >= nRegions; 0 = nRegions); d_ovfCounts[s * ovfStride];
// d_guard (nullable): a device word > 0 means "some rank lost k-mers of this group": build nothing (a fresh table is
// created empty) - every rank sees the same flags, so the group is applied everywhere or nowhere.
// the overflow segments of ALL sources in one launch (one count word and one direct-insert launch per source were 16
// serialised 6-microsecond kernels behind the build at 8 GPUs: 0.1 ms of a 0.84 ms build)
__global__ void __launch_bounds__(256) peer_overflow_insert_kernel(MgSlot *slots, uint32_t slotBits, const MgPeerSrc ovf,
                                                                   const uint32_t *__restrict__ counts, uint64_t ovfStride, uint32_t nSrc,
                                                                   uint64_t cap, const uint32_t *__restrict__ guard,
                                                                   unsigned long long *entries, uint32_t *error)
{
  if (guard && __ldg(guard) > 0u) return;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint32_t fresh = 0;
  for (uint32_t s = 0; s < nSrc; ++s)
    { uint64_t n = __ldg(counts + s * ovfStride);
      if (n > cap) n = cap;
      const uint64_t *kmers = ovf.p[s];
      for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        { const uint64_t key = kmers[i] & 0x3FFFFFFFFFFFFFFFull;
          bool isNew;
          const uint64_t sl = probe_insert(slots, slotBits, key, &isNew);
          if (sl == 0xFFFFFFFFFFFFFFFFull) { atomicExch(error, 1u); continue; }
          fresh += isNew ? 1u : 0u;
          atomicAdd(&slots[sl].count, 1u);
        }
    }
  fresh = mg_warp_sum(fresh);
  if (mg_lane() == 0 && fresh) atomicAdd(entries, (unsigned long long)fresh);
}

int mg_table_build_from_peers_ex(ModgpuTable *t, const uint64_t *const *d_buckets, const uint32_t *d_cursors, uint64_t cursorStride,
                                 uint32_t cap, uint32_t nSrc, const uint64_t *const *d_overflow, uint64_t overflowCap,
                                 const uint32_t *d_ovfCounts, uint64_t ovfStride, const uint32_t *d_guard, cudaStream_t st)
{
  if (nSrc < 1 || nSrc > MODGPU_MAX_PEERS) { mg_set_error("build from peers: %u sources out of range 1..%d", nSrc, MODGPU_MAX_PEERS); return MODGPU_EINVAL; }
  if (t->bulkOpen) { int rc = mg_table_bulk_close(t, st); if (rc) return rc; }
  const uint32_t nRegions = (uint32_t)(t->nSlots >> MG_REGION_BITS);
  if (!cursorStride) cursorStride = nRegions;
  MgPeerSrc src;
  for (uint32_t s = 0; s < MODGPU_MAX_PEERS; ++s) src.p[s] = d_buckets[s < nSrc ? s : 0];
  static int blocksPerSm = 0, nw = 0;
  if (!blocksPerSm)
    { const char *v = getenv("MODGPU_BUILD_NW");
      nw = (v && atoi(v) == 16) ? 16 : 8;     // 16 warps measured slower (0.63 against 0.53 ms on the local genome build): kept for A/B
      if (nw == 16) MG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, region_build_pipe_kernel<true, true, 1, 16, 2>, 512, 0));
      else MG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, region_build_pipe_kernel<true, true, 1>, 256, 0));
      if (blocksPerSm < 1) blocksPerSm = 1;
    }
  uint32_t grid = (uint32_t)mg_num_sms() * (uint32_t)blocksPerSm;
  if (grid > nRegions) grid = nRegions;
#define MG_PIPE_ARGS t->slots, t->slotBits, src, d_cursors, cap, nSrc, cursorStride, nRegions, t->dEntries, t->dError, d_guard, 0u, 1
#define MG_PIPE_LAUNCH(FR) do { if (nw == 16) region_build_pipe_kernel<FR, true, 1, 16, 2><<<grid, 512, 0, st>>>(MG_PIPE_ARGS); \
                                else if (nSrc <= 8) region_build_pipe_kernel<FR, true, 1><<<grid, 256, 0, st>>>(MG_PIPE_ARGS); \
                                else region_build_pipe_kernel<FR, true, 2><<<grid, 256, 0, st>>>(MG_PIPE_ARGS); } while (0)
  static int useTma = -1, tmaBlocks = 0, tmaHint = 0;
  if (useTma < 0)
    { const char *v = getenv("MODGPU_PEER_TMA"); useTma = v ? atoi(v) : 1;
      v = getenv("MODGPU_PEER_HINT"); tmaHint = v ? atoi(v) : 0;
    }
  const size_t ringBytes = 2 * (size_t)nSrc * cap * 8;
  bool aligned = (cap & 1u) == 0;                            // bulk copies want 16-byte aligned buckets
  for (uint32_t q = 0; q < nSrc; ++q) aligned = aligned && (reinterpret_cast<uintptr_t>(d_buckets[q]) & 15u) == 0;
  if (useTma && aligned && nSrc >= 2 && nSrc <= 8 && ringBytes <= 96 * 1024)
    { // the transfer on the TMA engine (bulk copies from the peers into a shared-memory ring)
      static size_t ringSet = 0;
      if (ringBytes > ringSet)
        { MG_CUDA(cudaFuncSetAttribute(region_build_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ringBytes));
          MG_CUDA(cudaFuncSetAttribute(region_build_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ringBytes));
          MG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&tmaBlocks, region_build_tma_kernel<true>, 256, ringBytes));
          if (tmaBlocks < 1) tmaBlocks = 1;
          ringSet = ringBytes;
        }
      uint32_t tgrid = (uint32_t)mg_num_sms() * (uint32_t)tmaBlocks;
      if (tgrid > nRegions) tgrid = nRegions;
      if (t->clearPending)
        { region_build_tma_kernel<true><<<tgrid, 256, ringBytes, st>>>(t->slots, t->slotBits, src, d_cursors, cap, nSrc, cursorStride, nRegions, t->dEntries, t->dError, d_guard, (uint32_t)tmaHint);
          t->clearPending = false;
        }
      else
        region_build_tma_kernel<false><<<tgrid, 256, ringBytes, st>>>(t->slots, t->slotBits, src, d_cursors, cap, nSrc, cursorStride, nRegions, t->dEntries, t->dError, d_guard, (uint32_t)tmaHint);
    }
  else if (t->clearPending) { MG_PIPE_LAUNCH(true); t->clearPending = false; }
  else MG_PIPE_LAUNCH(false);
#undef MG_PIPE_LAUNCH
#undef MG_PIPE_ARGS
  MG_LAUNCH_CHECK("region_build_peer");
  // the (rare) k-mers that did not fit their bucket at the sender: direct inserts reading the peer's segment
  if (d_overflow)
    { MgPeerSrc ovf;
      for (uint32_t s = 0; s < MODGPU_MAX_PEERS; ++s) ovf.p[s] = d_overflow[s < nSrc ? s : 0];
      peer_overflow_insert_kernel<<<grid_for(65536, 4), 256, 0, st>>>(t->slots, t->slotBits, ovf, d_ovfCounts, ovfStride, nSrc, overflowCap, d_guard,
                                                                      t->dEntries, t->dError);
      MG_LAUNCH_CHECK("overflow_insert");
    }
  return MODGPU_OK;
}

int mg_table_build_from_peers(ModgpuTable *t, const uint64_t *const *d_buckets, const uint32_t *d_cursors, uint32_t cap, uint32_t nSrc,
                              const uint64_t *const *d_overflow, uint64_t overflowCap, const uint32_t *d_ovfCounts, cudaStream_t st)
{ return mg_table_build_from_peers_ex(t, d_buckets, d_cursors, 0, cap, nSrc, d_overflow, overflowCap, d_ovfCounts, 1, nullptr, st); }

bool mg_table_clear_pending(const ModgpuTable *t) { return t->clearPending; }
// the guarded build kernels of mg_table_bulk_finish did nothing (overflow list too small): undo the host bookkeeping
void mg_table_bulk_abort(ModgpuTable *t, bool wasPending) { t->clearPending = wasPending; }

void mg_table_counters(ModgpuTable *t, unsigned long long **entries, uint32_t **error) { *entries = t->dEntries; *error = t->dError; }
uint32_t mg_table_regions(const ModgpuTable *t) { return (uint32_t)(t->nSlots >> MG_REGION_BITS); }
uint32_t mg_table_slot_bits(const ModgpuTable *t) { return t->slotBits; }

uint64_t mg_table_bulk_threshold(const ModgpuTable *t) { return t->nSlots / 8; }

extern "C" uint64_t modgpuTableSlots(const ModgpuTable *t) { return t->nSlots; }
extern "C" void *modgpuTableDevicePtr(const ModgpuTable *t)
{
  if (t->clearPending || t->bulkOpen)
    { ensure_cleared(const_cast<ModgpuTable *>(t), 0);
      cudaStreamSynchronize(0);
    }
  return t->slots;
}

// internal: the slot array for kernels launched on `st` (the pending clear / deferred build run on the same stream:
// ordered before them without a device-wide synchronisation)
MgSlot *mg_table_slots_on(ModgpuTable *t, cudaStream_t st)
{
  if ((t->clearPending || t->bulkOpen) && ensure_cleared(t, st)) return nullptr;
  return t->slots;
}

// internal: insert with the element count taken from device memory
int mg_table_insert_dev(ModgpuTable *t, const uint64_t *d_kmers, const uint64_t *d_n, uint64_t nMax,
                        uint32_t *d_slot, int exactOrder, cudaStream_t st)
{
  if (!nMax) return MODGPU_OK;
  { int rc = ensure_cleared(t, st); if (rc) return rc; }
  if (exactOrder && nMax >= (1ull << 30) - 2)
    { mg_set_error("exact-order insert batch of %llu exceeds 2^30", (unsigned long long)nMax); return MODGPU_EINVAL; }
  // two waves of resident blocks measured faster than one on B200 for this
  // latency-bound probe loop (profiles/README.md, insert sweep)
  unsigned grid = grid_for(nMax, 16);
  if (exactOrder == 2)                                   // find-or-insert without counting
    table_insert_kernel<true, false, false><<<grid, 256, 0, st>>>(t->slots, t->slotBits, d_kmers, (const unsigned long long *)d_n, nMax, d_slot, t->dEntries, t->dError);
  else if (exactOrder)
    table_insert_kernel<true><<<grid, 256, 0, st>>>(t->slots, t->slotBits, d_kmers, (const unsigned long long *)d_n, nMax, d_slot, t->dEntries, t->dError);
  else
    table_insert_kernel<false><<<grid, 256, 0, st>>>(t->slots, t->slotBits, d_kmers, (const unsigned long long *)d_n, nMax, d_slot, t->dEntries, t->dError);
  MG_LAUNCH_CHECK("table_insert");
  return MODGPU_OK;
}

extern "C" int modgpuTableInsert(ModgpuTable *t, const uint64_t *d_kmers, uint64_t n, uint32_t *d_slot,
                                 int exactOrder, void *stream)
{
  return mg_table_insert_dev(t, d_kmers, nullptr, n, d_slot, exactOrder, (cudaStream_t)stream);
}

extern "C" uint64_t modgpuTableEntries(ModgpuTable *t, void *stream)
{
  cudaStream_t st = (cudaStream_t)stream;
  if (t->bulkOpen && mg_table_bulk_close(t, st)) return 0xFFFFFFFFFFFFFFFFull;
  if (mg_check_cuda(cudaMemcpyAsync(t->hPinned, t->dEntries, 16, cudaMemcpyDeviceToHost, st), "entries readback", __FILE__, __LINE__) ||
      mg_check_cuda(cudaStreamSynchronize(st), "cudaStreamSynchronize", __FILE__, __LINE__))
    return 0xFFFFFFFFFFFFFFFFull;
  uint64_t entries = t->hPinned[0];
  uint32_t err = (uint32_t)(t->hPinned[1] & 0xFFFFFFFFu);
  if (err || entries > t->maxEntries)
    { // reference: die("hashTableSize %u is too small for %u"), modset.c:58
      mg_set_error("hashTableSize %llu is too small for %llu (table bits %d)%s",
                   (unsigned long long)(t->maxEntries + 1), (unsigned long long)entries, t->bits,
                   err == 2 ? " [duplicate key on import]" : "");
      return 0xFFFFFFFFFFFFFFFFull;
    }
  return entries;
}

template <class F>
static int run_compaction(ModgpuTable *t, F f, uint64_t n, uint64_t *totalOut, cudaStream_t st)
{
  uint64_t chunks = (n + MG_CP_CHUNK - 1) / MG_CP_CHUNK;
  if (chunks > t->scratchWords)
    { if (t->dScratch) cudaFree(t->dScratch);
      t->dScratch = nullptr;
      t->scratchWords = chunks + 1024;
      MG_CUDA(cudaMalloc(&t->dScratch, t->scratchWords * sizeof(uint32_t)));
    }
  unsigned long long *dTotal = t->dEntries + 2;
  int rc = mg_ordered_scan(f, n, t->dScratch, dTotal, st);
  if (rc) return rc;
  MG_CUDA(cudaMemcpyAsync(t->hPinned + 2, dTotal, 8, cudaMemcpyDeviceToHost, st));
  MG_CUDA(cudaStreamSynchronize(st));
  *totalOut = t->hPinned[2];
  return MODGPU_OK;
}

extern "C" int modgpuTableNumber(ModgpuTable *t, const uint32_t *d_slot, uint64_t n, uint32_t *d_index, void *stream)
{
  cudaStream_t st = (cudaStream_t)stream;
  { int rc = ensure_cleared(t, st); if (rc) return rc; }
  uint64_t added = 0;
  if (d_slot)
    { if (n)
        { FlagFirstOccurrence f; f.slots = t->slots; f.slotOf = d_slot; f.base = t->numbered;
          int rc = run_compaction(t, f, n, &added, st);
          if (rc) return rc;
        }
    }
  else
    { FlagNewSlot f; f.slots = t->slots; f.base = t->numbered;
      int rc = run_compaction(t, f, t->nSlots, &added, st);
      if (rc) return rc;
    }
  if (t->trackInfo && added) { int rc = ensure_info_hi(t, t->numbered + added, t->numbered, st); if (rc) return rc; }
  t->numbered += added;
  if (t->numbered > t->maxEntries)
    { mg_set_error("hashTableSize %llu is too small for %llu (table bits %d)",
                   (unsigned long long)(t->maxEntries + 1), (unsigned long long)t->numbered, t->bits);
      return MODGPU_EFULL;
    }
  if (d_slot && d_index && n)
    { gather_index_kernel<<<grid_for(n, 16), 256, 0, st>>>(t->slots, d_slot, n, d_index);
      MG_LAUNCH_CHECK("gather_index");
    }
  return MODGPU_OK;
}

uint64_t mg_table_numbered(const ModgpuTable *t) { return t->numbered; }

int mg_table_lookup_dev(const ModgpuTable *t, const uint64_t *d_kmers, const uint64_t *d_n, uint64_t nMax,
                        uint32_t *d_out, cudaStream_t st)
{
  if (!nMax) return MODGPU_OK;
  { int rc = ensure_cleared(const_cast<ModgpuTable *>(t), st); if (rc) return rc; }
  table_lookup_kernel<<<grid_for(nMax, 16), 256, 0, st>>>(t->slots, t->slotBits, d_kmers, (const unsigned long long *)d_n, nMax, d_out);
  MG_LAUNCH_CHECK("table_lookup");
  return MODGPU_OK;
}

extern "C" int modgpuTableLookup(const ModgpuTable *t, const uint64_t *d_kmers, uint64_t n, uint32_t *d_out, void *stream)
{
  return mg_table_lookup_dev(t, d_kmers, nullptr, n, d_out, (cudaStream_t)stream);
}

extern "C" int modgpuTableHistogram(const ModgpuTable *t, uint32_t *d_bins65536, void *stream)
{
  cudaStream_t st = (cudaStream_t)stream;
  MG_CUDA(cudaMemsetAsync(d_bins65536, 0, 65536 * sizeof(uint32_t), st));
  { int rc = ensure_cleared(const_cast<ModgpuTable *>(t), st); if (rc) return rc; }
  table_hist_kernel<<<grid_for(t->nSlots, 8), 256, 0, st>>>(t->slots, t->nSlots, d_bins65536);
  MG_LAUNCH_CHECK("table_hist");
  return MODGPU_OK;
}

// mode: 0 thresholds (-s), 1 copyM only (-sM), 2 exact multiplicity (modmap), 3 tally only
int mg_table_classify(ModgpuTable *t, int mode, int c1, int c2, int cM, int zeroDepth, uint32_t *d_classCounts, cudaStream_t st)
{
  if (d_classCounts) MG_CUDA(cudaMemsetAsync(d_classCounts, 0, 4 * sizeof(uint32_t), st));
  { int rc = ensure_cleared(t, st); if (rc) return rc; }
  table_classify_kernel<<<grid_for(t->nSlots, 8), 256, 0, st>>>(t->slots, t->nSlots, mode, c1, c2, cM, zeroDepth, d_classCounts);
  MG_LAUNCH_CHECK("table_classify");
  return MODGPU_OK;
}

extern "C" int modgpuTableClassify(ModgpuTable *t, int c1, int c2, int cM, int exact, uint32_t *d_classCounts, void *stream)
{
  int mode = exact ? 2 : (c1 < 0 && c2 < 0) ? 1 : 0;
  return mg_table_classify(t, mode, c1, c2, cM, 0, d_classCounts, (cudaStream_t)stream);
}

extern "C" int modgpuTableExport(ModgpuTable *t, uint64_t *d_value, uint16_t *d_depth, uint8_t *d_info,
                                 uint32_t *d_count32, void *stream)
{
  { int rc = ensure_cleared(t, (cudaStream_t)stream); if (rc) return rc; }
  table_export_kernel<<<grid_for(t->nSlots, 8), 256, 0, (cudaStream_t)stream>>>(t->slots, t->nSlots, d_value, d_depth, d_info, d_count32,
                                                                                t->trackInfo ? t->dInfoHi : nullptr);
  MG_LAUNCH_CHECK("table_export");
  return MODGPU_OK;
}

extern "C" int modgpuTableImport(ModgpuTable *t, const uint64_t *d_value, const uint16_t *d_depth,
                                 const uint8_t *d_info, uint64_t n, void *stream)
{
  if (!n) return MODGPU_OK;
  if (t->numbered + n > t->maxEntries)
    { mg_set_error("Modset size %llu is too big for %d bits", (unsigned long long)(t->numbered + n), t->bits); return MODGPU_EFULL; }
  { int rc = ensure_cleared(t, (cudaStream_t)stream); if (rc) return rc; }
  // an info array brings the whole byte: the flags beyond the copy bits live in the side array from here on
  if (d_info && !t->trackInfo)
    { int rc = ensure_info_hi(t, t->numbered, 0, (cudaStream_t)stream); if (rc) return rc;
      t->trackInfo = true;
    }
  if (t->trackInfo) { int rc = ensure_info_hi(t, t->numbered + n, t->numbered, (cudaStream_t)stream); if (rc) return rc; }
  table_import_kernel<<<grid_for(n, 16), 256, 0, (cudaStream_t)stream>>>(t->slots, t->slotBits, d_value, d_depth, d_info, n, t->numbered, t->dEntries, t->dError,
                                                                         t->trackInfo ? t->dInfoHi : nullptr);
  MG_LAUNCH_CHECK("table_import");
  t->numbered += n;
  return MODGPU_OK;
}
