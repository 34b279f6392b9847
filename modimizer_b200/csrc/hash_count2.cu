// hash_count2.cu - K1 + K2 (+ the bucket scatter of K3) in count mode, second generation.
//
// replaces, like hash_select.cu: modRCiterator / modRCnext / advanceHashRC / hashRC of the reference
// (seqhash.c:60-79,154-196) and the byte -> 2-bit conversion of seqIOread (seqio.c:322,328-331) for a whole batch;
// with OUT = 1 / 3 also the bucket scatter in front of the table build (table.cu).
//
// Same decomposition as hash_count_kernel (every WARP is autonomous: tiles of 2048 window starts from raw bytes,
// staged by one TMA bulk copy per tile behind a per-warp mbarrier; candidate masks per lane; a warp-level queue so
// that the per-candidate work is spread evenly over the lanes), rebuilt around the instruction counts the round-1
// profiles showed (ncu r01 v6: 12.8 issued instructions per base, a quarter of them tile bookkeeping; the full scan
// of a generic d: 37.9 per base of which 14 outside the scan itself):
//   - end flags are SPARSE: one byte per tile says whether any sequence ends in its reach; the 1-bit-per-base array
//     is neither copied nor read for the other tiles (a 3.1 Gb genome in 24 records: 387 MB of zeros per pass, three
//     shared-memory loads per lane and tile, and the second bulk copy with its issue sequence).  Batches of short
//     reads keep the staged flags (ENDS);
//   - the tile is read TRANSPOSED: lane l packs the 16-byte chunks l, l+32, l+64, l+96 (conflict-free as they lie)
//     and stores the four half-words where the packed tile wants them, instead of rotating its reads and the results;
//   - the queue offsets come from three ballots of the (small) per-lane counts instead of a shuffle scan, and the
//     masks are drained from the top bit (FLO alone, no BREV);
//   - evaluation and output run round by round with the bucket store one round behind its atomic, so that the L2
//     round trip of the position is hidden behind the next evaluation;
//   - a queue entry is evaluated in 32-bit pieces without a branch (mg_eval32_single: funnel-shifted window, the
//     reverse complement from two BREVs, three multiply-adds per 64-bit product);
//   - the region of a selected k-mer is one shift of the high product word.
// The scans are the first generation's: the 16 KiB candidate table (k >= 30, d = 2^t * odd with 64-2k+t <= 8: the
// modmap configuration k=31 d=64) and the full evaluation of every window in 32-bit pieces (any d, k >= 16: the
// reference's default k=19 d=31).  Everything else (k < 16, the arithmetic prefilter, packed input, per-owner
// segments) stays with hash_count_kernel, which mg_count2_launch reports by returning 1.
#include <string.h>
#include <stdlib.h>
#include "mg_select.cuh"


// resident blocks per SM the kernels are compiled for (register budget = 65536 / (blocks x threads))
#ifndef C2_LUT_BLOCKS
#define C2_LUT_BLOCKS 3
#endif
#ifndef C2_GEN_BLOCKS
#define C2_GEN_BLOCKS 2
#endif

// warps per block: measured on B200 (profiles/geometry_r02.txt), 3.1 Gbases.  Table-driven kernel, 3 blocks per SM:
// 10 / 11 / 12 / 13 / 14 / 15 / 16 warps -> 1.254 / 1.223 / 1.203 / 1.192 / 1.193 / 1.182 / 1.257 ms (16 warps need the
// 228 KiB shared-memory carve-out, up to 15 fit the 196 KiB one and leave 60 KiB of L1); full scan, 2 blocks per SM:
// 12 / 13 / 14 / 15 / 16 warps -> 3.648 / 3.609 / 3.592 / 3.667 / 3.656 ms (72 registers at 14 warps)
#ifndef C2_LUT_WARPS
#define C2_LUT_WARPS 15
#endif
#ifndef C2_GEN_WARPS
#define C2_GEN_WARPS 14
#endif
#ifndef C2_WQ_CAP
#define C2_WQ_CAP MG_WQ_CAP                                    // queue entries per warp
#endif
#define C2_WARPS(SCAN) ((SCAN) ? C2_LUT_WARPS : C2_GEN_WARPS)
#define C2_BOUNDS(SCAN) __launch_bounds__(C2_WARPS(SCAN) * 32, SCAN ? C2_LUT_BLOCKS : C2_GEN_BLOCKS)

template <bool ENDS> struct C2WarpSmem {
  __align__(16) uint8_t stage[MG_WS_RAW_BYTES];                // the raw tile + 32 bytes of overlap (TMA destination)
  __align__(16) uint32_t ends[ENDS ? MG_WS_ENDS_BYTES / 4 : 4];
  __align__(16) uint32_t half[2 * (MG_WT_RUNS + 2)];           // the packed tile as 32-bit halves (word w = half[2w+1] : half[2w])
  __align__(8) uint64_t bar;
  uint16_t queue[C2_WQ_CAP];
};

__device__ __forceinline__ uint32_t c2_top_bit(uint32_t x)     // position of the highest set bit (x != 0): FLO
{
  uint32_t r;
  asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}

template <bool ENDS>
__device__ __forceinline__ void c2_issue_tile(const SelectParams &P, C2WarpSmem<ENDS> *S, uint64_t tile)
{ // one elected lane: the bulk copies of a tile signal the warp's mbarrier
  mg_mbar_expect_tx(&S->bar, MG_WS_RAW_BYTES + (ENDS ? MG_WS_ENDS_BYTES : 0));
  mg_tma_load_1d_hint(S->stage, P.raw + tile * MG_WT_BASES, MG_WS_RAW_BYTES, &S->bar, MG_L2_EVICT_FIRST);
  if (ENDS) mg_tma_load_1d_hint(S->ends, P.ends + tile * MG_WT_RUNS, MG_WS_ENDS_BYTES, &S->bar, MG_L2_EVICT_FIRST);
}

// a selected k-mer goes to position pos of bucket `region` (pos from the bucket's cursor), or - bucket full - to the
// overflow list (PEER: of its owner).  The address is built from byte offsets; PEER buckets are many and
// small: their stores ask L2 to keep the line until its four k-mers have arrived (evict-last)
template <bool PEER>
__device__ __forceinline__ void c2_place(const SelectParams &P, uint32_t region, uint32_t pos, uint64_t km)
{
  if (pos < P.bucketCap)
    { char *row = reinterpret_cast<char *>(P.buckets) + (uint64_t)region * (P.bucketCap * 8u);       // (cap < 2^29)
      uint64_t *bp = reinterpret_cast<uint64_t *>(row + (uint64_t)pos * 8u);
      if (PEER) mg_st_keep(bp, km); else *bp = km;
    }
  else if (PEER)
    { const uint32_t ow = region / P.nRegions;
      const uint32_t o = atomicAdd(&P.ownerCursor[ow], 1u);
      if (o < P.overflowCap) P.overflow[(uint64_t)ow * P.overflowCap + o] = km;
    }
  else
    { const uint32_t o = atomicAdd(&P.cursors[P.nRegions], 1u);
      if (o < P.overflowCap) P.overflow[o] = km;
    }
}

// SCAN: 0 = every window evaluated in full (32-bit pieces, k >= 16; LUTK = 1 when d is odd: no low-bit test),
//       1 = table-driven candidates (LUTK = k)
// OUT:  0 = list, 1 = the table's region buckets, 3 = per-(owner, region) buckets
//       P2 (table scan only): d is a power of two - no odd-part test in the evaluation
template <int SCAN, int LUTK, int OUT, bool ASCII, bool ENDS, bool P2>
__global__ void C2_BOUNDS(SCAN) hash_count2_kernel(const SelectParams P)
{
  constexpr bool SCATTER = (OUT == 1 || OUT == 3);
  constexpr bool PEER = (OUT == 3);
  extern __shared__ __align__(128) uint8_t sDyn[];
  uint8_t *sLut = sDyn;                                                            // MG_LUT_SIZE bytes when SCAN
  C2WarpSmem<ENDS> *S = reinterpret_cast<C2WarpSmem<ENDS> *>(sDyn + (SCAN ? MG_LUT_SIZE : 0)) + (threadIdx.x >> 5);

  const MgKHasher &H = P.H;
  const uint32_t tid = threadIdx.x, lane = tid & 31;
  const MgEval32 &E = P.E;                                 // prepared on the host: constant-bank operands, nothing to rebuild per round
  // region of a k-mer = the top (slotBits - regionBits) bits of its slot hash = one shift of the high product word
  const uint32_t regionShift = P.regionShift;
  // tile schedule (as in the first generation): chunks of MG_CNT_CHUNK consecutive warp tiles, the first by warp index,
  // the following ones from an atomic ticket requested a whole chunk ahead
  constexpr uint32_t WARPS = C2_WARPS(SCAN);
  const uint32_t nWarps = gridDim.x * WARPS;
  uint64_t tile = (uint64_t)(blockIdx.x * WARPS + (tid >> 5)) * MG_CNT_CHUNK, tileNext = 0;
  uint32_t pendingChunk = 0;
  if (lane == 0) pendingChunk = nWarps + atomicAdd(P.ticket, 1u);
  // a tile can be bulk-copied when all of it and its overlap lie inside the batch (the raw bytes carry no slack)
  const uint64_t nBulk = P.nBases >= MG_WS_RAW_BYTES ? (P.nBases - MG_WS_RAW_BYTES) / MG_WT_BASES + 1 : 0;
  uint32_t nSelectedLocal = 0;
  uint32_t phase = 0;
  // scatter: the k-mer whose bucket position is still on its way (stored one round later, also across tiles)
  uint32_t pKl = 0, pKh = 0, pRegion = 0, pPos = 0;
  uint32_t pOn = 0;                                         // (a word, not a bool: no byte packing around the predicate)

  if (lane == 0)
    { mg_mbar_init(&S->bar, 1);
      mg_fence_barrier_init();
      mg_fence_proxy_async();
      if (tile < nBulk) c2_issue_tile<ENDS>(P, S, tile);
    }
  if (SCAN)
    { const uint4 *src = reinterpret_cast<const uint4 *>(P.lut);
      uint4 *dst = reinterpret_cast<uint4 *>(sLut);
      for (uint32_t i = tid; i < MG_LUT_SIZE / 16; i += WARPS * 32) dst[i] = __ldg(src + i);
    }
  __syncthreads();

  for (; tile < P.nTiles; tile = tileNext)
    { const uint32_t run0 = lane * 2;
      const uint64_t tileBase = tile * MG_WT_BASES;
      uint32_t tflag = 0;
      if (!ENDS) tflag = __ldg(P.tileFlags + tile);         // warp-uniform; consumed after the scan
      uint32_t e0 = 0, e1 = 0, e2 = 0, bl0 = 0, bl1 = 0;
      uint64_t *W = reinterpret_cast<uint64_t *>(S->half);
      if (tile < nBulk)
        { mg_mbar_wait(&S->bar, phase);
          phase ^= 1;
          // K1, transposed: chunk c = 16 bases = one half-word; lanes read consecutive chunks and put them in place
          const uint4 *src = reinterpret_cast<const uint4 *>(S->stage);
          const uint32_t h0 = pack16_dev<ASCII>(src[lane]), h1 = pack16_dev<ASCII>(src[32 + lane]),
                         h2 = pack16_dev<ASCII>(src[64 + lane]), h3 = pack16_dev<ASCII>(src[96 + lane]);
          const uint32_t at = lane ^ 1u;                     // little-endian halves of a 64-bit word: the first chunk is the high one
          S->half[at] = h0; S->half[32 + at] = h1; S->half[64 + at] = h2; S->half[96 + at] = h3;
          if (lane < 2) S->half[128 + at] = pack16_dev<ASCII>(src[128 + lane]);
          if (ENDS)
            { // the staged flags are turned into the blocked masks HERE, before the warp barrier that frees the staging
              // buffers for the next bulk copy: a load the compiler sinks below the copy's issue would read the next tile
              e0 = S->ends[run0]; e1 = S->ends[run0 + 1]; e2 = S->ends[run0 + 2];
              bl0 = mg_blocked_mask((uint64_t)e0 | ((uint64_t)e1 << 32), H.k);
              bl1 = mg_blocked_mask((uint64_t)e1 | ((uint64_t)e2 << 32), H.k);
            }
        }
      else
        { // the ragged end of the batch: guarded loads
          const uint64_t word = tile * MG_WT_RUNS + run0;
          W[run0] = pack32_raw<ASCII>(P.raw, word * MG_RUN, P.nBases);
          W[run0 + 1] = pack32_raw<ASCII>(P.raw, word * MG_RUN + 32, P.nBases);
          if (lane == 31) W[MG_WT_RUNS] = pack32_raw<ASCII>(P.raw, word * MG_RUN + 64, P.nBases);
          e0 = __ldg(P.ends + word); e1 = __ldg(P.ends + word + 1); e2 = __ldg(P.ends + word + 2);
          bl0 = mg_blocked_mask((uint64_t)e0 | ((uint64_t)e1 << 32), H.k);
          bl1 = mg_blocked_mask((uint64_t)e1 | ((uint64_t)e2 << 32), H.k);
          tflag = 1;
        }
      asm volatile("" : "+r"(bl0), "+r"(bl1) :: "memory");      // (the masks exist in registers before the barrier)
      __syncwarp();
      // every lane has consumed the staged tile: the next tile's copy streams in behind the computation of this one
      tileNext = tile + 1;
      if ((tileNext & (MG_CNT_CHUNK - 1)) == 0)
        { tileNext = (uint64_t)__shfl_sync(0xffffffffu, pendingChunk, 0) * MG_CNT_CHUNK;
          if (lane == 0 && tileNext < P.nTiles) pendingChunk = nWarps + atomicAdd(P.ticket, 1u);
        }
      if (lane == 0 && tileNext < nBulk) c2_issue_tile<ENDS>(P, S, tileNext);
      const uint4 w01 = *reinterpret_cast<const uint4 *>(S->half + 2 * run0);       // words run0, run0 + 1
      const uint2 w2h = *reinterpret_cast<const uint2 *>(S->half + 2 * run0 + 4);   // word run0 + 2
      // 32-bit halves of the three words, most significant first: a:b = w0, c:d = w1, e:f = w2
      const uint32_t a = w01.y, b = w01.x, c = w01.w, d = w01.z, e = w2h.y, f = w2h.x;

      uint32_t m0, m1;
      if (SCAN)
        mg_lut_scan<LUTK ? LUTK : 31>(sLut, ((uint64_t)a << 32) | b, ((uint64_t)c << 32) | d, ((uint64_t)e << 32) | f, &m0, &m1);
      else
        { // ONE unrolled copy of the scan, executed for both runs (two copies do not fit the instruction cache)
          uint64_t wa = ((uint64_t)a << 32) | b, wb = ((uint64_t)c << 32) | d;
          const uint64_t wc = ((uint64_t)e << 32) | f;
          m0 = 0; m1 = 0;
#pragma unroll 1
          for (int h = 0; h < 2; ++h)
            { const MgRun RR = mg_run_prepare(wa, wb, H.k);
              const MgRun32 Q = mg_run32(RR);
              uint32_t mh = 0;
#pragma unroll
              for (int i = 0; i < MG_RUN / 2; ++i)             // windows i and i + 16 share their middle funnel shifts
                { bool s0, s16;
                  mg_selected32_pair<LUTK == 1>(E, Q, i, &s0, &s16);
                  asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p or.b32 %0, %0, %2;\n\t}" : "+r"(mh) : "r"((uint32_t)s0), "r"(1u << i));
                  asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p or.b32 %0, %0, %2;\n\t}" : "+r"(mh) : "r"((uint32_t)s16), "r"(1u << (i + 16)));
                }
              if (h == 0) m0 = mh; else m1 = mh;
              wa = wb; wb = wc;
            }
        }
      // windows that would span two sequences or run off the batch are not usable
      { const bool inside = tileBase + MG_WT_BASES + MG_RUN <= P.nBases;
        if (!ENDS && tflag && tile < nBulk)                  // sparse flags: this tile has a sequence end in reach (warp-uniform)
          { const uint64_t word = tile * MG_WT_RUNS + run0;
            e0 = __ldg(P.ends + word); e1 = __ldg(P.ends + word + 1); e2 = __ldg(P.ends + word + 2);
            bl0 = mg_blocked_mask((uint64_t)e0 | ((uint64_t)e1 << 32), H.k);
            bl1 = mg_blocked_mask((uint64_t)e1 | ((uint64_t)e2 << 32), H.k);
          }
        m0 &= ~bl0; m1 &= ~bl1;
        if (!inside)                                         // the last tiles: window starts beyond the batch
          { const uint64_t p0 = tileBase + (uint64_t)run0 * MG_RUN;
            m0 &= mg_run_usable(0ull, H.k, p0, P.nBases);
            m1 &= mg_run_usable(0ull, H.k, p0 + MG_RUN, P.nBases);
          }
      }

      // ---- the warp's queue of (run, window).  Order is irrelevant in count mode: a lane's entries go behind those of the
      // lower lanes; the prefix of the (small) per-lane counts comes from three ballots, no shuffle and no shared memory
      const uint32_t cnt = __popc(m0) + __popc(m1);
      uint16_t *wq = S->queue;
      uint32_t nW, qoff;
      { const uint32_t lt = (1u << lane) - 1u;
        const uint32_t big = __ballot_sync(0xffffffffu, cnt > 7u);
        if (big == 0)
          { const uint32_t b0 = __ballot_sync(0xffffffffu, cnt & 1u), b1 = __ballot_sync(0xffffffffu, cnt & 2u), b2 = __ballot_sync(0xffffffffu, cnt & 4u);
            qoff = __popc(b0 & lt) + 2u * __popc(b1 & lt) + 4u * __popc(b2 & lt);
            nW = __popc(b0) + 2u * __popc(b1) + 4u * __popc(b2);
          }
        else
          { const uint32_t incl = mg_warp_incl_scan(cnt);
            qoff = incl - cnt;
            nW = __shfl_sync(0xffffffffu, incl, 31);
          }
      }
      if (nW == 0) { __syncwarp(); continue; }
      const bool queued = nW <= C2_WQ_CAP;                   // warp-uniform
      if (queued)
        { uint32_t mm = m1;                                  // from the top bit down: FLO alone finds it
          const uint32_t eb1 = (run0 + 1) << 5, eb0 = run0 << 5;
          uint32_t qa = mg_smem_addr(wq + qoff);             // shared-memory byte address of the lane's next entry
          while (mm)
            { const uint32_t i = c2_top_bit(mm);
              asm volatile("st.shared.u16 [%0], %1;" :: "r"(qa), "h"((uint16_t)(eb1 | i)) : "memory");
              mm ^= 1u << i; qa += 2;
            }
          mm = m0;
          while (mm)
            { const uint32_t i = c2_top_bit(mm);
              asm volatile("st.shared.u16 [%0], %1;" :: "r"(qa), "h"((uint16_t)(eb0 | i)) : "memory");
              mm ^= 1u << i; qa += 2;
            }
        }
      __syncwarp();

      // ---- evaluation and output, round by round (32 queue entries each).  Scatter: a selected k-mer is carried into the
      // next round (also across tiles; the last one is placed after the tile loop), where its bucket atomic is issued
      // before and its store behind that round's evaluation
      // one round: the lanes that `have` an entry evaluate it (ent = run * 32 + window)
      auto round = [&](uint32_t ent, const bool have) {
          if (SCATTER)
            { // the bucket position of the k-mer the previous round selected is requested HERE, at the top of the body, and
              // used below, after this round's evaluation: the L2 round trip of the atomic hides behind the evaluation and
              // no scoreboard is outstanding at the loop's back edge (where ptxas waits for every one: ncu r02, 19 % of all
              // stall samples sat on that wait when the atomic was issued at the end of the body)
              if (pOn) pPos = atomicAdd(&P.cursors[pRegion], 1u);
              asm volatile("" : "+r"(ent) :: "memory");
            }
          uint32_t kl = 0, kh = 0, ok = 0;
          bool isF = false;
          if (have)
            { const uint32_t src = ent >> 5, bit = ent & 31u;
              const uint2 x0 = *reinterpret_cast<const uint2 *>(S->half + 2 * src), x1 = *reinterpret_cast<const uint2 *>(S->half + 2 * src + 2);
              const uint32_t shiftK = SCAN == 1 ? (uint32_t)(64 - 2 * LUTK) : H.shift;     // the table scan knows k at compile time
              ok = mg_eval32_single<SCAN == 1 && P2>(E, shiftK, x0.y, x0.x, x1.y, x1.x, bit, &kl, &kh, &isF) ? 1u : 0u;
            }
          if (SCATTER)
            { asm volatile("" : "+r"(pPos), "+r"(kl), "+r"(kh) :: "memory");   // (the store below stays behind the evaluation)
              if (pOn) c2_place<PEER>(P, pRegion, pPos, ((uint64_t)pKh << 32) | pKl);      // the previous round's k-mer
              pOn = ok;
              if (ok)
                { ++nSelectedLocal;
                  const uint64_t km = ((uint64_t)kh << 32) | kl;
                  pRegion = (uint32_t)((km * 0x9E3779B97F4A7C15ull) >> 32) >> regionShift;            // mg_slot_hash >> regionBits
                  if (PEER) pRegion += mg_owner(km, P.nOwners) * P.nRegions;                           // bucket index = owner * R + region
                  pKl = kl; pKh = kh;
                }
            }
          else
            { // list: one reservation per warp and round
              const uint32_t ballot = __ballot_sync(0xffffffffu, ok);
              if (!ballot) return;
              unsigned long long wbase = 0;
              if (lane == 0) wbase = atomicAdd(P.count, (unsigned long long)__popc(ballot));
              wbase = __shfl_sync(0xffffffffu, wbase, 0);
              if (ok)
                { const uint64_t dst = wbase + __popc(ballot & ((1u << lane) - 1u));
                  if (dst < P.cap)
                    { uint64_t km = ((uint64_t)kh << 32) | kl;
                      if (P.strandBit && isF) km |= 1ull << 63;
                      P.outKmer[dst] = km;
                      if (P.outPos) P.outPos[dst] = (uint32_t)(tileBase + ent);   // ent = run*32 + window
                    }
                }
            }
      };
      if (queued)
        { // the entry is read whether the lane has one or not (the queue is readable to its end: nW <= C2_WQ_CAP rounds up
          // inside it): no divergent region around one load
          for (uint32_t base = 0; base < nW; base += 32) round(wq[base + lane], base + lane < nW);
        }
      else
        { // more candidates than the queue holds (d < 8, pathological sequence): every lane walks its own
          uint32_t own0 = m0, own1 = m1;
          for (;;)
            { const bool have = (own0 | own1) != 0;
              if (!__any_sync(0xffffffffu, have)) break;
              uint32_t ent = 0;
              if (own0) { const uint32_t i = __ffs(own0) - 1; own0 &= own0 - 1; ent = (run0 << 5) | i; }
              else if (own1) { const uint32_t i = __ffs(own1) - 1; own1 &= own1 - 1; ent = ((run0 + 1) << 5) | i; }
              round(ent, have);
            }
        }
      __syncwarp();                                         // the queue and the packed tile are free again
    }
  // the last pending k-mer
  if (SCATTER && pOn)
    { pPos = atomicAdd(&P.cursors[pRegion], 1u);
      c2_place<PEER>(P, pRegion, pPos, ((uint64_t)pKh << 32) | pKl);
    }
  if (SCATTER)
    { nSelectedLocal = mg_warp_sum(nSelectedLocal);
      if (lane == 0 && nSelectedLocal) atomicAdd(P.count, (unsigned long long)nSelectedLocal);
    }
}

// ------------------------------------------------------------------- host
template <int SCAN, int LUTK, int OUT, bool ASCII, bool ENDS, bool P2>
static int c2_launch(const SelectParams &P0, cudaStream_t st)
{
  static int blocksPerSm = 0;
  constexpr int WARPS = C2_WARPS(SCAN);
  const size_t smem = (SCAN ? MG_LUT_SIZE : 0) + WARPS * sizeof(C2WarpSmem<ENDS>);
  if (!blocksPerSm)
    { MG_CUDA(cudaFuncSetAttribute(hash_count2_kernel<SCAN, LUTK, OUT, ASCII, ENDS, P2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      MG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, hash_count2_kernel<SCAN, LUTK, OUT, ASCII, ENDS, P2>, WARPS * 32, smem));
      if (blocksPerSm < 1) blocksPerSm = 1;
    }
  SelectParams P = P0;
  P.nTiles = (uint32_t)((P.nBases + MG_WT_BASES - 1) / MG_WT_BASES);         // warp tiles
  P.E = mg_eval32_prepare(P.H);
  P.regionShift = 32u - (P.slotBits - P.regionBits);
  uint64_t grid = (uint64_t)mg_num_sms() * blocksPerSm;
  const uint64_t need = ((uint64_t)P.nTiles + WARPS * MG_CNT_CHUNK - 1) / (WARPS * MG_CNT_CHUNK);
  if (grid > need) grid = need;
  if (SCAN)
    { lut_build_kernel<<<MG_LUT_SIZE / 256, 256, 0, st>>>(P.H, const_cast<uint8_t *>(P.lut));
      MG_LAUNCH_CHECK("lut_build");
    }
  hash_count2_kernel<SCAN, LUTK, OUT, ASCII, ENDS, P2><<<(unsigned)grid, WARPS * 32, smem, st>>>(P);
  MG_LAUNCH_CHECK("hash_count2");
  return MODGPU_OK;
}

template <int SCAN, int LUTK, int OUT, bool P2>
static int c2_dispatch_io2(const SelectParams &P, bool ends, cudaStream_t st)
{
  if (P.rawAscii) return ends ? c2_launch<SCAN, LUTK, OUT, true, true, P2>(P, st) : c2_launch<SCAN, LUTK, OUT, true, false, P2>(P, st);
  return ends ? c2_launch<SCAN, LUTK, OUT, false, true, P2>(P, st) : c2_launch<SCAN, LUTK, OUT, false, false, P2>(P, st);
}

template <int SCAN, int LUTK, int OUT>
static int c2_dispatch_io(const SelectParams &P, bool ends, cudaStream_t st)
{
  if (SCAN == 1 && P.H.oddInv == 1) return c2_dispatch_io2<SCAN, LUTK, OUT, SCAN == 1>(P, ends, st);     // d a power of two
  return c2_dispatch_io2<SCAN, LUTK, OUT, false>(P, ends, st);
}

template <int OUT>
static int c2_dispatch_scan(const SelectParams &P, int scan, bool ends, cudaStream_t st)
{
  if (scan == 0) return P.H.tz == 0 ? c2_dispatch_io<0, 1, OUT>(P, ends, st) : c2_dispatch_io<0, 0, OUT>(P, ends, st);
  return P.H.k == 31 ? c2_dispatch_io<1, 31, OUT>(P, ends, st) : c2_dispatch_io<1, 30, OUT>(P, ends, st);
}

// ends: per-base flags staged with every tile (many short sequences) instead of the sparse per-tile bytes
int mg_count2_launch(const SelectParams &P, int out, int flags, cudaStream_t st)
{
  static int off = -1;
  if (off < 0) { const char *v = getenv("MODGPU_COUNT_GEN1"); off = (v && atoi(v)) ? 1 : 0; }
  if (off || (flags & MODGPU_SEL_GEN1)) return 1;
  if (!P.raw || (out != 0 && out != 1 && out != 3)) return 1;
  const bool pf = P.H.prefilter && !(flags & MODGPU_SEL_NOPREFILTER);
  int scan;
  if (pf && P.H.lut && !(flags & MODGPU_SEL_NOLUT)) scan = 1;
  else if (!pf && P.H.shift <= 32) scan = 0;
  else return 1;
  if (P.slotBits && (P.slotBits < P.regionBits || P.slotBits - P.regionBits > 31)) return 1;
  const bool ends = P.tileFlags == nullptr;
  if (out == 0) return c2_dispatch_scan<0>(P, scan, ends, st);
  if (out == 1) return c2_dispatch_scan<1>(P, scan, ends, st);
  return c2_dispatch_scan<3>(P, scan, ends, st);
}
