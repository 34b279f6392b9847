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
//   - the queue is filled through ONE shared-memory atomic per lane (order is irrelevant in count mode) instead of
//     a warp scan, and drained from the top bit (FLO alone, no BREV);
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

#define C2_ROUNDS 4                                             // queue entries a lane keeps in registers

template <bool ENDS> struct C2WarpSmem {
  __align__(16) uint8_t stage[MG_WS_RAW_BYTES];                // the raw tile + 32 bytes of overlap (TMA destination)
  __align__(16) uint32_t ends[ENDS ? MG_WS_ENDS_BYTES / 4 : 4];
  __align__(16) uint32_t half[2 * (MG_WT_RUNS + 2)];           // the packed tile as 32-bit halves (word w = half[2w+1] : half[2w])
  __align__(8) uint64_t bar;
  uint32_t qn[2];                                              // queue fill of this tile / the next one
  uint16_t queue[MG_WQ_CAP];
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

// SCAN: 0 = every window evaluated in full (32-bit pieces, k >= 16), 1 = table-driven candidates (LUTK = k)
// OUT:  0 = list, 1 = the table's region buckets, 3 = per-(owner, region) buckets
template <int SCAN, int LUTK, int OUT, bool ASCII, bool ENDS>
__global__ void __launch_bounds__(MG_CNT_THREADS, SCAN ? 3 : 2) hash_count2_kernel(const SelectParams P)
{
  constexpr bool SCATTER = (OUT == 1 || OUT == 3);
  constexpr bool PEER = (OUT == 3);
  extern __shared__ __align__(128) uint8_t sDyn[];
  uint8_t *sLut = sDyn;                                                            // MG_LUT_SIZE bytes when SCAN
  C2WarpSmem<ENDS> *S = reinterpret_cast<C2WarpSmem<ENDS> *>(sDyn + (SCAN ? MG_LUT_SIZE : 0)) + (threadIdx.x >> 5);

  const MgKHasher &H = P.H;
  const uint32_t tid = threadIdx.x, lane = tid & 31;
  const MgEval32 E = mg_eval32_prepare(H);
  const bool pow2 = H.oddInv == 1;                         // kernel-uniform: no odd-part test
  // region of a k-mer = the top (slotBits - regionBits) bits of its slot hash = one shift of the high product word
  const uint32_t regionShift = 32u - (P.slotBits - P.regionBits);
  // tile schedule (as in the first generation): chunks of MG_CNT_CHUNK consecutive warp tiles, the first by warp index,
  // the following ones from an atomic ticket requested a whole chunk ahead
  const uint32_t nWarps = gridDim.x * MG_CNT_WARPS;
  uint64_t tile = (uint64_t)(blockIdx.x * MG_CNT_WARPS + (tid >> 5)) * MG_CNT_CHUNK, tileNext = 0;
  uint32_t pendingChunk = 0;
  if (lane == 0) pendingChunk = nWarps + atomicAdd(P.ticket, 1u);
  // a tile can be bulk-copied when all of it and its overlap lie inside the batch (the raw bytes carry no slack)
  const uint64_t nBulk = P.nBases >= MG_WS_RAW_BYTES ? (P.nBases - MG_WS_RAW_BYTES) / MG_WT_BASES + 1 : 0;
  uint32_t nSelectedLocal = 0;
  uint32_t phase = 0, qp = 0;

  if (lane == 0)
    { mg_mbar_init(&S->bar, 1);
      mg_fence_barrier_init();
      mg_fence_proxy_async();
      S->qn[0] = 0; S->qn[1] = 0;
      if (tile < nBulk) c2_issue_tile<ENDS>(P, S, tile);
    }
  if (SCAN)
    { const uint4 *src = reinterpret_cast<const uint4 *>(P.lut);
      uint4 *dst = reinterpret_cast<uint4 *>(sLut);
      for (uint32_t i = tid; i < MG_LUT_SIZE / 16; i += MG_CNT_THREADS) dst[i] = __ldg(src + i);
    }
  __syncthreads();

  for (; tile < P.nTiles; tile = tileNext, qp ^= 1)
    { const uint32_t run0 = lane * 2;
      const uint64_t tileBase = tile * MG_WT_BASES;
      uint32_t tflag = 0;
      if (!ENDS) tflag = __ldg(P.tileFlags + tile);         // warp-uniform; consumed after the scan
      uint32_t e0 = 0, e1 = 0, e2 = 0;
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
          if (ENDS) { e0 = S->ends[run0]; e1 = S->ends[run0 + 1]; e2 = S->ends[run0 + 2]; }
        }
      else
        { // the ragged end of the batch: guarded loads
          const uint64_t word = tile * MG_WT_RUNS + run0;
          W[run0] = pack32_raw<ASCII>(P.raw, word * MG_RUN, P.nBases);
          W[run0 + 1] = pack32_raw<ASCII>(P.raw, word * MG_RUN + 32, P.nBases);
          if (lane == 31) W[MG_WT_RUNS] = pack32_raw<ASCII>(P.raw, word * MG_RUN + 64, P.nBases);
          e0 = __ldg(P.ends + word); e1 = __ldg(P.ends + word + 1); e2 = __ldg(P.ends + word + 2);
          tflag = 1;
        }
      __syncwarp();
      // every lane has consumed the staged tile: the next tile's copy streams in behind the computation of this one
      tileNext = tile + 1;
      if ((tileNext & (MG_CNT_CHUNK - 1)) == 0)
        { tileNext = (uint64_t)__shfl_sync(0xffffffffu, pendingChunk, 0) * MG_CNT_CHUNK;
          if (lane == 0 && tileNext < P.nTiles) pendingChunk = nWarps + atomicAdd(P.ticket, 1u);
        }
      if (lane == 0)
        { S->qn[qp ^ 1] = 0;                                 // the next tile's queue counter (last used two tiles ago)
          if (tileNext < nBulk) c2_issue_tile<ENDS>(P, S, tileNext);
        }
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
              if (H.tz == 0)
                {
#pragma unroll
                  for (int i = 0; i < MG_RUN; ++i)
                    { const bool ok = mg_selected32<true>(E, Q, i);
                      asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p or.b32 %0, %0, %2;\n\t}" : "+r"(mh) : "r"((uint32_t)ok), "r"(1u << i));
                    }
                }
              else
                {
#pragma unroll
                  for (int i = 0; i < MG_RUN; ++i)
                    { const bool ok = mg_selected32<false>(E, Q, i);
                      asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p or.b32 %0, %0, %2;\n\t}" : "+r"(mh) : "r"((uint32_t)ok), "r"(1u << i));
                    }
                }
              if (h == 0) m0 = mh; else m1 = mh;
              wa = wb; wb = wc;
            }
        }
      // windows that would span two sequences or run off the batch are not usable
      { bool plain = tileBase + MG_WT_BASES + MG_RUN <= P.nBases;
        if (ENDS) plain = plain && !__any_sync(0xffffffffu, (e0 | e1 | e2) != 0u);
        else plain = plain && !tflag;
        if (!plain)
          { if (!ENDS && tile < nBulk)
              { const uint64_t word = tile * MG_WT_RUNS + run0;
                e0 = __ldg(P.ends + word); e1 = __ldg(P.ends + word + 1); e2 = __ldg(P.ends + word + 2);
              }
            const uint64_t p0 = tileBase + (uint64_t)run0 * MG_RUN;
            m0 &= mg_run_usable((uint64_t)e0 | ((uint64_t)e1 << 32), H.k, p0, P.nBases);
            m1 &= mg_run_usable((uint64_t)e1 | ((uint64_t)e2 << 32), H.k, p0 + MG_RUN, P.nBases);
          }
      }

      // ---- the warp's queue of (run, window): one shared-memory atomic per lane reserves its entries
      const uint32_t cnt = __popc(m0) + __popc(m1);
      uint16_t *wq = S->queue;
      { uint32_t qoff = 0;
        if (cnt) qoff = atomicAdd(&S->qn[qp], cnt);
        if (cnt && qoff + cnt <= MG_WQ_CAP)
          { uint32_t mm = m1;                                // from the top bit down: FLO alone finds it
            const uint32_t eb1 = (run0 + 1) << 5, eb0 = run0 << 5;
            while (mm) { const uint32_t i = c2_top_bit(mm); mm ^= 1u << i; wq[qoff++] = (uint16_t)(eb1 | i); }
            mm = m0;
            while (mm) { const uint32_t i = c2_top_bit(mm); mm ^= 1u << i; wq[qoff++] = (uint16_t)(eb0 | i); }
          }
      }
      __syncwarp();
      const uint32_t nW = S->qn[qp];
      if (nW == 0) { __syncwarp(); continue; }
      const bool queued = nW <= MG_WQ_CAP;                   // warp-uniform

      // ---- evaluation and output
      if (queued && nW <= C2_ROUNDS * 32)
        { // the usual case: every lane evaluates its (<= C2_ROUNDS) queue entries into registers first
          uint32_t kl[C2_ROUNDS], kh[C2_ROUNDS], ent[C2_ROUNDS];
          uint32_t okMask = 0, fMask = 0;
#pragma unroll
          for (int r = 0; r < C2_ROUNDS; ++r)
            { kl[r] = 0; kh[r] = 0; ent[r] = 0;
              if (r * 32 < nW)                               // warp-uniform
                { const uint32_t q = r * 32 + lane;
                  if (q < nW)
                    { ent[r] = wq[q];
                      const uint32_t src = ent[r] >> 5, bit = ent[r] & 31u;
                      const uint2 x0 = *reinterpret_cast<const uint2 *>(S->half + 2 * src), x1 = *reinterpret_cast<const uint2 *>(S->half + 2 * src + 2);
                      bool isF, ok;
                      if (pow2) ok = mg_eval32_single<true>(E, H.shift, x0.y, x0.x, x1.y, x1.x, bit, &kl[r], &kh[r], &isF);
                      else ok = mg_eval32_single<false>(E, H.shift, x0.y, x0.x, x1.y, x1.x, bit, &kl[r], &kh[r], &isF);
                      if (ok) { okMask |= 1u << r; if (isF) fMask |= 1u << r; }
                    }
                }
            }
          if (SCATTER)
            { nSelectedLocal += __popc(okMask);
              uint32_t pos[C2_ROUNDS], region[C2_ROUNDS];
#pragma unroll
              for (int r = 0; r < C2_ROUNDS; ++r)
                if ((okMask >> r) & 1u)
                  { const uint64_t km = ((uint64_t)kh[r] << 32) | kl[r];
                    region[r] = (uint32_t)((km * 0x9E3779B97F4A7C15ull) >> 32) >> regionShift;       // mg_slot_hash >> regionBits
                    if (PEER) region[r] += mg_owner(km, P.nOwners) * P.nRegions;                      // bucket index = owner * R + region
                    pos[r] = atomicAdd(&P.cursors[region[r]], 1u);
                  }
#pragma unroll
              for (int r = 0; r < C2_ROUNDS; ++r)
                if ((okMask >> r) & 1u)
                  { const uint64_t km = ((uint64_t)kh[r] << 32) | kl[r];
                    if (pos[r] < P.bucketCap)
                      { uint64_t *bp = P.buckets + (uint64_t)region[r] * P.bucketCap + pos[r];
                        if (P.keepBuckets) mg_st_keep(bp, km); else *bp = km;
                      }
                    else if (PEER)
                      { const uint32_t ow = region[r] / P.nRegions;
                        const uint32_t o = atomicAdd(&P.ownerCursor[ow], 1u);
                        if (o < P.overflowCap) P.overflow[(uint64_t)ow * P.overflowCap + o] = km;
                      }
                    else
                      { const uint32_t o = atomicAdd(&P.cursors[P.nRegions], 1u);
                        if (o < P.overflowCap) P.overflow[o] = km;
                      }
                  }
            }
          else
            { // list: one reservation per warp and tile
              const uint32_t cs = __popc(okMask);
              const uint32_t inc2 = mg_warp_incl_scan(cs);
              const uint32_t total = __shfl_sync(0xffffffffu, inc2, 31);
              unsigned long long wbase = 0;
              if (lane == 0 && total) wbase = atomicAdd(P.count, (unsigned long long)total);
              wbase = __shfl_sync(0xffffffffu, wbase, 0);
              uint64_t dst = wbase + inc2 - cs;
#pragma unroll
              for (int r = 0; r < C2_ROUNDS; ++r)
                if ((okMask >> r) & 1u)
                  { if (dst < P.cap)
                      { const uint64_t km = ((uint64_t)kh[r] << 32) | kl[r];
                        P.outKmer[dst] = (P.strandBit && ((fMask >> r) & 1u)) ? (km | (1ull << 63)) : km;
                        if (P.outPos) P.outPos[dst] = (uint32_t)(tileBase + ent[r]);   // ent = run*32 + window
                      }
                    ++dst;
                  }
            }
        }
      else
        { // crowded warp: warp-aggregated reservations per round; per-lane loop when even the queue overflowed
          uint32_t own0 = m0, own1 = m1;
          for (uint32_t base = 0;; base += 32)
            { uint32_t ent = 0;
              bool have;
              if (queued)
                { if (base >= nW) break;
                  have = base + lane < nW;
                  if (have) ent = wq[base + lane];
                }
              else
                { have = (own0 | own1) != 0;
                  if (!__any_sync(0xffffffffu, have)) break;
                  if (own0) { const uint32_t i = __ffs(own0) - 1; own0 &= own0 - 1; ent = (run0 << 5) | i; }
                  else if (own1) { const uint32_t i = __ffs(own1) - 1; own1 &= own1 - 1; ent = ((run0 + 1) << 5) | i; }
                }
              uint32_t kl = 0, kh = 0;
              bool isF = false, ok = false;
              if (have)
                { const uint32_t src = ent >> 5, bit = ent & 31u;
                  const uint2 x0 = *reinterpret_cast<const uint2 *>(S->half + 2 * src), x1 = *reinterpret_cast<const uint2 *>(S->half + 2 * src + 2);
                  ok = mg_eval32_single<false>(E, H.shift, x0.y, x0.x, x1.y, x1.x, bit, &kl, &kh, &isF);
                }
              uint64_t km = ((uint64_t)kh << 32) | kl;
              const uint32_t ballot = __ballot_sync(0xffffffffu, ok);
              if (!ballot) continue;
              if (SCATTER)
                { if (ok)
                    { ++nSelectedLocal;
                      uint32_t region = (uint32_t)((km * 0x9E3779B97F4A7C15ull) >> 32) >> regionShift;
                      const uint32_t ow = PEER ? mg_owner(km, P.nOwners) : 0u;
                      if (PEER) region += ow * P.nRegions;
                      const uint32_t pos = atomicAdd(&P.cursors[region], 1u);
                      if (pos < P.bucketCap) { uint64_t *bp = P.buckets + (uint64_t)region * P.bucketCap + pos; if (P.keepBuckets) mg_st_keep(bp, km); else *bp = km; }
                      else if (PEER)
                        { const uint32_t o = atomicAdd(&P.ownerCursor[ow], 1u);
                          if (o < P.overflowCap) P.overflow[(uint64_t)ow * P.overflowCap + o] = km;
                        }
                      else
                        { const uint32_t o = atomicAdd(&P.cursors[P.nRegions], 1u);
                          if (o < P.overflowCap) P.overflow[o] = km;
                        }
                    }
                  continue;
                }
              unsigned long long wbase = 0;
              if (lane == 0) wbase = atomicAdd(P.count, (unsigned long long)__popc(ballot));
              wbase = __shfl_sync(0xffffffffu, wbase, 0);
              if (ok)
                { const uint64_t dst = wbase + __popc(ballot & ((1u << lane) - 1u));
                  if (dst < P.cap)
                    { if (P.strandBit && isF) km |= 1ull << 63;
                      P.outKmer[dst] = km;
                      if (P.outPos) P.outPos[dst] = (uint32_t)(tileBase + ent);
                    }
                }
            }
        }
      __syncwarp();                                         // the queue and the packed tile are free again
    }
  if (SCATTER)
    { nSelectedLocal = mg_warp_sum(nSelectedLocal);
      if (lane == 0 && nSelectedLocal) atomicAdd(P.count, (unsigned long long)nSelectedLocal);
    }
}

// ------------------------------------------------------------------- host
template <int SCAN, int LUTK, int OUT, bool ASCII, bool ENDS>
static int c2_launch(const SelectParams &P0, cudaStream_t st)
{
  static int blocksPerSm = 0;
  const size_t smem = (SCAN ? MG_LUT_SIZE : 0) + MG_CNT_WARPS * sizeof(C2WarpSmem<ENDS>);
  if (!blocksPerSm)
    { MG_CUDA(cudaFuncSetAttribute(hash_count2_kernel<SCAN, LUTK, OUT, ASCII, ENDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      MG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, hash_count2_kernel<SCAN, LUTK, OUT, ASCII, ENDS>, MG_CNT_THREADS, smem));
      if (blocksPerSm < 1) blocksPerSm = 1;
    }
  SelectParams P = P0;
  P.nTiles = (uint32_t)((P.nBases + MG_WT_BASES - 1) / MG_WT_BASES);         // warp tiles
  uint64_t grid = (uint64_t)mg_num_sms() * blocksPerSm;
  const uint64_t need = ((uint64_t)P.nTiles + MG_CNT_WARPS * MG_CNT_CHUNK - 1) / (MG_CNT_WARPS * MG_CNT_CHUNK);
  if (grid > need) grid = need;
  if (SCAN)
    { lut_build_kernel<<<MG_LUT_SIZE / 256, 256, 0, st>>>(P.H, const_cast<uint8_t *>(P.lut));
      MG_LAUNCH_CHECK("lut_build");
    }
  hash_count2_kernel<SCAN, LUTK, OUT, ASCII, ENDS><<<(unsigned)grid, MG_CNT_THREADS, smem, st>>>(P);
  MG_LAUNCH_CHECK("hash_count2");
  return MODGPU_OK;
}

template <int SCAN, int LUTK, int OUT>
static int c2_dispatch_io(const SelectParams &P, bool ends, cudaStream_t st)
{
  if (P.rawAscii) return ends ? c2_launch<SCAN, LUTK, OUT, true, true>(P, st) : c2_launch<SCAN, LUTK, OUT, true, false>(P, st);
  return ends ? c2_launch<SCAN, LUTK, OUT, false, true>(P, st) : c2_launch<SCAN, LUTK, OUT, false, false>(P, st);
}

template <int OUT>
static int c2_dispatch_scan(const SelectParams &P, int scan, bool ends, cudaStream_t st)
{
  if (scan == 0) return c2_dispatch_io<0, 0, OUT>(P, ends, st);
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
