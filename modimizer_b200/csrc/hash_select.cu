// hash_select.cu - K2: canonical k-mer hashing + hash % d == 0 selection.
//
// replaces: modRCiterator / modRCnext / advanceHashRC / hashRC of the
// reference (seqhash.c:60-79,154-196) for a whole batch of sequences at once.
//
// What the reference's serial "rolling" iterator computes per window start p is
// a pure function of the 2k-bit window, so positions are
// independent.  Layout of the work:
//   - a thread owns a RUN of 32 consecutive window starts = one packed word
//     plus the next (k-1 <= 30 overlap bases); forward and reverse-complement
//     k-mers of window i are constant-shift bit fields of two 128-bit registers
//     (mg_run_fwd / mg_run_rc), no per-base extraction, no serial dependence;
//   - a 256-thread tile = 8192 bases = 2 KiB packed + 1 KiB end flags, staged
//     in shared memory by one TMA bulk copy per stream (cp.async.bulk +
//     mbarrier, double buffered), so the next tile streams in while this one
//     is hashed;
//   - tiles are claimed through an atomic ticket; the selected list is written
//     in input order with a decoupled look-back scan (ORDERED) or at an
//     atomically reserved offset (count mode, order irrelevant);
//   - when d = 2^t * odd with t >= 3 and 64-2k+t <= 32 (e.g. k=31 d=64, the
//     modmap index configuration) a PREFILTER evaluates only the low product
//     word of both strands (2 IMAD + 2 compares per base) and the full 64-bit
//     canonical test runs on the ~2/2^t surviving candidates.
//
// Integer-issue bound, not HBM bound (0.25 + 8/d algorithmic bytes per base
// against ~9-28 instructions per base) - see DESIGN.md for the roofline.
#include <string.h>
#include <stdlib.h>
#include "mg_device.cuh"
#include "mg_select.cuh"

#define MG_TILE_PACK_BYTES (MG_TILE_THREADS * 8 + 16)     // 256 words + overlap word, 16 B multiple (258 words)
#define MG_TILE_ENDS_BYTES (MG_TILE_THREADS * 4 + 16)

// Phase 2 helper: evaluate queue entry e = (source thread << 5 | window) of the
// current tile from the shared-memory copy of the packed words.
__device__ __forceinline__ bool eval_entry(const MgKHasher &H, const uint64_t *sWords, uint32_t e,
                                           uint64_t *km, bool *isF)
{
  const uint32_t src = e >> 5, bit = e & 31u;
  return mg_eval_single(H, sWords[src], sWords[src + 1], bit, km, isF);
}

// LUTK != 0: k = LUTK at compile time; when d is also a power of two (kernel-uniform) the specialised
// evaluation applies (constant shifts, masked-product comparison, no odd-part test)
template <int LUTK>
__device__ __forceinline__ bool eval_entry_k(const MgKHasher &H, const uint64_t *sWords, uint32_t e,
                                             uint64_t *km, bool *isF)
{
  const uint32_t src = e >> 5, bit = e & 31u;
  if (LUTK != 0 && H.oddInv == 1) return mg_eval_single_pow2<LUTK ? LUTK : 31>(H, sWords[src], sWords[src + 1], bit, km, isF);
  return mg_eval_single(H, sWords[src], sWords[src + 1], bit, km, isF);
}

// One tile (256 runs = 8192 window starts) in three phases, by 128 threads
// owning two consecutive runs each (halves the per-run bookkeeping):
//  1. every thread scans its runs into bit masks: with the PREFILTER the
//     candidates (cheap low-word test, 7 instructions per window), otherwise
//     the windows that are selected (full canonical test);
//  2. the set bits of the whole tile are compacted into a shared-memory queue
//     in position order (one block scan), so that
//  3. the expensive part - full 64-bit evaluation of a candidate, extraction
//     of the winning strand's k-mer, the store - is spread evenly over the
//     threads whatever the distribution of hits among the runs (a per-thread
//     loop over its own hits costs max-over-lanes iterations per warp: measured
//     2/3 of all issued instructions in the first version of this kernel).
//     Tiles with more hits than the queue holds (d < 4, pathological
//     sequence) take the per-thread loop instead.
#define MG_SEL_THREADS 128
#define MG_SEL_RPT (MG_TILE_THREADS / MG_SEL_THREADS)          // runs per thread = 2
#define MG_QUEUE_CAP 4096
#define MG_SEL_ROUNDS 4                                          // queue entries a thread keeps in registers

template <bool PREFILTER>
__device__ __forceinline__ uint32_t scan_run(const MgKHasher &H, const MgRun &R)
{
  uint32_t m = 0;
  if (PREFILTER)
    {
#pragma unroll
      for (int i = 0; i < MG_RUN; ++i)
        { // candidate <=> min(low product word of fwd, of rc) < pfLim; one predicated OR per window
          const uint32_t pf = mg_run_fwd_lo(R, i) * H.pfMul, pr = mg_run_rc_lo(R, i) * H.pfMul;
          const uint32_t mn = min(pf, pr);
          asm("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}"
              : "+r"(m) : "r"(mn), "r"(H.pfLim), "r"(1u << i));
        }
    }
  else if (H.shift <= 32)                              // k >= 16 (kernel-uniform): the 32-bit evaluation of mg_common.cuh
    { const MgEval32 E = mg_eval32_prepare(H);
      const MgRun32 Q = mg_run32(R);
      if (H.tz == 0)
        {
#pragma unroll
          for (int i = 0; i < MG_RUN; ++i)
            { const bool ok = mg_selected32<true>(E, Q, i);
              asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p or.b32 %0, %0, %2;\n\t}" : "+r"(m) : "r"((uint32_t)ok), "r"(1u << i));
            }
        }
      else
        {
#pragma unroll
          for (int i = 0; i < MG_RUN; ++i)
            { const bool ok = mg_selected32<false>(E, Q, i);
              asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p or.b32 %0, %0, %2;\n\t}" : "+r"(m) : "r"((uint32_t)ok), "r"(1u << i));
            }
        }
    }
  else
    {
#pragma unroll
      for (int i = 0; i < MG_RUN; ++i)
        { uint64_t km; bool isF;
          if (mg_eval_window(H, R, i, &km, &isF)) m |= 1u << i;
        }
    }
  return m;
}

// ---------------------------------------------------------------- ordered --
// Input-order output (modmap hit lists, exact index numbering): block-level queue,
// a scan over the runs and a decoupled look-back over the tiles place the k-mers.
// TMA: the packed stream and the end flags are staged by bulk copies (double buffered).
template <bool PREFILTER, bool TMA>
__global__ void __launch_bounds__(MG_SEL_THREADS) hash_select_ordered_kernel(const SelectParams P)
{
  constexpr int NBUF = TMA ? 2 : 1;
  constexpr int NWARPS = MG_SEL_THREADS / 32;
  __shared__ __align__(128) uint64_t sPack[NBUF][MG_TILE_PACK_BYTES / 8];
  __shared__ __align__(128) uint32_t sEnds[TMA ? 2 : 1][TMA ? (MG_TILE_ENDS_BYTES / 4) : 4];
  __shared__ __align__(8) uint64_t sBar[2];
  __shared__ uint16_t sQueue[MG_QUEUE_CAP];
  __shared__ uint32_t sSel[MG_TILE_THREADS];      // selected windows per run
  __shared__ uint32_t sDst[MG_TILE_THREADS];      // output offset of each run
  __shared__ uint32_t sTile[2];
  __shared__ uint32_t sWarp[NWARPS];
  __shared__ uint64_t sBase;

  const MgKHasher &H = P.H;
  const uint32_t tid = threadIdx.x;

  if (tid == 0)
    { if (TMA)
        { mg_mbar_init(&sBar[0], 1);
          mg_mbar_init(&sBar[1], 1);
          mg_fence_barrier_init();
          mg_fence_proxy_async();
        }
      uint32_t t0 = atomicAdd(P.ticket, 1u);
      sTile[0] = t0;
      if (TMA && t0 < P.nTiles)
        { mg_mbar_expect_tx(&sBar[0], MG_TILE_PACK_BYTES + MG_TILE_ENDS_BYTES);
          mg_tma_load_1d(sPack[0], P.packed + (uint64_t)t0 * MG_TILE_THREADS, MG_TILE_PACK_BYTES, &sBar[0]);
          mg_tma_load_1d(sEnds[0], P.ends + (uint64_t)t0 * MG_TILE_THREADS, MG_TILE_ENDS_BYTES, &sBar[0]);
        }
    }

  for (uint32_t it = 0;; ++it)
    { const uint32_t stage = it & 1;
      const uint32_t buf = TMA ? stage : 0;
      __syncthreads();               // sTile[stage] visible; everyone is done with the other buffer and the queue
      const uint32_t tile = sTile[stage];
      if (tile >= P.nTiles) break;

      if (tid == 0)
        { uint32_t tn = atomicAdd(P.ticket, 1u);
          sTile[stage ^ 1] = tn;
          if (TMA && tn < P.nTiles)
            { mg_mbar_expect_tx(&sBar[stage ^ 1], MG_TILE_PACK_BYTES + MG_TILE_ENDS_BYTES);
              mg_tma_load_1d(sPack[stage ^ 1], P.packed + (uint64_t)tn * MG_TILE_THREADS, MG_TILE_PACK_BYTES, &sBar[stage ^ 1]);
              mg_tma_load_1d(sEnds[stage ^ 1], P.ends + (uint64_t)tn * MG_TILE_THREADS, MG_TILE_ENDS_BYTES, &sBar[stage ^ 1]);
            }
        }

      // ---- this thread's two runs: words 2t, 2t+1 (+ overlap word 2t+2), 96 end flags
      const uint32_t run0 = tid * MG_SEL_RPT;
      const uint64_t word = (uint64_t)tile * MG_TILE_THREADS + run0;
      uint64_t w0, w1, w2;
      uint32_t e0, e1, e2;
      if (TMA)
        { mg_mbar_wait(&sBar[stage], (it >> 1) & 1);
          w0 = sPack[buf][run0]; w1 = sPack[buf][run0 + 1]; w2 = sPack[buf][run0 + 2];
          e0 = sEnds[buf][run0]; e1 = sEnds[buf][run0 + 1]; e2 = sEnds[buf][run0 + 2];
        }
      else
        { w0 = __ldg(P.packed + word); w1 = __ldg(P.packed + word + 1); w2 = __ldg(P.packed + word + 2);
          e0 = __ldg(P.ends + word); e1 = __ldg(P.ends + word + 1); e2 = __ldg(P.ends + word + 2);
          sPack[0][run0] = w0;                             // phase 3 reads the tile from shared memory
          sPack[0][run0 + 1] = w1;
          if (tid == MG_SEL_THREADS - 1) sPack[0][MG_TILE_THREADS] = w2;
        }
      const uint64_t tileBase = (uint64_t)tile * MG_TILE_BASES;
      uint32_t m[MG_SEL_RPT];
      { const uint64_t p0 = tileBase + (uint64_t)run0 * MG_RUN;
        const MgRun RA = mg_run_prepare(w0, w1, H.k);
        m[0] = scan_run<PREFILTER>(H, RA) & mg_run_usable((uint64_t)e0 | ((uint64_t)e1 << 32), H.k, p0, P.nBases);
        const MgRun RB = mg_run_prepare(w1, w2, H.k);
        m[1] = scan_run<PREFILTER>(H, RB) & mg_run_usable((uint64_t)e1 | ((uint64_t)e2 << 32), H.k, p0 + MG_RUN, P.nBases);
      }

      // ---- phase 2: queue of (run, window) in position order
      uint32_t nQueue;
      uint32_t qoff = mg_block_excl_scan<NWARPS>(__popc(m[0]) + __popc(m[1]), sWarp, &nQueue);
      const bool queued = nQueue <= MG_QUEUE_CAP;          // block-uniform
      sSel[run0] = 0; sSel[run0 + 1] = 0;
      if (queued)
        {
#pragma unroll
          for (int r = 0; r < MG_SEL_RPT; ++r)
            { uint32_t mm = m[r];
              while (mm)
                { uint32_t i = __ffs(mm) - 1; mm &= mm - 1;
                  sQueue[qoff++] = (uint16_t)(((run0 + r) << 5) | i);
                }
            }
        }
      __syncthreads();

      // ---- phase 3: pass 1 marks the selected windows of every run ...
      const uint64_t *sWords = sPack[buf];
      if (!queued)
        { // overfull tile: every thread resolves its own windows
#pragma unroll
          for (int r = 0; r < MG_SEL_RPT; ++r)
            { uint32_t mm = m[r], sel = 0;
              while (mm)
                { uint32_t i = __ffs(mm) - 1; mm &= mm - 1;
                  uint64_t km; bool isF;
                  if (!PREFILTER || eval_entry(H, sWords, ((run0 + r) << 5) | i, &km, &isF)) sel |= 1u << i;
                }
              sSel[run0 + r] = sel;
            }
        }
      else if (PREFILTER)
        { for (uint32_t q = tid; q < nQueue; q += MG_SEL_THREADS)
            { const uint32_t e = sQueue[q];
              uint64_t km; bool isF;
              if (eval_entry(H, sWords, e, &km, &isF)) atomicOr(&sSel[e >> 5], 1u << (e & 31u));
            }
        }
      else
        { for (uint32_t q = tid; q < nQueue; q += MG_SEL_THREADS)
            { const uint32_t e = sQueue[q];
              atomicOr(&sSel[e >> 5], 1u << (e & 31u));          // the queue already holds the selected windows
            }
        }
      __syncthreads();
      // ... a scan over the runs + the look-back over the tiles place them ...
      uint32_t total;
      const uint32_t selA = sSel[run0], selB = sSel[run0 + 1];
      const uint32_t off = mg_block_excl_scan<NWARPS>(__popc(selA) + __popc(selB), sWarp, &total);
      sDst[run0] = off;
      sDst[run0 + 1] = off + __popc(selA);
      if (tid < 32)
        { uint32_t excl = mg_lookback(P.status, tile, total);
          if (tid == 0)
            { sBase = excl;
              if (tile == P.nTiles - 1) *P.count = (unsigned long long)excl + total;
            }
        }
      __syncthreads();
      const uint64_t outBase = sBase;
      // ... pass 2 writes them, again spread over all threads (or per owner when overfull)
      uint32_t own0 = selA, own1 = selB;
      for (uint32_t q = tid;; q += MG_SEL_THREADS)
        { uint32_t e;
          if (queued)
            { if (q >= nQueue) break;
              e = sQueue[q];
            }
          else
            { if (own0) { uint32_t i = __ffs(own0) - 1; own0 &= own0 - 1; e = (run0 << 5) | i; }
              else if (own1) { uint32_t i = __ffs(own1) - 1; own1 &= own1 - 1; e = ((run0 + 1) << 5) | i; }
              else break;
            }
          const uint32_t src = e >> 5, bit = e & 31u;
          const uint32_t selBits = sSel[src];
          if (!((selBits >> bit) & 1u)) continue;
          uint64_t km; bool isF;
          eval_entry(H, sWords, e, &km, &isF);
          const uint64_t dst = outBase + sDst[src] + __popc(selBits & ((1u << bit) - 1u));
          if (dst < P.cap)
            { if (P.strandBit && isF) km |= 1ull << 63;
              P.outKmer[dst] = km;
              if (P.outPos) P.outPos[dst] = (uint32_t)(tileBase + e);
            }
        }
    }
}

// ------------------------------------------------------------------ count --
// Count mode (order irrelevant): every WARP is autonomous.  A warp owns tiles of
// 64 runs (2048 window starts), t = global warp, + total warps, ...
//   - the tile (2 KiB of raw bytes, or 512 B of packed words, plus its end flags)
//     is staged in the warp's own shared-memory buffer by TMA bulk copies
//     (cp.async.bulk + a per-warp mbarrier; SASS UBLKCP); the copy of the NEXT tile
//     is issued as soon as the lanes have pulled the current one into registers,
//     so it streams in behind the whole computation of this tile;
//   - K1 fused: a lane packs its 64 bytes into two words (pack16_dev);
//   - candidates: table-driven (LUTK, one shared-memory lookup per 4 positions),
//     multiplicative low-word prefilter, or full evaluation of every window;
//   - the candidates of the warp's 2048 windows are compacted into the warp's queue
//     (warp scan), evaluated evenly by its lanes, kept in registers (MG_SEL_ROUNDS
//     per lane) so that the atomics of a lane are all in flight together, and written.
// No block barrier after the prologue; the only block-shared state is the read-only
// candidate table.
// OUT: 0 = list, 1 = scatter into the table's region buckets, 2 = per-owner segments,
//      3 = per-(owner, region) buckets: what the owner's region build consumes directly
// RAW: the batch as bytes (codes or ASCII, 16-byte aligned) instead of the packed stream
template <bool RAW> struct CountWarpSmem {
  __align__(16) uint8_t stage[RAW ? MG_WS_RAW_BYTES : MG_WS_PACK_BYTES];
  __align__(16) uint32_t ends[MG_WS_ENDS_BYTES / 4];
  __align__(16) uint64_t words[MG_WT_RUNS + 2];                // the packed tile, read by phase 3
  __align__(8) uint64_t bar;
  uint16_t queue[MG_WQ_CAP];
  uint32_t own[64];                                            // OUT == 2: per-owner counts, then bases
};


template <bool RAW>
__device__ __forceinline__ void count_issue_tile(const SelectParams &P, CountWarpSmem<RAW> *S, uint64_t tile)
{ // one elected lane: both bulk copies of a tile signal the warp's mbarrier
  mg_mbar_expect_tx(&S->bar, (RAW ? MG_WS_RAW_BYTES : MG_WS_PACK_BYTES) + MG_WS_ENDS_BYTES);
  if (RAW) mg_tma_load_1d_hint(S->stage, P.raw + tile * MG_WT_BASES, MG_WS_RAW_BYTES, &S->bar, MG_L2_EVICT_FIRST);
  else mg_tma_load_1d_hint(S->stage, P.packed + tile * MG_WT_RUNS, MG_WS_PACK_BYTES, &S->bar, MG_L2_EVICT_FIRST);
  mg_tma_load_1d_hint(S->ends, P.ends + tile * MG_WT_RUNS, MG_WS_ENDS_BYTES, &S->bar, MG_L2_EVICT_FIRST);
}

template <bool PREFILTER, bool RAW, int OUT, bool ASCII, int LUTK>
__global__ void __launch_bounds__(MG_CNT_THREADS, LUTK ? 3 : 2) hash_count_kernel(const SelectParams P)
{
  static_assert(LUTK == 0 || PREFILTER, "the table-driven scan is a prefilter");
  constexpr bool SCATTER = (OUT == 1 || OUT == 3);
  constexpr bool OWNERS = (OUT == 2);
  constexpr bool PEER = (OUT == 3);
  extern __shared__ __align__(128) uint8_t sDyn[];
  uint8_t *sLut = sDyn;                                                           // MG_LUT_SIZE bytes when LUTK
  CountWarpSmem<RAW> *S = reinterpret_cast<CountWarpSmem<RAW> *>(sDyn + (LUTK ? MG_LUT_SIZE : 0)) + (threadIdx.x >> 5);

  const MgKHasher &H = P.H;
  const uint32_t tid = threadIdx.x;
  const uint32_t lane = tid & 31;
  // tile schedule: chunks of MG_CNT_CHUNK consecutive warp tiles; a warp's first chunk is its global index, the
  // following ones come from an atomic ticket requested a whole chunk ahead (its latency never shows), so that
  // warps on faster SMs take more of the work (a static split left 20 % of the warp slots idle at the tail)
  const uint32_t nWarps = gridDim.x * MG_CNT_WARPS;
  uint64_t tile = (uint64_t)(blockIdx.x * MG_CNT_WARPS + (tid >> 5)) * MG_CNT_CHUNK, tileNext = 0;
  uint32_t pendingChunk = 0;
  if (lane == 0) pendingChunk = nWarps + atomicAdd(P.ticket, 1u);
  // a tile can be bulk-copied when all of it (and the overlap) lies inside the batch; the packed
  // stream and the flags carry slack words past the end, the raw bytes do not
  const uint64_t nBulk = RAW ? (P.nBases >= MG_WS_RAW_BYTES ? (P.nBases - MG_WS_RAW_BYTES) / MG_WT_BASES + 1 : 0) : P.nTiles;
  uint32_t nSelectedLocal = 0;                             // SCATTER: this thread's share of the total
  uint32_t phase = 0;

  if (lane == 0)
    { mg_mbar_init(&S->bar, 1);
      mg_fence_barrier_init();
      mg_fence_proxy_async();
      if (tile < nBulk) count_issue_tile<RAW>(P, S, tile);
    }
  if (LUTK)
    { const uint4 *src = reinterpret_cast<const uint4 *>(P.lut);
      uint4 *dst = reinterpret_cast<uint4 *>(sLut);
      for (uint32_t i = tid; i < MG_LUT_SIZE / 16; i += MG_CNT_THREADS) dst[i] = __ldg(src + i);
    }
  __syncthreads();

  for (; tile < P.nTiles; tile = tileNext)
    { // ---- this lane's two runs: words 2l, 2l+1 (+ overlap word 2l+2), 96 end flags
      const uint32_t run0 = lane * MG_SEL_RPT;
      const uint64_t tileBase = tile * MG_WT_BASES;
      uint64_t w0, w1, w2;
      uint32_t e0, e1, e2;
      if (tile < nBulk)
        { mg_mbar_wait(&S->bar, phase);
          phase ^= 1;
          e0 = S->ends[run0]; e1 = S->ends[run0 + 1]; e2 = S->ends[run0 + 2];
          uint64_t wx = 0;                                  // the overlap word, from lanes 0 and 1
          if (RAW)
            { // a lane's 64 bytes are four 16-byte chunks at a 64-byte stride: read in lane order they would hit the
              // same banks four ways.  Lane pairs start at a rotated chunk instead (conflict-free), and the four
              // packed values are rotated back with two rounds of selects.
              const uint4 *src = reinterpret_cast<const uint4 *>(S->stage) + lane * 4;
              const uint32_t rot = (lane >> 1) & 3u;
              const uint32_t t0 = pack16_dev<ASCII>(src[rot]), t1 = pack16_dev<ASCII>(src[(rot + 1) & 3u]),
                             t2 = pack16_dev<ASCII>(src[(rot + 2) & 3u]), t3 = pack16_dev<ASCII>(src[(rot + 3) & 3u]);
              const bool r1 = rot & 1u, r2 = rot & 2u;
              const uint32_t u0 = r1 ? t3 : t0, u1 = r1 ? t0 : t1, u2 = r1 ? t1 : t2, u3 = r1 ? t2 : t3;
              w0 = ((uint64_t)(r2 ? u2 : u0) << 32) | (r2 ? u3 : u1);
              w1 = ((uint64_t)(r2 ? u0 : u2) << 32) | (r2 ? u1 : u3);
              uint32_t x = 0;
              if (lane < 2) x = pack16_dev<ASCII>(reinterpret_cast<const uint4 *>(S->stage)[128 + lane]);
              wx = ((uint64_t)__shfl_sync(0xffffffffu, x, 0) << 32) | __shfl_sync(0xffffffffu, x, 1);
            }
          else
            { const uint64_t *src = reinterpret_cast<const uint64_t *>(S->stage);
              w0 = src[run0]; w1 = src[run0 + 1];
              wx = src[MG_WT_RUNS];
            }
          S->words[run0] = w0; S->words[run0 + 1] = w1;
          if (lane == 0) S->words[MG_WT_RUNS] = wx;
          // the staged flags must sit in registers before the warp barrier below frees the staging buffers for the next
          // bulk copy: a load the compiler sinks past the copy's issue reads the NEXT tile's flags (seen in hash_count2.cu)
          asm volatile("" : "+r"(e0), "+r"(e1), "+r"(e2) :: "memory");
        }
      else
        { // the ragged end of a raw batch: guarded loads
          const uint64_t word = tile * MG_WT_RUNS + run0;
          e0 = __ldg(P.ends + word); e1 = __ldg(P.ends + word + 1); e2 = __ldg(P.ends + word + 2);
          w0 = pack32_raw<ASCII>(P.raw, word * MG_RUN, P.nBases);
          w1 = pack32_raw<ASCII>(P.raw, word * MG_RUN + 32, P.nBases);
          S->words[run0] = w0; S->words[run0 + 1] = w1;
          if (lane == 31) S->words[MG_WT_RUNS] = pack32_raw<ASCII>(P.raw, word * MG_RUN + 64, P.nBases);
        }
      __syncwarp();
      // every lane has consumed its part of the staged tile (the stores above depend on it): the buffer is
      // free, and the next tile's copy streams in behind the whole computation of this one
      tileNext = tile + 1;
      if ((tileNext & (MG_CNT_CHUNK - 1)) == 0)
        { tileNext = (uint64_t)__shfl_sync(0xffffffffu, pendingChunk, 0) * MG_CNT_CHUNK;
          if (lane == 0 && tileNext < P.nTiles) pendingChunk = nWarps + atomicAdd(P.ticket, 1u);
        }
      if (lane == 0 && tileNext < nBulk) count_issue_tile<RAW>(P, S, tileNext);
      w2 = S->words[run0 + 2];

      uint32_t m0, m1;
      { const uint64_t p0 = tileBase + (uint64_t)run0 * MG_RUN;
        if (LUTK)
          mg_lut_scan<LUTK ? LUTK : 31>(sLut, w0, w1, w2, &m0, &m1);
        else if (PREFILTER)
          { const MgRun RA = mg_run_prepare(w0, w1, H.k);
            m0 = scan_run<PREFILTER>(H, RA);
            const MgRun RB = mg_run_prepare(w1, w2, H.k);
            m1 = scan_run<PREFILTER>(H, RB);
          }
        else
          { // the full scan is ~700 unrolled instructions per run: ONE copy, executed twice, keeps the tile loop
            // inside the 32 KB instruction cache (two copies: 21 % of the stall samples were instruction fetches)
            uint64_t wa = w0, wb = w1;
            m0 = 0; m1 = 0;
#pragma unroll 1
            for (int h = 0; h < 2; ++h)
              { const MgRun RR = mg_run_prepare(wa, wb, H.k);
                const uint32_t mh = scan_run<PREFILTER>(H, RR);
                if (h == 0) m0 = mh; else m1 = mh;
                wa = w1; wb = w2;
              }
          }
        // windows that would span two sequences or run off the batch are not usable; a tile with no sequence end
        // in sight that lies inside the batch (nearly all of them on long sequences) skips the whole computation
        const bool plain = tileBase + MG_WT_BASES + MG_RUN <= P.nBases && !__any_sync(0xffffffffu, (e0 | e1 | e2) != 0u);
        if (!plain)
          { m0 &= mg_run_usable((uint64_t)e0 | ((uint64_t)e1 << 32), H.k, p0, P.nBases);
            m1 &= mg_run_usable((uint64_t)e1 | ((uint64_t)e2 << 32), H.k, p0 + MG_RUN, P.nBases);
          }
      }

      // ---- phase 2: queue of (run, window) of this warp's 64 runs
      const uint32_t cnt = __popc(m0) + __popc(m1);
      const uint32_t incl = mg_warp_incl_scan(cnt);
      const uint32_t nW = __shfl_sync(0xffffffffu, incl, 31);
      if (nW == 0) { __syncwarp(); continue; }
      const bool queued = nW <= MG_WQ_CAP;                 // warp-uniform
      uint16_t *wq = S->queue;
      if (queued)
        { uint32_t qoff = incl - cnt;
          uint32_t mm = m0;
          while (mm) { const uint32_t i = __ffs(mm) - 1; mm &= mm - 1; wq[qoff++] = (uint16_t)((run0 << 5) | i); }
          mm = m1;
          while (mm) { const uint32_t i = __ffs(mm) - 1; mm &= mm - 1; wq[qoff++] = (uint16_t)(((run0 + 1) << 5) | i); }
        }
      __syncwarp();

      // ---- phase 3: evaluation and output
      const uint64_t *sWords = S->words;
      if (queued && nW <= MG_SEL_ROUNDS * 32)
        { // the usual case: every lane evaluates its (<= MG_SEL_ROUNDS) queue entries into registers first
          uint64_t km[MG_SEL_ROUNDS];
          uint32_t ent[MG_SEL_ROUNDS];
          uint32_t okMask = 0, fMask = 0;
#pragma unroll
          for (int r = 0; r < MG_SEL_ROUNDS; ++r)
            { const uint32_t q = r * 32 + lane;
              km[r] = 0; ent[r] = 0;
              if (q < nW)
                { bool isF;
                  ent[r] = wq[q];
                  if (eval_entry_k<LUTK>(H, sWords, ent[r], &km[r], &isF)) { okMask |= 1u << r; if (isF) fMask |= 1u << r; }
                }
            }
          if (OWNERS)
            { // per-owner segments: rank inside the warp's tile through shared memory, one reservation per owner
              uint32_t own[MG_SEL_ROUNDS], rk[MG_SEL_ROUNDS];
              S->own[lane] = 0; S->own[lane + 32] = 0;
              __syncwarp();
#pragma unroll
              for (int r = 0; r < MG_SEL_ROUNDS; ++r)
                if ((okMask >> r) & 1u)
                  { own[r] = mg_owner(km[r], P.nOwners);
                    rk[r] = atomicAdd(&S->own[own[r]], 1u);
                  }
              __syncwarp();
#pragma unroll
              for (uint32_t o = lane; o < 64; o += 32)
                { const uint32_t c = (o < P.nOwners) ? S->own[o] : 0u;
                  S->own[o] = c ? atomicAdd(&P.ownerCursor[o], c) : 0u;
                }
              __syncwarp();
#pragma unroll
              for (int r = 0; r < MG_SEL_ROUNDS; ++r)
                if ((okMask >> r) & 1u)
                  { const uint64_t dst = (uint64_t)S->own[own[r]] + rk[r];
                    if (dst < P.ownerCap) P.ownerBuf[(uint64_t)own[r] * P.ownerCap + dst] = km[r];
                  }
            }
          else if (SCATTER)
            { nSelectedLocal += __popc(okMask);
              uint32_t pos[MG_SEL_ROUNDS], region[MG_SEL_ROUNDS];
#pragma unroll
              for (int r = 0; r < MG_SEL_ROUNDS; ++r)
                if ((okMask >> r) & 1u)
                  { region[r] = (uint32_t)(mg_slot_hash(km[r], P.slotBits) >> P.regionBits);
                    if (PEER) region[r] += mg_owner(km[r], P.nOwners) * P.nRegions;      // bucket index = owner * R + region
                    pos[r] = atomicAdd(&P.cursors[region[r]], 1u);
                  }
#pragma unroll
              for (int r = 0; r < MG_SEL_ROUNDS; ++r)
                if ((okMask >> r) & 1u)
                  { if (pos[r] < P.bucketCap) { uint64_t *bp = P.buckets + (uint64_t)region[r] * P.bucketCap + pos[r]; if (P.keepBuckets) mg_st_keep(bp, km[r]); else *bp = km[r]; }
                    else if (PEER)
                      { const uint32_t ow = region[r] / P.nRegions;
                        const uint32_t o = atomicAdd(&P.ownerCursor[ow], 1u);
                        if (o < P.overflowCap) P.overflow[(uint64_t)ow * P.overflowCap + o] = km[r];
                      }
                    else
                      { const uint32_t o = atomicAdd(&P.cursors[P.nRegions], 1u);
                        if (o < P.overflowCap) P.overflow[o] = km[r];
                      }
                  }
            }
          else
            { // list: one reservation per warp and tile
              const uint32_t c = __popc(okMask);
              const uint32_t inc2 = mg_warp_incl_scan(c);
              const uint32_t total = __shfl_sync(0xffffffffu, inc2, 31);
              unsigned long long wbase = 0;
              if (lane == 0 && total) wbase = atomicAdd(P.count, (unsigned long long)total);
              wbase = __shfl_sync(0xffffffffu, wbase, 0);
              uint64_t dst = wbase + inc2 - c;
#pragma unroll
              for (int r = 0; r < MG_SEL_ROUNDS; ++r)
                if ((okMask >> r) & 1u)
                  { if (dst < P.cap)
                      { P.outKmer[dst] = (P.strandBit && ((fMask >> r) & 1u)) ? (km[r] | (1ull << 63)) : km[r];
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
            { uint64_t km = 0; bool isF = false, ok = false;
              uint32_t e = 0;
              bool have;
              if (queued)
                { if (base >= nW) break;
                  have = base + lane < nW;
                  if (have) e = wq[base + lane];
                }
              else
                { have = (own0 | own1) != 0;
                  if (!__any_sync(0xffffffffu, have)) break;
                  if (own0) { uint32_t i = __ffs(own0) - 1; own0 &= own0 - 1; e = (run0 << 5) | i; }
                  else if (own1) { uint32_t i = __ffs(own1) - 1; own1 &= own1 - 1; e = ((run0 + 1) << 5) | i; }
                }
              if (have) ok = eval_entry_k<LUTK>(H, sWords, e, &km, &isF);
              const uint32_t ballot = __ballot_sync(0xffffffffu, ok);
              if (!ballot) continue;
              if (OWNERS)
                { if (ok)
                    { const uint32_t o = mg_owner(km, P.nOwners);
                      const uint64_t dst = atomicAdd(&P.ownerCursor[o], 1u);
                      if (dst < P.ownerCap) P.ownerBuf[(uint64_t)o * P.ownerCap + dst] = km;
                    }
                  continue;
                }
              if (SCATTER)
                { if (ok)
                    { ++nSelectedLocal;
                      uint32_t region = (uint32_t)(mg_slot_hash(km, P.slotBits) >> P.regionBits);
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
                      if (P.outPos) P.outPos[dst] = (uint32_t)(tileBase + e);   // e = run*32 + window
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
extern "C" uint64_t modgpuHashSelectWorkspace(uint64_t nBases)
{
  uint64_t words = (nBases + 31) / 32;
  uint64_t tiles = (words + MG_TILE_THREADS - 1) / MG_TILE_THREADS;
  return MG_WS_STATUS + tiles * sizeof(uint64_t);  // ticket (+pad), candidate table, then descriptors
}

MgKHasher mg_khasher_from(const ModgpuHasher *h) { return mg_make_khasher(h->k, h->w, h->factor1); }

// count mode with the table-driven prefilter launches lut_build_kernel + hash_count_kernel, everything else one kernel
int mg_select_launches(const ModgpuHasher *h, int flags)
{
  const MgKHasher H = mg_khasher_from(h);
  const bool lut = H.lut && !(flags & (MODGPU_SEL_ORDERED | MODGPU_SEL_NOPREFILTER | MODGPU_SEL_NOLUT));
  return lut ? 2 : 1;
}

template <bool PF, bool TMA>
static int launch_ordered(const SelectParams &P, cudaStream_t st)
{
  static int blocksPerSm = 0;
  if (!blocksPerSm)
    { MG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, hash_select_ordered_kernel<PF, TMA>, MG_SEL_THREADS, 0));
      if (blocksPerSm < 1) blocksPerSm = 1;
    }
  uint64_t grid = (uint64_t)mg_num_sms() * blocksPerSm;
  if (grid > P.nTiles) grid = P.nTiles;
  hash_select_ordered_kernel<PF, TMA><<<(unsigned)grid, MG_SEL_THREADS, 0, st>>>(P);
  MG_LAUNCH_CHECK("hash_select_ordered");
  return MODGPU_OK;
}

template <bool PF, bool RAW, int OUT, bool ASCII, int LUTK>
static int launch_count(const SelectParams &P0, cudaStream_t st)
{
  static int blocksPerSm = 0;
  const size_t smem = (LUTK ? MG_LUT_SIZE : 0) + MG_CNT_WARPS * sizeof(CountWarpSmem<RAW>);
  if (!blocksPerSm)
    { MG_CUDA(cudaFuncSetAttribute(hash_count_kernel<PF, RAW, OUT, ASCII, LUTK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      MG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, hash_count_kernel<PF, RAW, OUT, ASCII, LUTK>, MG_CNT_THREADS, smem));
      if (blocksPerSm < 1) blocksPerSm = 1;
    }
  SelectParams P = P0;
  P.nTiles = (uint32_t)((P.nBases + MG_WT_BASES - 1) / MG_WT_BASES);         // warp tiles
  uint64_t grid = (uint64_t)mg_num_sms() * blocksPerSm;
  const uint64_t need = ((uint64_t)P.nTiles + MG_CNT_WARPS * MG_CNT_CHUNK - 1) / (MG_CNT_WARPS * MG_CNT_CHUNK);
  if (grid > need) grid = need;
  if (LUTK)
    { lut_build_kernel<<<MG_LUT_SIZE / 256, 256, 0, st>>>(P.H, const_cast<uint8_t *>(P.lut));
      MG_LAUNCH_CHECK("lut_build");
    }
  hash_count_kernel<PF, RAW, OUT, ASCII, LUTK><<<(unsigned)grid, MG_CNT_THREADS, smem, st>>>(P);
  MG_LAUNCH_CHECK("hash_count");
  return MODGPU_OK;
}

// count mode dispatch over (scan, input) for one output mode.  scan: 0 full evaluation of every window,
// 1 multiplicative low-word prefilter, 30/31 table-driven prefilter for that k
template <int OUT, bool PF, int LUTK>
static int dispatch_count_load(const SelectParams &P, cudaStream_t st)
{
  if (P.raw) return P.rawAscii ? launch_count<PF, true, OUT, true, LUTK>(P, st) : launch_count<PF, true, OUT, false, LUTK>(P, st);
  return launch_count<PF, false, OUT, false, LUTK>(P, st);
}

template <int OUT>
static int dispatch_count(const SelectParams &P, bool pf, int flags, cudaStream_t st)
{
  // raw bytes with the table-driven or the full 32-bit scan: the second-generation kernel (hash_count2.cu)
  { const int rc = mg_count2_launch(P, OUT, flags, st);
    if (rc != 1) return rc;
  }
  if (!pf) return dispatch_count_load<OUT, false, 0>(P, st);
  if (P.H.lut && !(flags & MODGPU_SEL_NOLUT))
    return P.H.k == 31 ? dispatch_count_load<OUT, true, 31>(P, st) : dispatch_count_load<OUT, true, 30>(P, st);
  return dispatch_count_load<OUT, true, 0>(P, st);
}

extern "C" int modgpuHashSelect(const ModgpuHasher *h, const uint64_t *d_packed, const uint32_t *d_ends,
                                uint64_t nBases, uint64_t *d_kmers, uint32_t *d_gpos, uint64_t cap,
                                uint64_t *d_count, void *d_workspace, int flags, void *stream)
{
  cudaStream_t st = (cudaStream_t)stream;
  if (!h || h->k < 1 || h->k > 31 || h->w < 1) { mg_set_error("modgpuHashSelect: bad hasher"); return MODGPU_EINVAL; }
  if (nBases >= (1ull << 32)) { mg_set_error("modgpuHashSelect: batch of %llu bases exceeds 2^32-1", (unsigned long long)nBases); return MODGPU_EINVAL; }
  MG_CUDA(cudaMemsetAsync(d_count, 0, sizeof(uint64_t), st));
  if (!nBases) return MODGPU_OK;
  SelectParams P;
  memset(&P, 0, sizeof(P));
  P.H = mg_khasher_from(h);
  P.packed = d_packed; P.ends = d_ends; P.nBases = nBases;
  uint64_t words = (nBases + 31) / 32;
  P.nTiles = (uint32_t)((words + MG_TILE_THREADS - 1) / MG_TILE_THREADS);
  P.strandBit = (flags & MODGPU_SEL_STRAND) ? 1u : 0u;
  P.outKmer = d_kmers; P.outPos = d_gpos; P.cap = cap;
  P.count = (unsigned long long *)d_count;
  P.ticket = (uint32_t *)d_workspace;
  P.status = (uint64_t *)((char *)d_workspace + MG_WS_STATUS);
  P.lut = (const uint8_t *)d_workspace + MG_WS_LUT;
  MG_CUDA(cudaMemsetAsync(d_workspace, 0, modgpuHashSelectWorkspace(nBases), st));
  const bool pf = P.H.prefilter && !(flags & MODGPU_SEL_NOPREFILTER);
  const bool ord = (flags & MODGPU_SEL_ORDERED) != 0;
  const bool tma = !(flags & MODGPU_SEL_NOTMA);
  if (!ord) return dispatch_count<0>(P, pf, flags, st);
  if (pf) return tma ? launch_ordered<true, true>(P, st) : launch_ordered<true, false>(P, st);
  return tma ? launch_ordered<false, true>(P, st) : launch_ordered<false, false>(P, st);
}

// K2 fused with the bucket scatter of the bulk insert (count mode): no list.
// cursors must be zeroed by the caller; *d_count receives the number selected.
int mg_hash_select_scatter(const ModgpuHasher *h, const uint64_t *d_packed, const uint32_t *d_ends, uint64_t nBases,
                           uint64_t *d_count, void *d_workspace, int flags, uint32_t slotBits, uint32_t regionBits,
                           uint32_t bucketCap, uint32_t *d_cursors, uint64_t *d_buckets, uint64_t *d_overflow,
                           uint64_t overflowCap, const uint8_t *d_raw, int rawAscii, const uint8_t *d_tileFlags, cudaStream_t st)
{
  if (nBases >= (1ull << 32)) { mg_set_error("hash_select: batch of %llu bases exceeds 2^32-1", (unsigned long long)nBases); return MODGPU_EINVAL; }
  MG_CUDA(cudaMemsetAsync(d_count, 0, sizeof(uint64_t), st));
  if (!nBases) return MODGPU_OK;
  SelectParams P;
  memset(&P, 0, sizeof(P));
  P.H = mg_khasher_from(h);
  P.packed = d_packed; P.ends = d_ends; P.nBases = nBases;
  uint64_t words = (nBases + 31) / 32;
  P.nTiles = (uint32_t)((words + MG_TILE_THREADS - 1) / MG_TILE_THREADS);
  P.count = (unsigned long long *)d_count;
  P.ticket = (uint32_t *)d_workspace;
  P.status = (uint64_t *)((char *)d_workspace + MG_WS_STATUS);
  P.lut = (const uint8_t *)d_workspace + MG_WS_LUT;
  P.slotBits = slotBits; P.regionBits = regionBits; P.nRegions = 1u << (slotBits - regionBits); P.bucketCap = bucketCap;
  P.cursors = d_cursors; P.buckets = d_buckets; P.overflow = d_overflow; P.overflowCap = overflowCap;
  { static int keep = -1; if (keep < 0) { const char *v = getenv("MODGPU_KEEP_BUCKETS"); keep = v ? atoi(v) : 0; } P.keepBuckets = (uint32_t)keep; }
  MG_CUDA(cudaMemsetAsync(d_workspace, 0, 64, st));
  const bool pf = P.H.prefilter && !(flags & MODGPU_SEL_NOPREFILTER);
  const bool tma = !(flags & MODGPU_SEL_NOTMA);
  if (d_raw) { P.raw = d_raw; P.rawAscii = rawAscii ? 1u : 0u; }
  P.tileFlags = d_tileFlags;
  return dispatch_count<1>(P, pf, flags, st);
}

// K2 with the selected k-mers bucketed by owner GPU (multi-GPU count mode): segment o of d_buf
// (ownerCap entries each) receives the k-mers owned by rank o, d_cursors[o] (zeroed here) their number
// (which exceeds ownerCap when the segment overflowed: the caller must then fall back to the list path).
int mg_hash_select_owners(const ModgpuHasher *h, const uint64_t *d_packed, const uint32_t *d_ends, uint64_t nBases,
                          void *d_workspace, int flags, uint32_t nOwners, uint32_t *d_cursors, uint64_t *d_buf,
                          uint64_t ownerCap, cudaStream_t st)
{
  if (nBases >= (1ull << 32)) { mg_set_error("hash_select: batch of %llu bases exceeds 2^32-1", (unsigned long long)nBases); return MODGPU_EINVAL; }
  if (nOwners < 1 || nOwners > 64) { mg_set_error("hash_select: nOwners %u out of range 1..64", nOwners); return MODGPU_EINVAL; }
  MG_CUDA(cudaMemsetAsync(d_cursors, 0, nOwners * sizeof(uint32_t), st));
  if (!nBases) return MODGPU_OK;
  SelectParams P;
  memset(&P, 0, sizeof(P));
  P.H = mg_khasher_from(h);
  P.packed = d_packed; P.ends = d_ends; P.nBases = nBases;
  uint64_t words = (nBases + 31) / 32;
  P.nTiles = (uint32_t)((words + MG_TILE_THREADS - 1) / MG_TILE_THREADS);
  P.ticket = (uint32_t *)d_workspace;
  P.status = (uint64_t *)((char *)d_workspace + MG_WS_STATUS);
  P.lut = (const uint8_t *)d_workspace + MG_WS_LUT;
  P.count = (unsigned long long *)((char *)d_workspace + 8);       // unused total
  P.nOwners = nOwners; P.ownerCursor = d_cursors; P.ownerBuf = d_buf; P.ownerCap = ownerCap;
  MG_CUDA(cudaMemsetAsync(d_workspace, 0, 64, st));
  const bool pf = P.H.prefilter && !(flags & MODGPU_SEL_NOPREFILTER);
  const bool tma = !(flags & MODGPU_SEL_NOTMA);
  return dispatch_count<2>(P, pf, flags, st);
}

// K2 with the selected k-mers written into per-(owner, region) buckets (multi-GPU, fully fused):
// bucket (o * nRegions + r) of d_buckets (bucketCap entries) holds the k-mers owned by rank o that fall
// into region r of its table; d_cursors[o * nRegions + r] (zeroed here) counts them; k-mers beyond a
// bucket's capacity go to the owner's overflow segment d_overflow[o * overflowCap ..] / d_ovfCounts[o].
int mg_hash_select_peer(const ModgpuHasher *h, const uint64_t *d_packed, const uint32_t *d_ends, uint64_t nBases,
                        uint64_t *d_count, void *d_workspace, int flags, uint32_t slotBits, uint32_t regionBits,
                        uint32_t nOwners, uint32_t bucketCap, uint32_t *d_cursors, uint64_t *d_buckets,
                        uint64_t *d_overflow, uint64_t overflowCap, uint32_t *d_ovfCounts,
                        const uint8_t *d_raw, int rawAscii, const uint8_t *d_tileFlags, cudaStream_t st)
{
  if (nBases >= (1ull << 32)) { mg_set_error("hash_select: batch of %llu bases exceeds 2^32-1", (unsigned long long)nBases); return MODGPU_EINVAL; }
  if (nOwners < 1 || nOwners > 64) { mg_set_error("hash_select: nOwners %u out of range 1..64", nOwners); return MODGPU_EINVAL; }
  const uint32_t nRegions = 1u << (slotBits - regionBits);
  if (!(flags & MODGPU_SEL_APPEND))                    // APPEND: the batch joins what earlier batches left in the buckets
    { MG_CUDA(cudaMemsetAsync(d_count, 0, sizeof(uint64_t), st));
      MG_CUDA(cudaMemsetAsync(d_cursors, 0, (size_t)nOwners * nRegions * sizeof(uint32_t), st));
      MG_CUDA(cudaMemsetAsync(d_ovfCounts, 0, nOwners * sizeof(uint32_t), st));
    }
  if (!nBases) return MODGPU_OK;
  SelectParams P;
  memset(&P, 0, sizeof(P));
  P.H = mg_khasher_from(h);
  P.packed = d_packed; P.ends = d_ends; P.nBases = nBases;
  uint64_t words = (nBases + 31) / 32;
  P.nTiles = (uint32_t)((words + MG_TILE_THREADS - 1) / MG_TILE_THREADS);
  P.count = (unsigned long long *)d_count;
  P.ticket = (uint32_t *)d_workspace;
  P.status = (uint64_t *)((char *)d_workspace + MG_WS_STATUS);
  P.lut = (const uint8_t *)d_workspace + MG_WS_LUT;
  P.slotBits = slotBits; P.regionBits = regionBits; P.nRegions = nRegions; P.bucketCap = bucketCap;
  P.cursors = d_cursors; P.buckets = d_buckets; P.overflow = d_overflow; P.overflowCap = overflowCap;
  P.nOwners = nOwners; P.ownerCursor = d_ovfCounts;
  { static int keep = -1; if (keep < 0) { const char *v = getenv("MODGPU_KEEP_PEER_BUCKETS"); keep = v ? atoi(v) : 1; }
    P.keepBuckets = (uint32_t)keep; }                  // many small buckets: their tail sectors must survive in L2
  MG_CUDA(cudaMemsetAsync(d_workspace, 0, 64, st));
  const bool pf = P.H.prefilter && !(flags & MODGPU_SEL_NOPREFILTER);
  const bool tma = !(flags & MODGPU_SEL_NOTMA);
  if (d_raw) { P.raw = d_raw; P.rawAscii = rawAscii ? 1u : 0u; }
  P.tileFlags = d_tileFlags;
  return dispatch_count<3>(P, pf, flags, st);
}

// ---------------------------------------------------------------- locate --
// global offset -> (sequence id, offset in sequence): the id/offset columns
// the reference fills at modmap.c:113-116.  Binary search over the offsets.
__global__ void __launch_bounds__(256) locate_kernel(const uint32_t *__restrict__ gpos, uint64_t n,
                                                     const uint64_t *__restrict__ offs, uint64_t nSeq,
                                                     uint32_t *__restrict__ id, uint32_t *__restrict__ pos)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    { uint64_t g = gpos[i];
      uint64_t lo = 0, hi = nSeq;                     // last r with offs[r] <= g
      while (hi - lo > 1)
        { uint64_t mid = (lo + hi) >> 1;
          if (__ldg(offs + mid) <= g) lo = mid; else hi = mid;
        }
      if (id) id[i] = (uint32_t)lo;
      if (pos) pos[i] = (uint32_t)(g - __ldg(offs + lo));
    }
}

extern "C" int modgpuLocate(const uint32_t *d_gpos, uint64_t n, const uint64_t *d_offs, uint64_t nSeq,
                            uint32_t *d_id, uint32_t *d_pos, void *stream)
{
  if (!n) return MODGPU_OK;
  if (!nSeq) { mg_set_error("modgpuLocate: no sequences"); return MODGPU_EINVAL; }
  uint64_t blocks = (n + 255) / 256;
  uint64_t maxBlocks = (uint64_t)mg_num_sms() * 8;
  if (blocks > maxBlocks) blocks = maxBlocks;
  locate_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_gpos, n, d_offs, nSeq, d_id, d_pos);
  MG_LAUNCH_CHECK("locate");
  return MODGPU_OK;
}
