// api.cu - host-level batched modset: the reference's caller loops as C-ABI
// calls taking HOST (or device) buffers.
//
//   modgpuModsetAdd  ==  the addSequence loop of addSequenceFile
//                        (reference modutils.c:19-51): for every sequence,
//                        every selected modimizer is found-or-inserted and its
//                        depth incremented.
//
// Pipeline per chunk of whole sequences: H2D (copy stream, double buffered,
// overlapped with the previous chunk's kernels) -> K1 pack2bit + end flags ->
// K2 hash_select -> K3 table insert.  There is no CPU path: every step is a
// kernel launch or a CUDA copy, and errors surface as return codes.
#include <string.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>
#include "mg_device.cuh"
#include "mg_api.h"

// ---------------------------------------------------------------- create --
extern "C" ModgpuModset *modgpuModsetCreateWithHasher(int bits, const ModgpuHasher *h);

extern "C" ModgpuModset *modgpuModsetCreate(int bits, int k, int w, int seed)
{
  ModgpuHasher h;
  if (modgpuHasherInit(&h, k, w, seed)) return nullptr;
  return modgpuModsetCreateWithHasher(bits, &h);
}

// modsetCreate with an existing hasher (e.g. the Seqhash stored in a .mod file, modset.c:96-97)
extern "C" ModgpuModset *modgpuModsetCreateWithHasher(int bits, const ModgpuHasher *hasher)
{
  if (modgpuDeviceCount() < 1)
    { if (!modgpuLastError()[0]) mg_set_error("no CUDA device: libmodgpu has no CPU fallback");
      return nullptr;
    }
  if (!hasher || hasher->k < 1 || hasher->k > 31 || hasher->w < 1 || !(hasher->factor1 & 1))
    { mg_set_error("modsetCreate: bad hasher"); return nullptr; }
  ModgpuModset *ms = new ModgpuModset();
  ms->hasher = *hasher;
  ms->bits = bits;
  cudaGetDevice(&ms->device);
  if (mg_check_cuda(cudaStreamCreateWithFlags(&ms->stream, cudaStreamNonBlocking), "cudaStreamCreate", __FILE__, __LINE__) ||
      mg_check_cuda(cudaStreamCreateWithFlags(&ms->copyStream, cudaStreamNonBlocking), "cudaStreamCreate", __FILE__, __LINE__))
    { delete ms; return nullptr; }
  ms->ownStream = true;
  for (int i = 0; i < 2; ++i)
    { cudaEventCreateWithFlags(&ms->evCopied[i], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&ms->evFree[i], cudaEventDisableTiming);
    }
  ms->table = modgpuTableCreate(bits, ms->stream);
  if (!ms->table) { modgpuModsetDestroy(ms); return nullptr; }
  if (ms->misc.ensure(16384) || ms->hMisc.ensure(4096)) { modgpuModsetDestroy(ms); return nullptr; }
  return ms;
}

extern "C" void modgpuModsetDestroy(ModgpuModset *ms)
{
  if (!ms) return;
  if (ms->stream) cudaStreamSynchronize(ms->stream);
  if (ms->copyStream) cudaStreamSynchronize(ms->copyStream);
  prof_collect(ms);
  for (cudaEvent_t e : ms->evPool) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i)
    { ms->bases[i].release(); ms->offs[i].release(); ms->hOffs[i].release(); ms->pk[i].release();
      if (ms->evCopied[i]) cudaEventDestroy(ms->evCopied[i]);
      if (ms->evFree[i]) cudaEventDestroy(ms->evFree[i]);
    }
  ms->packed.release(); ms->ends.release(); ms->kmers.release(); ms->kmers2.release(); ms->gpos.release(); ms->slot.release();
  ms->work.release(); ms->misc.release(); ms->expo.release(); ms->hMisc.release();
  if (ms->table) modgpuTableDestroy(ms->table);
  if (ms->ownStream && ms->stream) cudaStreamDestroy(ms->stream);
  if (ms->copyStream) cudaStreamDestroy(ms->copyStream);
  delete ms;
}

extern "C" const ModgpuHasher *modgpuModsetHasher(const ModgpuModset *ms) { return &ms->hasher; }
extern "C" int modgpuModsetDevice(const ModgpuModset *ms) { return ms->device; }
extern "C" ModgpuTable *modgpuModsetTable(ModgpuModset *ms) { return ms->table; }

extern "C" int modgpuModsetSetStream(ModgpuModset *ms, void *stream)
{
  MG_CUDA(cudaStreamSynchronize(ms->stream));
  if (ms->ownStream) cudaStreamDestroy(ms->stream);
  ms->stream = (cudaStream_t)stream;
  ms->ownStream = false;
  return MODGPU_OK;
}

extern "C" int modgpuModsetSetFlags(ModgpuModset *ms, int flags)
{
  ms->selFlags = flags & 0xFF00FF;
  // bits 8..15: insert locality override + 1 (0 = auto): 1 = off, 2.. = 2^(v-1) regions
  const int v = (flags >> 8) & 0xFF;
  ms->regionBits = v ? v - 1 : -1;
  return MODGPU_OK;
}
extern "C" int modgpuModsetSetExactOrder(ModgpuModset *ms, int e) { ms->exactOrder = e ? 1 : 0; return MODGPU_OK; }

extern "C" int modgpuModsetProfile(ModgpuModset *ms, int enable)
{
  MG_CUDA(cudaStreamSynchronize(ms->stream));
  prof_collect(ms);
  ms->profile = enable != 0;
  for (int i = 0; i < MODGPU_T_N; ++i) { ms->ms[i] = 0; ms->launches[i] = 0; }
  return MODGPU_OK;
}

extern "C" int modgpuModsetTimes(ModgpuModset *ms, double ms_out[MODGPU_T_N], uint64_t launches_out[MODGPU_T_N])
{
  MG_CUDA(cudaStreamSynchronize(ms->stream));
  prof_collect(ms);
  for (int i = 0; i < MODGPU_T_N; ++i) { if (ms_out) ms_out[i] = ms->ms[i]; if (launches_out) launches_out[i] = ms->launches[i]; }
  return MODGPU_OK;
}

// End flags with a clean-buffer invariant: between batches ms->ends is all zero, so a batch only pays for its own
// nSeq flags - set before the select, cleared right after it - instead of a 1-bit-per-base memset (0.06 ms per
// 3.1 Gbases).  Any early return between the two leaves endsCleanCap at 0 and the next batch clears the buffer.
static int ends_mark(ModgpuModset *ms, const uint64_t *d_offs, uint64_t nSeq, uint64_t nBases, cudaStream_t st)
{
  // the per-base flag words, then one byte per 2048-base tile of the count kernels (set and cleared with the flags)
  const uint64_t flagBytes = modgpuEndsWords(nBases) * 4, nTiles = (nBases + 2047) / 2048;
  int rc = ms->ends.ensure(flagBytes + nTiles + 64);
  if (rc) return rc;
  if (ms->endsCleanCap != ms->ends.cap) MG_CUDA(cudaMemsetAsync(ms->ends.p, 0, ms->ends.cap, st));
  ms->endsCleanCap = 0;
  // sparse when at most one tile in four can hold a sequence end (genomes, long reads); short reads keep per-tile staging
  ms->tileFlags = (nSeq * 4 <= nTiles) ? (uint8_t *)ms->ends.p + flagBytes : nullptr;
  ms->tileFlagsAt = (uint8_t *)ms->ends.p + flagBytes;
  return mg_ends_sparse(d_offs, nSeq, (uint32_t *)ms->ends.p, ms->tileFlags, 1, st);
}

static int ends_unmark(ModgpuModset *ms, const uint64_t *d_offs, uint64_t nSeq, cudaStream_t st)
{
  ProfScope p(ms, MODGPU_T_PACK, 1);
  int rc = mg_ends_sparse(d_offs, nSeq, (uint32_t *)ms->ends.p, ms->tileFlags, 0, st);
  if (rc) return rc;
  ms->endsCleanCap = ms->ends.cap;
  return MODGPU_OK;
}

// ----------------------------------------------------------------- chunks --
// K1 + K2 over one device-resident chunk; leaves the selected k-mers (and, if
// wantPos, their global offsets) in ms->kmers / ms->gpos and returns the count.
// Synchronises the stream once (to learn the count and size the next kernels).
int mg_modset_select_chunk(ModgpuModset *ms, const uint8_t *d_bases, const uint64_t *d_offs,
                           uint64_t nSeq, uint64_t nBases, int isAscii, bool wantPos, int extraFlags,
                           uint64_t *nSelected)
{
  cudaStream_t st = ms->stream;
  *nSelected = 0;
  if (!nBases) return MODGPU_OK;
  const uint64_t words = modgpuPackedWords(nBases);
  int rc;
  if ((rc = ms->packed.ensure(words * 8)) || (rc = ms->ends.ensure(words * 4)) ||
      (rc = ms->work.ensure(modgpuHashSelectWorkspace(nBases))))
    return rc;
  { ProfScope p(ms, MODGPU_T_PACK, 2);
    if ((rc = modgpuPack2bit(d_bases, nBases, isAscii, (uint64_t *)ms->packed.p, st))) return rc;
    if ((rc = ends_mark(ms, d_offs, nSeq, nBases, st))) return rc;
  }
  const uint64_t d = (uint64_t)ms->hasher.w;
  uint64_t cap = (d <= 2) ? nBases : (nBases / d + nBases / (4 * d) + 65536);
  if (cap > nBases) cap = nBases;
  uint64_t *dCount = (uint64_t *)ms->misc.p;
  volatile uint64_t *hCount = (volatile uint64_t *)ms->hMisc.p;
  const int flags = ms->selFlags | extraFlags | (ms->exactOrder ? MODGPU_SEL_ORDERED : 0);
  for (int attempt = 0; attempt < 2; ++attempt)
    { if ((rc = ms->kmers.ensure(cap * 8))) return rc;
      if (wantPos && (rc = ms->gpos.ensure(cap * 4))) return rc;
      { ProfScope p(ms, MODGPU_T_SELECT, mg_select_launches(&ms->hasher, ms->selFlags | (ms->exactOrder ? MODGPU_SEL_ORDERED : 0)));
        if ((rc = modgpuHashSelect(&ms->hasher, (const uint64_t *)ms->packed.p, (const uint32_t *)ms->ends.p, nBases,
                                   (uint64_t *)ms->kmers.p, wantPos ? (uint32_t *)ms->gpos.p : nullptr, cap, dCount,
                                   ms->work.p, flags, st)))
          return rc;
      }
      MG_CUDA(cudaMemcpyAsync((void *)hCount, dCount, 8, cudaMemcpyDeviceToHost, st));
      MG_CUDA(cudaStreamSynchronize(st));
      if (*hCount <= cap) break;
      cap = *hCount;                                     // denser than expected: redo with the exact size
    }
  if ((rc = ends_unmark(ms, d_offs, nSeq, st))) return rc;
  *nSelected = *hCount;
  return MODGPU_OK;
}

// K3, three ways:
//  - bulk (default for long lists in count mode): scatter into per-region buckets,
//    build every region in shared memory, stream the table once (table.cu);
//  - direct: one random HBM probe per k-mer (short lists, and exact-order lists,
//    whose input ordinals matter);
//  - region-partitioned direct (A/B only, MODGPU flags bits 8..15 = 2..): group
//    the list by 2^rb table regions first so that the probes of one region hit L2.
static int insert_list(ModgpuModset *ms, const uint64_t *d_kmers, uint64_t n, uint32_t *dSlot)
{
  int rc;
  const int rb = ms->regionBits;                       // -1 auto, 0 direct, 1..8 partitioned direct, 254 bulk
  const bool bulk = !ms->exactOrder && (rb == 254 || (rb < 0 && n >= mg_table_bulk_threshold(ms->table)));
  if (bulk)
    { ProfScope p(ms, MODGPU_T_INSERT, 3);
      return mg_table_insert_bulk(ms->table, d_kmers, n, ms->stream);
    }
  if (rb > 0 && rb <= 8 && !ms->exactOrder)
    { uint32_t slotBits = 0;
      while ((1ull << slotBits) < modgpuTableSlots(ms->table)) ++slotBits;
      if ((rc = ms->kmers2.ensure(n * 8))) return rc;
      uint64_t *scratch = (uint64_t *)((char *)ms->misc.p + 512);
      { ProfScope p(ms, MODGPU_T_INSERT, 3);
        if ((rc = mg_slot_partition(d_kmers, n, slotBits, (uint32_t)rb, (uint64_t *)ms->kmers2.p, scratch, ms->stream))) return rc;
      }
      d_kmers = (const uint64_t *)ms->kmers2.p;
    }
  ProfScope p(ms, MODGPU_T_INSERT, 1);
  return mg_table_insert_dev(ms->table, d_kmers, nullptr, n, dSlot, ms->exactOrder, ms->stream);
}

// count mode, long batch: K1 -> K2 with the selected k-mers scattered straight
// into the table's per-region buckets -> shared-memory region build.  No list.
// Returns 1 when the batch turned out too skewed for the buckets (nothing was
// applied to the table; the caller falls back to the list path).
static int add_chunk_fused(ModgpuModset *ms, const uint8_t *d_bases, const uint64_t *d_offs, uint64_t nSeq,
                           uint64_t nBases, int isAscii, uint64_t expected, uint64_t *nHashes)
{
  cudaStream_t st = ms->stream;
  const uint64_t words = modgpuPackedWords(nBases);
  int rc;
  // K1 fused into K2's tile loader when the batch is 16-byte aligned: no packed stream is written at all
  const bool fusePack = ((((uintptr_t)d_bases) & 15) == 0) && !(ms->selFlags & MODGPU_SEL_NOFUSEPACK);
  if ((rc = (fusePack ? MODGPU_OK : ms->packed.ensure(words * 8))) || (rc = ms->ends.ensure(words * 4)) ||
      (rc = ms->work.ensure(modgpuHashSelectWorkspace(nBases))))
    return rc;
  { ProfScope p(ms, MODGPU_T_PACK, fusePack ? 1 : 2);
    if (!fusePack && (rc = modgpuPack2bit(d_bases, nBases, isAscii, (uint64_t *)ms->packed.p, st))) return rc;
    if ((rc = ends_mark(ms, d_offs, nSeq, nBases, st))) return rc;
  }
  MgBulk b;
  const bool deferred = ms->accumulate > 1;              // the buckets stay open over several chunks
  const bool wasPending = mg_table_clear_pending(ms->table);
  if ((rc = deferred ? mg_table_bulk_open(ms->table, expected, ms->accumulate, &b, st)
                     : mg_table_bulk_begin(ms->table, expected, 2 * expected + 65536, &b, st)))
    return rc;
  uint64_t *dCount = (uint64_t *)ms->misc.p;
  volatile uint64_t *hCount = (volatile uint64_t *)ms->hMisc.p;
  { ProfScope p(ms, MODGPU_T_SELECT, mg_select_launches(&ms->hasher, ms->selFlags | (ms->exactOrder ? MODGPU_SEL_ORDERED : 0)));
    if ((rc = mg_hash_select_scatter(&ms->hasher, (const uint64_t *)ms->packed.p, (const uint32_t *)ms->ends.p, nBases, dCount,
                                     ms->work.p, ms->selFlags, b.slotBits, b.regionBits, b.cap, b.cursors, b.buckets,
                                     b.overflow, b.overflowCap, fusePack ? d_bases : nullptr, isAscii, ms->tileFlags, st)))
      return rc;
  }
  // the build follows without a host round trip: its kernels check on the device that the overflow list was
  // large enough and leave the table untouched otherwise (skewed batch: the caller falls back to the list path)
  { static int gap = -1; if (gap < 0) { const char *v = getenv("MODGPU_GAP"); gap = v ? atoi(v) : 0; }
    if (gap) MG_CUDA(cudaStreamSynchronize(st));
  }
  if (!deferred)
    { ProfScope p(ms, MODGPU_T_INSERT, 2);
      if ((rc = mg_table_bulk_finish(ms->table, &b, st))) return rc;
    }
  if ((rc = ends_unmark(ms, d_offs, nSeq, st))) return rc;
  MG_CUDA(cudaMemcpyAsync((void *)hCount, dCount, 8, cudaMemcpyDeviceToHost, st));
  MG_CUDA(cudaMemcpyAsync((void *)(hCount + 1), mg_table_bulk_overflow_count(ms->table), 4, cudaMemcpyDeviceToHost, st));
  MG_CUDA(cudaStreamSynchronize(st));
  const uint64_t overflowed = hCount[1] & 0xFFFFFFFFull;  // deferred: cumulative over the chunks waiting in the buckets
  if (overflowed > b.overflowCap)                        // buckets and overflow list too small: nothing was applied
    { if (deferred)
        { ProfScope p(ms, MODGPU_T_INSERT, 2);
          if ((rc = mg_table_bulk_rollback(ms->table, st))) return rc;    // builds what was waiting before this chunk
          ms->dirty = true;
        }
      else mg_table_bulk_abort(ms->table, wasPending);
      return 1;
    }
  if (deferred) mg_table_bulk_commit(ms->table, expected, overflowed);
  *nHashes = hCount[0];
  ms->dirty = true;
  return MODGPU_OK;
}

static int add_chunk_device(ModgpuModset *ms, const uint8_t *d_bases, const uint64_t *d_offs, uint64_t nSeq,
                            uint64_t nBases, int isAscii, uint64_t *nHashes)
{
  uint64_t n = 0;
  int rc;
  const uint64_t expected = nBases / (uint64_t)ms->hasher.w;
  // the bucket path pays one pass over the table per build: worth it when the k-mers sharing that pass (this chunk,
  // or up to `accumulate` chunks in deferred mode) are at least 1/8 of the slots
  const bool fuse = !ms->exactOrder && ms->hasher.w >= 4 && !(ms->selFlags & 0x10) &&
                    (ms->regionBits == 254 ||
                     (ms->regionBits < 0 && expected * (uint64_t)ms->accumulate >= mg_table_bulk_threshold(ms->table)));
  if (fuse)
    { rc = add_chunk_fused(ms, d_bases, d_offs, nSeq, nBases, isAscii, expected, nHashes);
      if (rc <= 0) return rc;
    }
  // exact order: entries an earlier count-mode / bulk / merge insert left un-numbered get their (slot-order) indices
  // first, otherwise a reappearance in this batch would number them as first occurrences of THIS batch (modset.c:57)
  if (ms->exactOrder && ms->dirty && (rc = mg_modset_ensure_numbered(ms))) return rc;
  rc = mg_modset_select_chunk(ms, d_bases, d_offs, nSeq, nBases, isAscii, false, 0, &n);
  if (rc) return rc;
  *nHashes = n;
  if (!n) return MODGPU_OK;
  uint32_t *dSlot = nullptr;
  if (ms->exactOrder)
    { if ((rc = ms->slot.ensure(n * 4))) return rc;
      dSlot = (uint32_t *)ms->slot.p;
    }
  if ((rc = insert_list(ms, (const uint64_t *)ms->kmers.p, n, dSlot))) return rc;
  if (ms->exactOrder)
    { ProfScope p(ms, MODGPU_T_OTHER, 3);
      if ((rc = modgpuTableNumber(ms->table, dSlot, n, nullptr, ms->stream))) return rc;
    }
  else ms->dirty = true;
  return MODGPU_OK;
}

// greedy groups of whole sequences of at most `limit` bases
static void plan_chunks(const uint64_t *offs, uint64_t nSeq, uint64_t limit, std::vector<uint64_t> &cuts)
{
  cuts.clear();
  cuts.push_back(0);
  uint64_t r = 0;
  while (r < nSeq)
    { uint64_t r1 = r + 1;
      while (r1 < nSeq && offs[r1 + 1] - offs[r] <= limit) ++r1;
      cuts.push_back(r1);
      r = r1;
    }
}

static int check_offsets(const uint64_t *offs, uint64_t nSeq)
{
  if (!offs || offs[0] != 0) { mg_set_error("offsets must start at 0"); return MODGPU_EINVAL; }
  for (uint64_t r = 0; r < nSeq; ++r)
    { if (offs[r + 1] < offs[r]) { mg_set_error("offsets must be non-decreasing (sequence %llu)", (unsigned long long)r); return MODGPU_EINVAL; }
      if (offs[r + 1] - offs[r] > 0x7FFFFFFFull)         // reference: int len, seqhash.h:50
        { mg_set_error("sequence %llu longer than 2^31-1", (unsigned long long)r); return MODGPU_EINVAL; }
    }
  return MODGPU_OK;
}

// stage chunk c (sequences cuts[c]..cuts[c+1]) into device buffer c&1
static int stage_chunk(ModgpuModset *ms, const char *bases, const uint64_t *offs, const std::vector<uint64_t> &cuts, size_t c)
{
  const int b = (int)(c & 1);
  const uint64_t r0 = cuts[c], r1 = cuts[c + 1];
  const uint64_t nb = offs[r1] - offs[r0], ns = r1 - r0;
  int rc;
  // the device buffer is free once the kernels of chunk c-2 are done
  MG_CUDA(cudaStreamWaitEvent(ms->copyStream, ms->evFree[b], 0));
  // the pinned offsets buffer is free once the copy of chunk c-2 has completed
  MG_CUDA(cudaEventSynchronize(ms->evCopied[b]));
  if ((rc = ms->bases[b].ensure(nb + 64)) || (rc = ms->offs[b].ensure((ns + 1) * 8)) || (rc = ms->hOffs[b].ensure((ns + 1) * 8)))
    return rc;
  uint64_t *ho = (uint64_t *)ms->hOffs[b].p;
  for (uint64_t r = 0; r <= ns; ++r) ho[r] = offs[r0 + r] - offs[r0];
  if (nb) MG_CUDA(cudaMemcpyAsync(ms->bases[b].p, bases + offs[r0], nb, cudaMemcpyHostToDevice, ms->copyStream));
  MG_CUDA(cudaMemcpyAsync(ms->offs[b].p, ho, (ns + 1) * 8, cudaMemcpyHostToDevice, ms->copyStream));
  MG_CUDA(cudaEventRecord(ms->evCopied[b], ms->copyStream));
  return MODGPU_OK;
}

// The same double-buffered staging for the modmap entry points (refmap.cu): chunk c+1 crosses PCIe on the copy stream
// while the kernels of chunk c run.  begin plans the chunks and starts the first copy; chunk(c) starts the copy of
// c+1, makes the compute stream wait for c and hands out its device buffers; release(c) marks the buffers reusable.
struct MgFeed { const char *bases; const uint64_t *offs; std::vector<uint64_t> cuts; };

MgFeed *mg_feed_begin(ModgpuModset *ms, const char *bases, const uint64_t *offs, uint64_t nSeq, uint64_t limit, size_t *nChunks)
{
  MgFeed *f = new MgFeed();
  f->bases = bases; f->offs = offs;
  plan_chunks(offs, nSeq, limit, f->cuts);
  *nChunks = f->cuts.size() - 1;
  if (*nChunks && stage_chunk(ms, bases, offs, f->cuts, 0)) { delete f; return nullptr; }
  return f;
}

int mg_feed_chunk(ModgpuModset *ms, MgFeed *f, size_t c, const uint8_t **d_bases, const uint64_t **d_offs, uint64_t *r0, uint64_t *r1)
{
  const int b = (int)(c & 1);
  int rc;
  if (c + 2 < f->cuts.size() && (rc = stage_chunk(ms, f->bases, f->offs, f->cuts, c + 1))) return rc;
  MG_CUDA(cudaStreamWaitEvent(ms->stream, ms->evCopied[b], 0));
  *d_bases = (const uint8_t *)ms->bases[b].p; *d_offs = (const uint64_t *)ms->offs[b].p;
  *r0 = f->cuts[c]; *r1 = f->cuts[c + 1];
  return MODGPU_OK;
}

int mg_feed_release(ModgpuModset *ms, size_t c)
{
  MG_CUDA(cudaEventRecord(ms->evFree[c & 1], ms->stream));
  return MODGPU_OK;
}

void mg_feed_end(ModgpuModset *ms, MgFeed *f)
{ // a copy still in flight (early exit) must not outlive the caller's buffers
  cudaStreamSynchronize(ms->copyStream);
  delete f;
}

extern "C" uint64_t modgpuModsetAdd(ModgpuModset *ms, const char *bases, const uint64_t *offs, uint64_t nSeq, int isAscii)
{
  const uint64_t FAIL = 0xFFFFFFFFFFFFFFFFull;
  if (!ms || !offs) { mg_set_error("modgpuModsetAdd: null argument"); return FAIL; }
  if (!nSeq) return 0;
  if (check_offsets(offs, nSeq)) return FAIL;
  if (!bases && offs[nSeq]) { mg_set_error("modgpuModsetAdd: null bases"); return FAIL; }
  std::vector<uint64_t> cuts;
  plan_chunks(offs, nSeq, MG_HOST_CHUNK, cuts);
  const size_t nChunks = cuts.size() - 1;
  uint64_t total = 0;
  // a buffer whose allocation grows must not be in use: growth only happens while staging,
  // after the waits in stage_chunk have retired the previous user of that buffer
  if (stage_chunk(ms, bases, offs, cuts, 0)) return FAIL;
  for (size_t c = 0; c < nChunks; ++c)
    { const int b = (int)(c & 1);
      if (c + 1 < nChunks && stage_chunk(ms, bases, offs, cuts, c + 1)) return FAIL;
      if (mg_check_cuda(cudaStreamWaitEvent(ms->stream, ms->evCopied[b], 0), "cudaStreamWaitEvent", __FILE__, __LINE__)) return FAIL;
      uint64_t n = 0;
      const uint64_t r0 = cuts[c], r1 = cuts[c + 1];
      if (add_chunk_device(ms, (const uint8_t *)ms->bases[b].p, (const uint64_t *)ms->offs[b].p, r1 - r0,
                           offs[r1] - offs[r0], isAscii, &n))
        return FAIL;
      if (mg_check_cuda(cudaEventRecord(ms->evFree[b], ms->stream), "cudaEventRecord", __FILE__, __LINE__)) return FAIL;
      total += n;
    }
  ms->totalHashes += total;
  if (ms->accumulate <= 1 && modgpuTableEntries(ms->table, ms->stream) == FAIL) return FAIL;      // reference: die() on overflow
  return total;
}

extern "C" uint64_t modgpuModsetAddDevice(ModgpuModset *ms, const uint8_t *d_bases, const uint64_t *d_offs,
                                          uint64_t nSeq, uint64_t nBases, int isAscii)
{
  const uint64_t FAIL = 0xFFFFFFFFFFFFFFFFull;
  if (!ms) { mg_set_error("modgpuModsetAddDevice: null modset"); return FAIL; }
  if (!nSeq || !nBases) return 0;
  uint64_t total = 0;
  if (nBases <= MG_DEV_CHUNK)
    { if (add_chunk_device(ms, d_bases, d_offs, nSeq, nBases, isAscii, &total)) return FAIL; }
  else
    { // split at sequence boundaries: needs the offsets on the host
      std::vector<uint64_t> ho(nSeq + 1);
      if (mg_check_cuda(cudaMemcpyAsync(ho.data(), d_offs, (nSeq + 1) * 8, cudaMemcpyDeviceToHost, ms->stream), "offsets readback", __FILE__, __LINE__) ||
          mg_check_cuda(cudaStreamSynchronize(ms->stream), "sync", __FILE__, __LINE__))
        return FAIL;
      if (check_offsets(ho.data(), nSeq)) return FAIL;
      std::vector<uint64_t> cuts;
      plan_chunks(ho.data(), nSeq, MG_DEV_CHUNK, cuts);
      for (size_t c = 0; c + 1 < cuts.size(); ++c)
        { const uint64_t r0 = cuts[c], r1 = cuts[c + 1];
          const uint64_t nb = ho[r1] - ho[r0], ns = r1 - r0;
          if (nb >= (1ull << 32)) { mg_set_error("sequence group of %llu bases exceeds 2^32-1", (unsigned long long)nb); return FAIL; }
          // local offsets for the group
          if (ms->offs[0].ensure((ns + 1) * 8) || ms->hOffs[0].ensure((ns + 1) * 8)) return FAIL;
          if (mg_check_cuda(cudaStreamSynchronize(ms->stream), "sync", __FILE__, __LINE__)) return FAIL;
          uint64_t *hl = (uint64_t *)ms->hOffs[0].p;
          for (uint64_t r = 0; r <= ns; ++r) hl[r] = ho[r0 + r] - ho[r0];
          if (mg_check_cuda(cudaMemcpyAsync(ms->offs[0].p, hl, (ns + 1) * 8, cudaMemcpyHostToDevice, ms->stream), "offsets upload", __FILE__, __LINE__)) return FAIL;
          uint64_t n = 0;
          if (add_chunk_device(ms, d_bases + ho[r0], (const uint64_t *)ms->offs[0].p, ns, nb, isAscii, &n)) return FAIL;
          total += n;
        }
    }
  ms->totalHashes += total;
  if (ms->accumulate <= 1 && modgpuTableEntries(ms->table, ms->stream) == FAIL) return FAIL;     // deferred: modgpuModsetFlush reports
  return total;
}

// ------------------------------------------------------------ packed input --
// The reference's own 2-bit sequence packing (sqioSeqPack, seqio.c:557-570: what its "binary" seqio files hold):
// four bases per byte, first base in the top two bits, every sequence starting on a byte, and a last byte with fewer
// than four bases holding them in its LOW bits.  A caller that has such records ships 0.25 bytes per base over PCIe
// instead of 1; the device expands them to codes (one streaming pass, 1.25 B/base) and the batch then takes the
// normal path.  One thread expands 16 bases: five packed bytes, a funnel shift, and per output word one multiply
// that spreads a byte's four codes (b * 0x40100401 >> 6 & 0x03030303) - unless the group touches a sequence boundary.
__global__ void __launch_bounds__(256) unpack_seqio_kernel(const uint8_t *__restrict__ packed, const uint64_t *__restrict__ byteOffs,
                                                           const uint64_t *__restrict__ baseOffs, uint64_t nSeq, uint64_t nBases,
                                                           uint8_t *__restrict__ out)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t nGroups = (nBases + 15) / 16;
  for (uint64_t grp = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; grp < nGroups; grp += stride)
    { const uint64_t g0 = grp * 16;
      uint64_t lo = 0, hi = nSeq;                          // last sequence starting at or before g0 (it holds base g0)
      while (hi - lo > 1) { const uint64_t mid = (lo + hi) >> 1; if (__ldg(baseOffs + mid) <= g0) lo = mid; else hi = mid; }
      uint64_t r = lo;
      const uint64_t s0 = __ldg(baseOffs + r), s1 = __ldg(baseOffs + r + 1);
      const uint64_t j0 = g0 - s0, len = s1 - s0, whole = len & ~(uint64_t)3;   // bases in fully packed bytes
      uint4 w;
      if (j0 + 16 <= whole)
        { const uint8_t *p = packed + __ldg(byteOffs + r) + (j0 >> 2);
          const uint64_t v = ((uint64_t)p[0] << 32) | ((uint64_t)p[1] << 24) | ((uint64_t)p[2] << 16) | ((uint64_t)p[3] << 8) | (uint64_t)p[4];
          const uint32_t x = (uint32_t)(v >> (8 - 2 * (uint32_t)(j0 & 3)));      // 16 bases, the first in the top two bits
          w.x = (((x >> 24) * 0x40100401u) >> 6) & 0x03030303u;
          w.y = ((((x >> 16) & 0xFFu) * 0x40100401u) >> 6) & 0x03030303u;
          w.z = ((((x >> 8) & 0xFFu) * 0x40100401u) >> 6) & 0x03030303u;
          w.w = (((x & 0xFFu) * 0x40100401u) >> 6) & 0x03030303u;
        }
      else
        { uint32_t c[4] = { 0, 0, 0, 0 };
          for (uint32_t t = 0; t < 16; ++t)
            { const uint64_t g = g0 + t;
              if (g >= nBases) break;
              while (g >= __ldg(baseOffs + r + 1)) ++r;
              const uint64_t a = __ldg(baseOffs + r), j = g - a, L = __ldg(baseOffs + r + 1) - a;
              const uint32_t b = packed[__ldg(byteOffs + r) + (j >> 2)], rem = (uint32_t)(L & 3);
              const bool tail = rem && (j >> 2) == ((L - 1) >> 2);               // the short last byte is right-aligned
              const uint32_t shift = tail ? 2 * (rem - 1 - (uint32_t)(j & 3)) : 6 - 2 * (uint32_t)(j & 3);
              c[t >> 2] |= ((b >> shift) & 3u) << (8 * (t & 3));
            }
          w.x = c[0]; w.y = c[1]; w.z = c[2]; w.w = c[3];
        }
      reinterpret_cast<uint4 *>(out)[grp] = w;
    }
}

extern "C" uint64_t modgpuModsetAddPacked(ModgpuModset *ms, const uint8_t *packed, const uint64_t *byteOffs,
                                          const uint64_t *offs, uint64_t nSeq)
{
  const uint64_t FAIL = 0xFFFFFFFFFFFFFFFFull;
  if (!ms || !offs || !byteOffs) { mg_set_error("modgpuModsetAddPacked: null argument"); return FAIL; }
  if (!nSeq) return 0;
  if (check_offsets(offs, nSeq)) return FAIL;
  if (!packed && offs[nSeq]) { mg_set_error("modgpuModsetAddPacked: null bases"); return FAIL; }
  for (uint64_t r = 0; r < nSeq; ++r)
    if (byteOffs[r + 1] < byteOffs[r] + (offs[r + 1] - offs[r] + 3) / 4)
      { mg_set_error("modgpuModsetAddPacked: sequence %llu has fewer packed bytes than (len+3)/4", (unsigned long long)r); return FAIL; }
  std::vector<uint64_t> cuts;
  plan_chunks(offs, nSeq, MG_HOST_CHUNK, cuts);
  const size_t nChunks = cuts.size() - 1;
  cudaStream_t st = ms->stream;
  uint64_t total = 0;
  // per chunk: the packed bytes and the two rebased offset arrays cross PCIe on the copy stream into buffer c & 1 while the
  // kernels of chunk c - 1 run (the same events as the byte path); pkOffs[b] holds byteOffs then baseOffs
  auto stage = [&](size_t c) -> int {
    const int b = (int)(c & 1);
    const uint64_t r0 = cuts[c], r1 = cuts[c + 1], ns = r1 - r0, nbytes = byteOffs[r1] - byteOffs[r0];
    int rc;
    MG_CUDA(cudaStreamWaitEvent(ms->copyStream, ms->evFree[b], 0));
    MG_CUDA(cudaEventSynchronize(ms->evCopied[b]));
    if ((rc = ms->pk[b].ensure(nbytes + 64)) || (rc = ms->offs[b].ensure(2 * (ns + 1) * 8)) || (rc = ms->hOffs[b].ensure(2 * (ns + 1) * 8))) return rc;
    uint64_t *ho = (uint64_t *)ms->hOffs[b].p;
    for (uint64_t r = 0; r <= ns; ++r) { ho[r] = byteOffs[r0 + r] - byteOffs[r0]; ho[ns + 1 + r] = offs[r0 + r] - offs[r0]; }
    if (nbytes) MG_CUDA(cudaMemcpyAsync(ms->pk[b].p, packed + byteOffs[r0], nbytes, cudaMemcpyHostToDevice, ms->copyStream));
    MG_CUDA(cudaMemcpyAsync(ms->offs[b].p, ho, 2 * (ns + 1) * 8, cudaMemcpyHostToDevice, ms->copyStream));
    MG_CUDA(cudaEventRecord(ms->evCopied[b], ms->copyStream));
    return MODGPU_OK;
  };
  if (stage(0)) return FAIL;
  for (size_t c = 0; c < nChunks; ++c)
    { const int b = (int)(c & 1);
      if (c + 1 < nChunks && stage(c + 1)) return FAIL;
      if (mg_check_cuda(cudaStreamWaitEvent(st, ms->evCopied[b], 0), "cudaStreamWaitEvent", __FILE__, __LINE__)) return FAIL;
      const uint64_t r0 = cuts[c], r1 = cuts[c + 1], ns = r1 - r0, nb = offs[r1] - offs[r0];
      uint64_t n = 0;
      if (nb)
        { if (ms->bases[0].ensure(nb + 64)) return FAIL;
          const uint64_t *dByte = (const uint64_t *)ms->offs[b].p, *dBase = dByte + ns + 1;
          { ProfScope p(ms, MODGPU_T_PACK, 1);
            uint64_t blocks = ((nb + 15) / 16 + 255) / 256, maxBlocks = (uint64_t)mg_num_sms() * 16;
            if (blocks > maxBlocks) blocks = maxBlocks;
            unpack_seqio_kernel<<<(unsigned)blocks, 256, 0, st>>>((const uint8_t *)ms->pk[b].p, dByte, dBase, ns, nb, (uint8_t *)ms->bases[0].p);
            if (mg_check_cuda(cudaGetLastError(), "unpack_seqio", __FILE__, __LINE__)) return FAIL;
          }
          if (add_chunk_device(ms, (const uint8_t *)ms->bases[0].p, dBase, ns, nb, 0, &n)) return FAIL;
        }
      if (mg_check_cuda(cudaEventRecord(ms->evFree[b], st), "cudaEventRecord", __FILE__, __LINE__)) return FAIL;
      total += n;
    }
  ms->totalHashes += total;
  if (ms->accumulate <= 1 && modgpuTableEntries(ms->table, st) == FAIL) return FAIL;
  return total;
}

extern "C" int modgpuModsetSelectDevice(ModgpuModset *ms, const uint8_t *d_bases, const uint64_t *d_offs,
                                        uint64_t nSeq, uint64_t nBases, int isAscii,
                                        const uint64_t **d_kmers, uint64_t *nSelected)
{
  if (nBases >= (1ull << 32)) { mg_set_error("modgpuModsetSelectDevice: batch exceeds 2^32-1 bases"); return MODGPU_EINVAL; }
  int rc = mg_modset_select_chunk(ms, d_bases, d_offs, nSeq, nBases, isAscii, false, 0, nSelected);
  if (rc) return rc;
  *d_kmers = (const uint64_t *)ms->kmers.p;
  return MODGPU_OK;
}

extern "C" int modgpuModsetSelectHost(ModgpuModset *ms, const char *bases, const uint64_t *offs,
                                      uint64_t nSeq, int isAscii, const uint64_t **d_kmers, uint64_t *nSelected)
{
  *nSelected = 0; *d_kmers = nullptr;
  if (!nSeq) return MODGPU_OK;
  int rc = check_offsets(offs, nSeq);
  if (rc) return rc;
  const uint64_t nb = offs[nSeq];
  if (nb >= (1ull << 32)) { mg_set_error("modgpuModsetSelectHost: batch exceeds 2^32-1 bases"); return MODGPU_EINVAL; }
  if ((rc = ms->bases[0].ensure(nb + 64)) || (rc = ms->offs[0].ensure((nSeq + 1) * 8))) return rc;
  if (nb) MG_CUDA(cudaMemcpyAsync(ms->bases[0].p, bases, nb, cudaMemcpyHostToDevice, ms->stream));
  MG_CUDA(cudaMemcpyAsync(ms->offs[0].p, offs, (nSeq + 1) * 8, cudaMemcpyHostToDevice, ms->stream));
  return modgpuModsetSelectDevice(ms, (const uint8_t *)ms->bases[0].p, (const uint64_t *)ms->offs[0].p, nSeq, nb, isAscii, d_kmers, nSelected);
}

// ------------------------------------------------------------ scanner --
// modRCiterator / modRCnext over a batch of sequences (seqhash.c:154-196) with HOST buffers on both sides:
// the modimizers of every sequence in (sequence, position) order - k-mer (bit 63 = isForward), position inside
// its sequence, and seqOff[r] .. seqOff[r+1] = the results of sequence r.  Returns the total, UINT64_MAX on error;
// when the total exceeds cap only the first cap results are stored.
struct ModgpuScanner { ModgpuModset *ms; DevBuf id, pos; };

extern "C" ModgpuScanner *modgpuScannerCreate(const ModgpuHasher *h)
{
  ModgpuModset *ms = modgpuModsetCreateWithHasher(20, h);        // only its hasher, stream and scratch buffers are used
  if (!ms) return nullptr;
  ModgpuScanner *sc = new ModgpuScanner();
  sc->ms = ms;
  return sc;
}

extern "C" void modgpuScannerDestroy(ModgpuScanner *sc)
{
  if (!sc) return;
  sc->id.release(); sc->pos.release();
  modgpuModsetDestroy(sc->ms);
  delete sc;
}

extern "C" uint64_t modgpuScannerScan(ModgpuScanner *sc, const char *bases, const uint64_t *offs, uint64_t nSeq, int isAscii,
                                      uint64_t *kmers, uint32_t *pos, uint64_t *seqOff, uint64_t cap)
{
  const uint64_t FAIL = 0xFFFFFFFFFFFFFFFFull;
  ModgpuModset *ms = sc->ms;
  cudaStream_t st = ms->stream;
  if (seqOff) for (uint64_t r = 0; r <= nSeq; ++r) seqOff[r] = 0;
  if (!nSeq) return 0;
  if (check_offsets(offs, nSeq)) return FAIL;
  const uint64_t nb = offs[nSeq];
  if (nb >= (1ull << 32)) { mg_set_error("modgpuScannerScan: batch exceeds 2^32-1 bases"); return FAIL; }
  if (!nb) return 0;
  if (ms->bases[0].ensure(nb + 64) || ms->offs[0].ensure((nSeq + 1) * 8)) return FAIL;
  if (mg_check_cuda(cudaMemcpyAsync(ms->bases[0].p, bases, nb, cudaMemcpyHostToDevice, st), "H2D bases", __FILE__, __LINE__) ||
      mg_check_cuda(cudaMemcpyAsync(ms->offs[0].p, offs, (nSeq + 1) * 8, cudaMemcpyHostToDevice, st), "H2D offsets", __FILE__, __LINE__))
    return FAIL;
  uint64_t n = 0;
  if (mg_modset_select_chunk(ms, (const uint8_t *)ms->bases[0].p, (const uint64_t *)ms->offs[0].p, nSeq, nb, isAscii, true,
                             MODGPU_SEL_ORDERED | MODGPU_SEL_STRAND, &n))
    return FAIL;
  const uint64_t m = n < cap ? n : cap;
  if (m || (seqOff && n))                               // the per-sequence ranges need the ids even when nothing is stored
    { if (sc->id.ensure(n * 4) || sc->pos.ensure(n * 4)) return FAIL;
      // global offset -> (sequence, position in sequence)
      if (modgpuLocate((const uint32_t *)ms->gpos.p, n, (const uint64_t *)ms->offs[0].p, nSeq, (uint32_t *)sc->id.p, (uint32_t *)sc->pos.p, st))
        return FAIL;
      if (m && kmers && mg_check_cuda(cudaMemcpyAsync(kmers, ms->kmers.p, m * 8, cudaMemcpyDeviceToHost, st), "D2H kmers", __FILE__, __LINE__)) return FAIL;
      if (m && pos && mg_check_cuda(cudaMemcpyAsync(pos, sc->pos.p, m * 4, cudaMemcpyDeviceToHost, st), "D2H pos", __FILE__, __LINE__)) return FAIL;
    }
  std::vector<uint32_t> id;
  if (seqOff && n)
    { id.resize(n);
      if (mg_check_cuda(cudaMemcpyAsync(id.data(), sc->id.p, n * 4, cudaMemcpyDeviceToHost, st), "D2H ids", __FILE__, __LINE__)) return FAIL;
    }
  if (mg_check_cuda(cudaStreamSynchronize(st), "cudaStreamSynchronize", __FILE__, __LINE__)) return FAIL;
  if (seqOff && n)
    { // results are in sequence order: a counting pass gives the per-sequence ranges
      for (uint64_t i = 0; i < n; ++i) ++seqOff[id[i] + 1];
      for (uint64_t r = 0; r < nSeq; ++r) seqOff[r + 1] += seqOff[r];
    }
  return n;
}

// multi-GPU, sync-free: K1 + K2 with the selected k-mers written into nOwners segments of the caller's
// send buffer (segment o = k-mers owned by rank o), counts in d_counts (uint32 per owner)
extern "C" int modgpuModsetSelectOwnersDevice(ModgpuModset *ms, const uint8_t *d_bases, const uint64_t *d_offs,
                                              uint64_t nSeq, uint64_t nBases, int isAscii, uint32_t nOwners,
                                              uint64_t *d_segments, uint64_t segCap, uint32_t *d_counts)
{
  cudaStream_t st = ms->stream;
  if (nBases >= (1ull << 32)) { mg_set_error("modgpuModsetSelectOwnersDevice: batch exceeds 2^32-1 bases"); return MODGPU_EINVAL; }
  const uint64_t words = modgpuPackedWords(nBases);
  int rc;
  if ((rc = ms->packed.ensure(words * 8)) || (rc = ms->ends.ensure(words * 4)) ||
      (rc = ms->work.ensure(modgpuHashSelectWorkspace(nBases))))
    return rc;
  { ProfScope p(ms, MODGPU_T_PACK, 2);
    if ((rc = modgpuPack2bit(d_bases, nBases, isAscii, (uint64_t *)ms->packed.p, st))) return rc;
    if ((rc = ends_mark(ms, d_offs, nSeq, nBases, st))) return rc;
  }
  { ProfScope p(ms, MODGPU_T_SELECT, mg_select_launches(&ms->hasher, ms->selFlags | (ms->exactOrder ? MODGPU_SEL_ORDERED : 0)));
    if ((rc = mg_hash_select_owners(&ms->hasher, (const uint64_t *)ms->packed.p, (const uint32_t *)ms->ends.p, nBases, ms->work.p,
                                    ms->selFlags, nOwners, d_counts, d_segments, segCap, st)))
      return rc;
  }
  return ends_unmark(ms, d_offs, nSeq, st);
}

extern "C" int modgpuModsetSelectOwnersHost(ModgpuModset *ms, const char *bases, const uint64_t *offs, uint64_t nSeq,
                                            int isAscii, uint32_t nOwners, uint64_t *d_segments, uint64_t segCap,
                                            uint32_t *d_counts)
{
  if (!nSeq) return MODGPU_OK;
  int rc = check_offsets(offs, nSeq);
  if (rc) return rc;
  const uint64_t nb = offs[nSeq];
  if (nb >= (1ull << 32)) { mg_set_error("modgpuModsetSelectOwnersHost: batch exceeds 2^32-1 bases"); return MODGPU_EINVAL; }
  if ((rc = ms->bases[0].ensure(nb + 64)) || (rc = ms->offs[0].ensure((nSeq + 1) * 8))) return rc;
  if (nb) MG_CUDA(cudaMemcpyAsync(ms->bases[0].p, bases, nb, cudaMemcpyHostToDevice, ms->stream));
  MG_CUDA(cudaMemcpyAsync(ms->offs[0].p, offs, (nSeq + 1) * 8, cudaMemcpyHostToDevice, ms->stream));
  return modgpuModsetSelectOwnersDevice(ms, (const uint8_t *)ms->bases[0].p, (const uint64_t *)ms->offs[0].p, nSeq, nb, isAscii,
                                        nOwners, d_segments, segCap, d_counts);
}

extern "C" uint32_t modgpuModsetRegions(ModgpuModset *ms) { return mg_table_regions(ms->table); }

// multi-GPU, fully fused: K1 + K2 scattering into per-(owner, region) buckets of the caller's send buffer
extern "C" int modgpuModsetSelectBucketsDevice(ModgpuModset *ms, const uint8_t *d_bases, const uint64_t *d_offs,
                                               uint64_t nSeq, uint64_t nBases, int isAscii, uint32_t nOwners,
                                               uint64_t *d_buckets, uint32_t bucketCap, uint32_t *d_cursors,
                                               uint64_t *d_overflow, uint64_t overflowCap, uint32_t *d_ovfCounts,
                                               uint64_t *d_count)
{
  cudaStream_t st = ms->stream;
  if (nBases >= (1ull << 32)) { mg_set_error("modgpuModsetSelectBucketsDevice: batch exceeds 2^32-1 bases"); return MODGPU_EINVAL; }
  const uint64_t words = modgpuPackedWords(nBases);
  int rc;
  const bool fusePack = ((((uintptr_t)d_bases) & 15) == 0) && !(ms->selFlags & MODGPU_SEL_NOFUSEPACK);
  if ((rc = (fusePack ? MODGPU_OK : ms->packed.ensure(words * 8))) || (rc = ms->ends.ensure(words * 4)) ||
      (rc = ms->work.ensure(modgpuHashSelectWorkspace(nBases))))
    return rc;
  { ProfScope p(ms, MODGPU_T_PACK, fusePack ? 1 : 2);
    if (!fusePack && (rc = modgpuPack2bit(d_bases, nBases, isAscii, (uint64_t *)ms->packed.p, st))) return rc;
    if ((rc = ends_mark(ms, d_offs, nSeq, nBases, st))) return rc;
  }
  { ProfScope p(ms, MODGPU_T_SELECT, mg_select_launches(&ms->hasher, ms->selFlags | (ms->exactOrder ? MODGPU_SEL_ORDERED : 0)));
    if ((rc = mg_hash_select_peer(&ms->hasher, (const uint64_t *)ms->packed.p, (const uint32_t *)ms->ends.p, nBases, d_count, ms->work.p,
                                  ms->selFlags, mg_table_slot_bits(ms->table), 11, nOwners, bucketCap, d_cursors, d_buckets,
                                  d_overflow, overflowCap, d_ovfCounts, fusePack ? d_bases : nullptr, isAscii, ms->tileFlags, st)))
      return rc;
  }
  return ends_unmark(ms, d_offs, nSeq, st);
}

extern "C" int modgpuModsetSelectBucketsHost(ModgpuModset *ms, const char *bases, const uint64_t *offs, uint64_t nSeq,
                                             int isAscii, uint32_t nOwners, uint64_t *d_buckets, uint32_t bucketCap,
                                             uint32_t *d_cursors, uint64_t *d_overflow, uint64_t overflowCap,
                                             uint32_t *d_ovfCounts, uint64_t *d_count)
{
  if (!nSeq) return MODGPU_OK;
  int rc = check_offsets(offs, nSeq);
  if (rc) return rc;
  const uint64_t nb = offs[nSeq];
  if (nb >= (1ull << 32)) { mg_set_error("modgpuModsetSelectBucketsHost: batch exceeds 2^32-1 bases"); return MODGPU_EINVAL; }
  if ((rc = ms->bases[0].ensure(nb + 64)) || (rc = ms->offs[0].ensure((nSeq + 1) * 8))) return rc;
  if (nb) MG_CUDA(cudaMemcpyAsync(ms->bases[0].p, bases, nb, cudaMemcpyHostToDevice, ms->stream));
  MG_CUDA(cudaMemcpyAsync(ms->offs[0].p, offs, (nSeq + 1) * 8, cudaMemcpyHostToDevice, ms->stream));
  return modgpuModsetSelectBucketsDevice(ms, (const uint8_t *)ms->bases[0].p, (const uint64_t *)ms->offs[0].p, nSeq, nb, isAscii,
                                         nOwners, d_buckets, bucketCap, d_cursors, d_overflow, overflowCap, d_ovfCounts, d_count);
}

// build the object's table regions from nSrc received bucket arrays (+ overflow segments)
extern "C" int modgpuModsetBuildFromBuckets(ModgpuModset *ms, const uint64_t *d_buckets, const uint32_t *d_cursors,
                                            uint32_t bucketCap, uint32_t nSrc, const uint64_t *d_overflow,
                                            uint64_t overflowCap, const uint32_t *d_ovfCounts)
{
  ProfScope p(ms, MODGPU_T_INSERT, 2);
  int rc = mg_table_build_from_buckets(ms->table, d_buckets, d_cursors, bucketCap, nSrc, d_overflow, overflowCap, d_ovfCounts, ms->stream);
  if (rc) return rc;
  ms->dirty = true;
  return MODGPU_OK;
}

// the same with the buckets left in the source ranks' memory: the region build reads them over NVLink
extern "C" int modgpuModsetBuildFromPeers(ModgpuModset *ms, const uint64_t *const *d_buckets, const uint32_t *d_cursors,
                                          uint32_t bucketCap, uint32_t nSrc, const uint64_t *const *d_overflow,
                                          uint64_t overflowCap, const uint32_t *d_ovfCounts)
{
  ProfScope p(ms, MODGPU_T_INSERT, 2);
  int rc = mg_table_build_from_peers(ms->table, d_buckets, d_cursors, bucketCap, nSrc, d_overflow, overflowCap, d_ovfCounts, ms->stream);
  if (rc) return rc;
  ms->dirty = true;
  return MODGPU_OK;
}

// insert + count nSegs received segments (device counts, uint32 each) into the object's table
extern "C" int modgpuModsetInsertSegments(ModgpuModset *ms, const uint64_t *d_segments, uint32_t nSegs, uint64_t segCap,
                                          const uint32_t *d_counts, uint64_t expectedN)
{
  ProfScope p(ms, MODGPU_T_INSERT, 3);
  int rc = mg_table_insert_segments(ms->table, d_segments, nSegs, segCap, d_counts, expectedN, ms->stream);
  if (rc) return rc;
  ms->dirty = true;
  return MODGPU_OK;
}

extern "C" int modgpuModsetInsertDevice(ModgpuModset *ms, const uint64_t *d_kmers, uint64_t n)
{
  if (!n) return MODGPU_OK;
  { const int keep = ms->exactOrder;
    ms->exactOrder = 0;
    int rc = insert_list(ms, d_kmers, n, nullptr);
    ms->exactOrder = keep;
    if (rc) return rc;
  }
  ms->dirty = true;
  ms->totalHashes += n;
  return MODGPU_OK;
}

// Deferred build for streaming many batches into one set: the selected k-mers of up to nChunks device chunks wait in
// the table's region buckets and the regions are built once for all of them.  Results are identical (counts are sums);
// what changes is WHEN a full table is reported: by modgpuModsetFlush, or by the first call that reads the set (every
// reader flushes), instead of by the modgpuModsetAdd* call that overfilled it.  nChunks <= 1 restores the default.
extern "C" int modgpuModsetSetAccumulate(ModgpuModset *ms, int nChunks)
{
  if (!ms) { mg_set_error("modgpuModsetSetAccumulate: null modset"); return MODGPU_EINVAL; }
  if (nChunks > 64) nChunks = 64;
  if (nChunks <= 1)
    { int rc = mg_table_bulk_close(ms->table, ms->stream);
      if (rc) return rc;
    }
  ms->accumulate = nChunks > 1 ? (uint32_t)nChunks : 1u;
  return MODGPU_OK;
}

extern "C" int modgpuModsetFlush(ModgpuModset *ms)
{
  if (!ms) { mg_set_error("modgpuModsetFlush: null modset"); return MODGPU_EINVAL; }
  { ProfScope p(ms, MODGPU_T_INSERT, mg_table_bulk_is_open(ms->table) ? 2 : 0);
    int rc = mg_table_bulk_close(ms->table, ms->stream);
    if (rc) return rc;
  }
  return modgpuTableEntries(ms->table, ms->stream) == 0xFFFFFFFFFFFFFFFFull ? MODGPU_EFULL : MODGPU_OK;
}

extern "C" int modgpuModsetClear(ModgpuModset *ms)
{
  ProfScope p(ms, MODGPU_T_OTHER, 1);
  ms->dirty = false; ms->totalHashes = 0;
  return modgpuTableClear(ms->table, ms->stream);
}

// ------------------------------------------------------------- whole set --
static int ensure_numbered(ModgpuModset *ms)
{
  if (!ms->dirty) return MODGPU_OK;
  ProfScope p(ms, MODGPU_T_OTHER, 3);
  int rc = modgpuTableNumber(ms->table, nullptr, 0, nullptr, ms->stream);
  if (rc) return rc;
  ms->dirty = false;
  return MODGPU_OK;
}

int mg_modset_ensure_numbered(ModgpuModset *ms) { return ensure_numbered(ms); }
void mg_modset_mark(ModgpuModset *ms, bool dirty, bool depthIsZero) { ms->dirty = dirty; ms->depthIsZero = depthIsZero; }
extern "C" int modgpuModsetBits(const ModgpuModset *ms) { return ms->bits; }
cudaStream_t mg_modset_stream(ModgpuModset *ms) { return ms->stream; }
void *mg_modset_kmers(ModgpuModset *ms) { return ms->kmers.p; }
void *mg_modset_gpos(ModgpuModset *ms) { return ms->gpos.p; }

extern "C" uint32_t modgpuModsetMax(ModgpuModset *ms)
{
  // ms->max is the number of distinct entries: the device counter has it without
  // numbering the entries (numbering is only needed to export or look up)
  uint64_t e = modgpuTableEntries(ms->table, ms->stream);
  if (e == 0xFFFFFFFFFFFFFFFFull) return 0xFFFFFFFFu;
  return (uint32_t)e;
}

extern "C" int modgpuModsetExport(ModgpuModset *ms, uint64_t *value, uint16_t *depth, uint8_t *info)
{
  int rc = ensure_numbered(ms);
  if (rc) return rc;
  const uint64_t n = mg_table_numbered(ms->table);
  if (!n) return MODGPU_OK;
  cudaStream_t st = ms->stream;
  if ((rc = ms->expo.ensure(n * 11 + 64))) return rc;
  uint64_t *dV = (uint64_t *)ms->expo.p;
  uint16_t *dD = (uint16_t *)(dV + n);
  uint8_t *dI = (uint8_t *)(dD + n);
  { ProfScope p(ms, MODGPU_T_OTHER, 1);
    if ((rc = modgpuTableExport(ms->table, dV, dD, dI, nullptr, st))) return rc;
  }
  if (value) MG_CUDA(cudaMemcpyAsync(value, dV, n * 8, cudaMemcpyDeviceToHost, st));
  if (depth)
    { if (ms->depthIsZero) memset(depth, 0, n * 2);
      else MG_CUDA(cudaMemcpyAsync(depth, dD, n * 2, cudaMemcpyDeviceToHost, st));
    }
  if (info) MG_CUDA(cudaMemcpyAsync(info, dI, n, cudaMemcpyDeviceToHost, st));
  MG_CUDA(cudaStreamSynchronize(st));
  return MODGPU_OK;
}

extern "C" int modgpuModsetHistogram(ModgpuModset *ms, uint32_t *bins65536)
{
  int rc = ensure_numbered(ms);
  if (rc) return rc;
  if ((rc = ms->expo.ensure(65536 * 4))) return rc;
  if (ms->depthIsZero)
    { memset(bins65536, 0, 65536 * 4); bins65536[0] = (uint32_t)mg_table_numbered(ms->table); return MODGPU_OK; }
  { ProfScope p(ms, MODGPU_T_OTHER, 1);
    if ((rc = modgpuTableHistogram(ms->table, (uint32_t *)ms->expo.p, ms->stream))) return rc;
  }
  MG_CUDA(cudaMemcpyAsync(bins65536, ms->expo.p, 65536 * 4, cudaMemcpyDeviceToHost, ms->stream));
  MG_CUDA(cudaStreamSynchronize(ms->stream));
  return MODGPU_OK;
}

// mode: 0 thresholds (-s), 1 copyM only (-sM), 2 exact multiplicity (modmap), 3 tally only
static int classify(ModgpuModset *ms, int mode, int c1, int c2, int cM, uint32_t classCounts[4])
{
  int rc = ensure_numbered(ms);
  if (rc) return rc;
  uint32_t *dC = (uint32_t *)((char *)ms->misc.p + 256);
  { ProfScope p(ms, MODGPU_T_OTHER, 1);
    if ((rc = mg_table_classify(ms->table, mode, c1, c2, cM, ms->depthIsZero ? 1 : 0, dC, ms->stream))) return rc;
  }
  uint32_t *hC = (uint32_t *)((char *)ms->hMisc.p + 256);
  MG_CUDA(cudaMemcpyAsync(hC, dC, 16, cudaMemcpyDeviceToHost, ms->stream));
  MG_CUDA(cudaStreamSynchronize(ms->stream));
  if (classCounts) memcpy(classCounts, hC, 16);
  return MODGPU_OK;
}

int mg_modset_classify(ModgpuModset *ms, int mode, int c1, int c2, int cM, uint32_t classCounts[4])
{ return classify(ms, mode, c1, c2, cM, classCounts); }

extern "C" int modgpuModsetSetCopy(ModgpuModset *ms, int c1, int c2, int cM, uint32_t classCounts[4])
{ return classify(ms, 0, c1, c2, cM, classCounts); }

extern "C" int modgpuModsetSetCopyM(ModgpuModset *ms, int cM, uint32_t classCounts[4])
{ return classify(ms, 1, 0, 0, cM, classCounts); }

extern "C" int modgpuModsetFind(ModgpuModset *ms, const uint64_t *kmers, uint64_t n, uint32_t *index, uint8_t *copy)
{
  int rc = ensure_numbered(ms);
  if (rc) return rc;
  if (!n) return MODGPU_OK;
  cudaStream_t st = ms->stream;
  if ((rc = ms->kmers.ensure(n * 8)) || (rc = ms->slot.ensure(n * 4))) return rc;
  MG_CUDA(cudaMemcpyAsync(ms->kmers.p, kmers, n * 8, cudaMemcpyHostToDevice, st));
  { ProfScope p(ms, MODGPU_T_OTHER, 1);
    if ((rc = mg_table_lookup_dev(ms->table, (const uint64_t *)ms->kmers.p, nullptr, n, (uint32_t *)ms->slot.p, st))) return rc;
  }
  std::vector<uint32_t> aux(n);
  MG_CUDA(cudaMemcpyAsync(aux.data(), ms->slot.p, n * 4, cudaMemcpyDeviceToHost, st));
  MG_CUDA(cudaStreamSynchronize(st));
  for (uint64_t i = 0; i < n; ++i)
    { if (index) index[i] = aux[i] >> 2;
      if (copy) copy[i] = (uint8_t)(aux[i] & 3u);
    }
  return MODGPU_OK;
}

// modsetSummary (reference modset.c:130-153): the histogram and the class
// tallies come from the device, only the printf stays here.  Keeps the
// reference's 32-bit products (depth * binCount is U32 * U32 there).
extern "C" int modgpuModsetSummary(ModgpuModset *ms, char *buf, int n)
{
  if (ensure_numbered(ms)) return -1;
  const ModgpuHasher *h = &ms->hasher;
  const uint32_t max = (uint32_t)mg_table_numbered(ms->table);
  int o = 0;
  o += snprintf(buf + o, (size_t)(n - o), "SH k %d  w/m %d  s %d\n", h->k, h->w, h->seed);
  o += snprintf(buf + o, (size_t)(n - o), "MS table bits %d size %llu number of entries %u", ms->bits,
                (unsigned long long)(1ull << ms->bits), max);
  if (!max) { o += snprintf(buf + o, (size_t)(n - o), "\n"); return o; }
  std::vector<uint32_t> bins(65536);
  uint32_t copy[4];
  if (modgpuModsetHistogram(ms, bins.data()) || classify(ms, 3, 0, 0, 0, copy)) return -1;
  uint32_t top = 0;
  for (uint32_t i = 0; i < 65536; ++i) if (bins[i]) top = i + 1;
  uint64_t sum = 0, tot = 0;
  for (uint32_t i = 0; i < top; ++i) { sum += bins[i]; tot += (uint32_t)(i * bins[i]); }
  int64_t half = (int64_t)(tot / 2);
  uint32_t n50;
  for (n50 = 0; n50 < top; ++n50) { half -= (uint32_t)(n50 * bins[n50]); if (half < 0) break; }
  o += snprintf(buf + o, (size_t)(n - o), " total count %llu\nMS average depth %.1f N50 depth %u",
                (unsigned long long)tot, tot / (double)sum, n50);
  if (copy[0] < max)
    o += snprintf(buf + o, (size_t)(n - o), " copy0 %u copy1 %u copy2 %u copyM %u", copy[0], copy[1], copy[2], copy[3]);
  o += snprintf(buf + o, (size_t)(n - o), "\n");
  return o;
}

extern "C" int modgpuModsetImport(ModgpuModset *ms, const uint64_t *value, const uint16_t *depth,
                                  const uint8_t *info, uint64_t n)
{
  int rc = ensure_numbered(ms);
  if (rc) return rc;
  if (!n) return MODGPU_OK;
  cudaStream_t st = ms->stream;
  if ((rc = ms->expo.ensure(n * 11 + 64))) return rc;
  uint64_t *dV = (uint64_t *)ms->expo.p;
  uint16_t *dD = (uint16_t *)(dV + n);
  uint8_t *dI = (uint8_t *)(dD + n);
  MG_CUDA(cudaMemcpyAsync(dV, value, n * 8, cudaMemcpyHostToDevice, st));
  if (depth) MG_CUDA(cudaMemcpyAsync(dD, depth, n * 2, cudaMemcpyHostToDevice, st));
  if (info) MG_CUDA(cudaMemcpyAsync(dI, info, n, cudaMemcpyHostToDevice, st));
  { ProfScope p(ms, MODGPU_T_OTHER, 1);
    if ((rc = modgpuTableImport(ms->table, dV, depth ? dD : nullptr, info ? dI : nullptr, n, st))) return rc;
  }
  if (modgpuTableEntries(ms->table, st) == 0xFFFFFFFFFFFFFFFFull) return MODGPU_EFULL;
  return MODGPU_OK;
}
