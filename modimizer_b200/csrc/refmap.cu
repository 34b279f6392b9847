// refmap.cu - the modmap side of the hot path on the device.
//
//   modgpuReferenceBuild == referenceFastaRead's insert loop + classification
//                           + modsetPack + referencePack (reference
//                           modmap.c:93-134, 74-91): per selected reference
//                           k-mer the triple (index, offset, id), per index the
//                           multiplicity, loc[] = exclusive prefix sum of the
//                           multiplicities, rev[] = occurrences grouped by index.
//   modgpuReferenceQuery == the seed loop of queryProcess (modmap.c:196-231):
//                           lookup of every read modimizer, the Q-line
//                           counters and the reference hits of copy-1/2 seeds.
//
// Index numbering is the reference's (first occurrence, modset.c:57), so the
// arrays are identical to the reference's, not merely isomorphic.
// rev[] is the reference's counting sort of the occurrence list by index (modmap.c:88-90), done here as a
// stable least-significant-digit radix sort of (index, ordinal) pairs: per pass a per-tile digit histogram,
// one exclusive scan over (digit, tile), and a scatter that ranks equal digits in input order.
#include <vector>
#include <string.h>
#include "mg_device.cuh"
#include "mg_scan.cuh"

struct ModgpuModset;
int mg_modset_select_chunk(ModgpuModset *ms, const uint8_t *d_bases, const uint64_t *d_offs,
                           uint64_t nSeq, uint64_t nBases, int isAscii, bool wantPos, int extraFlags,
                           uint64_t *nSelected);
int mg_table_insert_dev(ModgpuTable *t, const uint64_t *d_kmers, const uint64_t *d_n, uint64_t nMax,
                        uint32_t *d_slot, int exactOrder, cudaStream_t st);
int mg_table_lookup_dev(const ModgpuTable *t, const uint64_t *d_kmers, const uint64_t *d_n, uint64_t nMax,
                        uint32_t *d_out, cudaStream_t st);
uint64_t mg_table_numbered(const ModgpuTable *t);
int mg_modset_classify(ModgpuModset *ms, int mode, int c1, int c2, int cM, uint32_t classCounts[4]);
int mg_modset_ensure_numbered(ModgpuModset *ms);
void mg_modset_mark(ModgpuModset *ms, bool dirty, bool depthIsZero);
cudaStream_t mg_modset_stream(ModgpuModset *ms);
// scratch of the modset object (api.cu)
void *mg_modset_kmers(ModgpuModset *ms);
void *mg_modset_gpos(ModgpuModset *ms);
// double-buffered staging of host batches (api.cu): the copy of chunk c+1 runs behind the kernels of chunk c
struct MgFeed;
MgFeed *mg_feed_begin(ModgpuModset *ms, const char *bases, const uint64_t *offs, uint64_t nSeq, uint64_t limit, size_t *nChunks);
int mg_feed_chunk(ModgpuModset *ms, MgFeed *f, size_t c, const uint8_t **d_bases, const uint64_t **d_offs, uint64_t *r0, uint64_t *r1);
int mg_feed_release(ModgpuModset *ms, size_t c);
void mg_feed_end(ModgpuModset *ms, MgFeed *f);

static const uint64_t MG_REF_CHUNK = 1ull << 28;       // bases per staged chunk (whole sequences; a longer sequence is its own chunk)
static const uint64_t MG_QUERY_CHUNK = 1ull << 27;
static const uint32_t MG_REF_CAP = 1u << 26;      // modmap.c:363: referenceCreate(ms, 1 << 26)

struct RBuf {
  void *p = nullptr; size_t cap = 0;
  int ensure(size_t bytes, cudaStream_t st, bool keep = false)
  {
    if (bytes <= cap) return MODGPU_OK;
    void *q = nullptr;
    size_t want = bytes + bytes / 4 + 256;
    MG_CUDA(cudaMalloc(&q, want));
    if (p && keep) MG_CUDA(cudaMemcpyAsync(q, p, cap, cudaMemcpyDeviceToDevice, st));
    if (p) { MG_CUDA(cudaStreamSynchronize(st)); cudaFree(p); }
    p = q; cap = want;
    return MODGPU_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct ModgpuReference {
  ModgpuModset *ms = nullptr;
  uint32_t max = 0;                      // number of reference hits (ref->max)
  RBuf index, offset, id;                // per hit                      modmap.c:38-41
  RBuf depth, loc;                       // per modset index, max+1      modmap.c:42,44
  RBuf rev;                              // per hit, grouped by index    modmap.c:43
  RBuf tmp, tmp2, sortTmp, slot, misc;
  void *hPinned = nullptr;
};

// ------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(256) add_const_kernel(uint32_t *a, uint64_t n, uint32_t c)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a[i] += c;
}

// loc[i] = sum of depth[0..i-1]  (modmap.c:84-86; depth[0] = 0)
struct LocScan {
  const uint32_t *depth; uint32_t *loc;
  __device__ uint32_t value(uint64_t i) const { return depth[i]; }
  __device__ void emit(uint64_t i, uint32_t prefix, uint32_t) const { loc[i] = prefix; }
};

// per seed: counters of the Q line and the reference hits of the -v lines
__global__ void __launch_bounds__(256) query_resolve_kernel(const uint32_t *__restrict__ aux, uint64_t n,
                                                            const uint32_t *__restrict__ readId,
                                                            const uint32_t *__restrict__ loc, const uint32_t *__restrict__ rev,
                                                            const uint32_t *__restrict__ refId, const uint32_t *__restrict__ refOffset,
                                                            uint32_t *__restrict__ seedIndex, uint32_t *__restrict__ hitId,
                                                            uint32_t *__restrict__ hitOffset, int32_t *counters)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    { uint32_t a = aux[i];
      uint32_t ix = a >> 2, cls = a & 3u;
      seedIndex[i] = ix;
      uint32_t h0 = 0xFFFFFFFFu, o0 = 0xFFFFFFFFu, h1 = 0xFFFFFFFFu, o1 = 0xFFFFFFFFu;
      if (!ix) atomicAdd(&counters[4 * (uint64_t)readId[i]], 1);               // miss         modmap.c:207
      else
        { if (cls) atomicAdd(&counters[4 * (uint64_t)readId[i] + cls], 1);     // ++copy[msCopy], modmap.c:206
          if (cls != 3)                                                        // modmap.c:217
            { uint32_t l = rev[loc[ix]];                                       // modmap.c:219
              h0 = refId[l]; o0 = refOffset[l];
              if (cls != 1) { uint32_t l2 = rev[loc[ix] + 1]; h1 = refId[l2]; o1 = refOffset[l2]; }   // modmap.c:226
            }
        }
      hitId[2 * i] = h0; hitId[2 * i + 1] = h1;
      hitOffset[2 * i] = o0; hitOffset[2 * i + 1] = o1;
    }
}

// seedOff[r] = number of seeds whose read id is < r (ids are non-decreasing)
__global__ void __launch_bounds__(256) seed_offsets_kernel(const uint32_t *__restrict__ readId, uint64_t n,
                                                           uint64_t nSeq, uint64_t base, uint64_t *seedOff)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= nSeq; r += stride)
    { uint64_t lo = 0, hi = n;                       // first i with readId[i] >= r
      while (lo < hi)
        { uint64_t mid = (lo + hi) >> 1;
          if (readId[mid] < r) lo = mid + 1; else hi = mid;
        }
      seedOff[r] = base + lo;
    }
}

// ---- stable LSD radix sort of (key, ordinal) pairs, 8 bits per pass --------------------------------------
// A tile is 4096 consecutive elements; warp w of the tile's block owns elements [512 w, 512 w + 512) and walks them
// 32 at a time, so "input order" inside a tile is (warp, iteration, lane).
#define MG_RS_TILE 4096
#define MG_RS_ITERS 16

__global__ void __launch_bounds__(256) rs_hist_kernel(const uint32_t *__restrict__ keys, uint64_t n, int shift,
                                                      uint32_t *__restrict__ hist, uint32_t nTiles)
{
  __shared__ uint32_t sH[256];
  sH[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * MG_RS_TILE;
#pragma unroll 4
  for (int r = 0; r < MG_RS_ITERS; ++r)
    { const uint64_t i = base + (uint64_t)r * 256 + threadIdx.x;
      if (i < n) atomicAdd(&sH[(keys[i] >> shift) & 255u], 1u);
    }
  __syncthreads();
  hist[(uint64_t)threadIdx.x * nTiles + blockIdx.x] = sH[threadIdx.x];        // digit-major: the scan runs over (digit, tile)
}

struct RsScan {                 // exclusive scan of the (digit, tile) counts in place
  uint32_t *h;
  __device__ uint32_t value(uint64_t i) const { return h[i]; }
  __device__ void emit(uint64_t i, uint32_t prefix, uint32_t) const { h[i] = prefix; }
};

// valsIn == nullptr: the ordinals themselves (first pass); keysOut == nullptr: the sorted keys are not wanted (last pass)
__global__ void __launch_bounds__(256) rs_scatter_kernel(const uint32_t *__restrict__ keysIn, const uint32_t *__restrict__ valsIn,
                                                         uint64_t n, int shift, const uint32_t *__restrict__ offsets, uint32_t nTiles,
                                                         uint32_t *__restrict__ keysOut, uint32_t *__restrict__ valsOut)
{
  __shared__ uint32_t sCnt[8][256];                     // per warp: elements of each digit seen so far, then their base
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (uint32_t i = tid; i < 8 * 256; i += 256) (&sCnt[0][0])[i] = 0;
  __syncthreads();
  const uint64_t first = (uint64_t)blockIdx.x * MG_RS_TILE + (uint64_t)warp * (MG_RS_TILE / 8) + lane;
  uint32_t key[MG_RS_ITERS], rank[MG_RS_ITERS];
#pragma unroll
  for (int it = 0; it < MG_RS_ITERS; ++it)
    { const uint64_t i = first + (uint64_t)it * 32;
      const bool valid = i < n;
      const uint32_t k = valid ? keysIn[i] : 0u;
      const uint32_t d = valid ? ((k >> shift) & 255u) : 256u;          // lanes past the end form their own group
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
      const uint32_t r = __popc(peers & ((1u << lane) - 1u));
      uint32_t old = 0;
      if (valid && r == 0) { old = sCnt[warp][d]; sCnt[warp][d] = old + __popc(peers); }   // one leader per digit: no conflict
      old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1);
      key[it] = k; rank[it] = old + r;
      __syncwarp();
    }
  __syncthreads();
  { uint32_t run = offsets[(uint64_t)tid * nTiles + blockIdx.x];          // thread d: where digit d of this tile starts
#pragma unroll
    for (int w = 0; w < 8; ++w) { const uint32_t c = sCnt[w][tid]; sCnt[w][tid] = run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < MG_RS_ITERS; ++it)
    { const uint64_t i = first + (uint64_t)it * 32;
      if (i < n)
        { const uint32_t dst = sCnt[warp][(key[it] >> shift) & 255u] + rank[it];
          if (keysOut) keysOut[dst] = key[it];
          valsOut[dst] = valsIn ? valsIn[i] : (uint32_t)i;
        }
    }
}

// out[j] = ordinal of the j-th pair in (key, ordinal) order; keys < 2^keyBits.  kA, kB, vA, vB: n words of scratch each;
// hist: 256 * tiles + scan scratch (see rs_scratch_words)
static uint64_t rs_scratch_words(uint64_t n)
{
  const uint64_t tiles = (n + MG_RS_TILE - 1) / MG_RS_TILE, h = 256 * tiles;
  return h + (h + MG_CP_CHUNK - 1) / MG_CP_CHUNK + 64;
}

static int rs_sort_ordinals(const uint32_t *dKeys, uint64_t n, int keyBits, uint32_t *kA, uint32_t *kB, uint32_t *vA, uint32_t *vB,
                            uint32_t *hist, uint32_t *out, cudaStream_t st)
{
  const uint32_t tiles = (uint32_t)((n + MG_RS_TILE - 1) / MG_RS_TILE);
  const uint64_t h = 256ull * tiles;
  uint32_t *scanScratch = hist + h + 16;
  unsigned long long *dTotal = (unsigned long long *)(hist + h);          // 8-byte aligned: h is a multiple of 256
  const int passes = keyBits <= 8 ? 1 : (keyBits + 7) / 8;
  const uint32_t *kin = dKeys, *vin = nullptr;
  for (int p = 0; p < passes; ++p)
    { const bool last = p == passes - 1;
      uint32_t *kout = last ? nullptr : ((p & 1) ? kB : kA), *vout = last ? out : ((p & 1) ? vB : vA);
      rs_hist_kernel<<<tiles, 256, 0, st>>>(kin, n, 8 * p, hist, tiles);
      MG_LAUNCH_CHECK("rs_hist");
      RsScan f; f.h = hist;
      int rc = mg_ordered_scan(f, h, scanScratch, dTotal, st);
      if (rc) return rc;
      rs_scatter_kernel<<<tiles, 256, 0, st>>>(kin, vin, n, 8 * p, hist, tiles, kout, vout);
      MG_LAUNCH_CHECK("rs_scatter");
      kin = kout; vin = vout;
    }
  return MODGPU_OK;
}

static unsigned rgrid(uint64_t n)
{
  uint64_t blocks = (n + 255) / 256;
  uint64_t maxBlocks = (uint64_t)mg_num_sms() * 16;
  if (blocks > maxBlocks) blocks = maxBlocks;
  if (!blocks) blocks = 1;
  return (unsigned)blocks;
}

// ------------------------------------------------------------------- build
extern "C" void modgpuReferenceDestroy(ModgpuReference *R)
{
  if (!R) return;
  RBuf *all[] = { &R->index, &R->offset, &R->id, &R->depth, &R->loc, &R->rev, &R->tmp, &R->tmp2, &R->sortTmp,
                  &R->slot, &R->misc };
  for (RBuf *b : all) b->release();
  if (R->hPinned) cudaFreeHost(R->hPinned);
  if (R->ms) modgpuModsetDestroy(R->ms);
  delete R;
}

extern "C" ModgpuReference *modgpuReferenceBuild(int bits, int k, int w, int seed, const char *bases,
                                                 const uint64_t *offs, uint64_t nSeq, int isAscii, uint32_t counts[4])
{
  ModgpuModset *ms = modgpuModsetCreate(bits, k, w, seed);
  if (!ms) return nullptr;
  ModgpuReference *R = new ModgpuReference();
  R->ms = ms;
  modgpuModsetSetExactOrder(ms, 1);                  // first-occurrence numbering, ordered selection
  cudaStream_t st = mg_modset_stream(ms);
  ModgpuTable *t = modgpuModsetTable(ms);
#define RB_FAIL() do { modgpuReferenceDestroy(R); return nullptr; } while (0)
  if (mg_check_cuda(cudaMallocHost(&R->hPinned, 256), "cudaMallocHost", __FILE__, __LINE__)) RB_FAIL();
  if (!offs || offs[0] != 0) { mg_set_error("offsets must start at 0"); RB_FAIL(); }
  for (uint64_t r = 0; r < nSeq; ++r)
    if (offs[r + 1] < offs[r] || offs[r + 1] - offs[r] > 0x7FFFFFFFull)
      { mg_set_error("bad length of reference sequence %llu", (unsigned long long)r); RB_FAIL(); }

  // the hit arrays are sized once from the expected density (they still grow if the genome is denser)
  { const uint64_t guess = offs[nSeq] / (uint64_t)(w > 0 ? w : 1), want = guess + guess / 8 + 4096;
    const uint64_t n0 = want < MG_REF_CAP ? want : MG_REF_CAP;
    if (R->index.ensure(n0 * 4, st) || R->offset.ensure(n0 * 4, st) || R->id.ensure(n0 * 4, st)) RB_FAIL();
  }
  size_t nChunks = 0;
  MgFeed *feed = mg_feed_begin(ms, bases, offs, nSeq, MG_REF_CHUNK, &nChunks);
  if (!feed) RB_FAIL();
#define RB_FAIL2() do { mg_feed_end(ms, feed); RB_FAIL(); } while (0)
  uint64_t nHits = 0;
  for (size_t c = 0; c < nChunks; ++c)
    { const uint8_t *dBases; const uint64_t *dOffs; uint64_t r0, r1;
      if (mg_feed_chunk(ms, feed, c, &dBases, &dOffs, &r0, &r1)) RB_FAIL2();
      const uint64_t nb = offs[r1] - offs[r0], ns = r1 - r0;
      if (nb >= (1ull << 32)) { mg_set_error("reference sequence group exceeds 2^32-1 bases"); RB_FAIL2(); }
      uint64_t n = 0;
      if (mg_modset_select_chunk(ms, dBases, dOffs, ns, nb, isAscii, true, MODGPU_SEL_ORDERED, &n)) RB_FAIL2();
      if (n)
        { if (nHits + n + 1 >= MG_REF_CAP)               // modmap.c:111: die ("reference size overflow")
            { mg_set_error("reference size overflow (%llu hits, capacity 2^26)", (unsigned long long)(nHits + n)); RB_FAIL2(); }
          if (R->slot.ensure(n * 4, st) || R->index.ensure((nHits + n) * 4, st, true) ||
              R->offset.ensure((nHits + n) * 4, st, true) || R->id.ensure((nHits + n) * 4, st, true))
            RB_FAIL2();
          uint32_t *dIndex = (uint32_t *)R->index.p + nHits, *dOffset = (uint32_t *)R->offset.p + nHits, *dId = (uint32_t *)R->id.p + nHits;
          // find-or-insert with the multiplicity in the slot count (++ref->depth[index], modmap.c:113)
          if (mg_table_insert_dev(t, (const uint64_t *)mg_modset_kmers(ms), nullptr, n, (uint32_t *)R->slot.p, 1, st)) RB_FAIL2();
          if (modgpuTableNumber(t, (const uint32_t *)R->slot.p, n, dIndex, st)) RB_FAIL2();
          if (modgpuLocate((const uint32_t *)mg_modset_gpos(ms), n, dOffs, ns, dId, dOffset, st)) RB_FAIL2();
          if (r0)
            { add_const_kernel<<<rgrid(n), 256, 0, st>>>(dId, n, (uint32_t)r0);
              if (mg_check_cuda(cudaGetLastError(), "add_const", __FILE__, __LINE__)) RB_FAIL2();
            }
          nHits += n;
        }
      if (mg_feed_release(ms, c)) RB_FAIL2();
    }
  mg_feed_end(ms, feed);
#undef RB_FAIL2
  if (modgpuTableEntries(t, st) == 0xFFFFFFFFFFFFFFFFull) RB_FAIL();
  R->max = (uint32_t)nHits;
  mg_modset_mark(ms, false, true);                   // numbered; ms->depth stays 0 in modmap

  // classes by exact multiplicity (modmap.c:125-129)
  uint32_t cls[4] = { 0, 0, 0, 0 };
  if (mg_modset_classify(ms, 2, 0, 0, 0, cls)) RB_FAIL();
  if (counts) { counts[0] = R->max; counts[1] = cls[1]; counts[2] = cls[2]; counts[3] = cls[3]; }

  // referencePack (modmap.c:74-91)
  const uint64_t m = mg_table_numbered(t) + 1;
  if (R->depth.ensure(m * 4, st) || R->loc.ensure(m * 4, st) || R->rev.ensure((nHits + 1) * 4, st) ||
      R->misc.ensure(((m + MG_CP_CHUNK - 1) / MG_CP_CHUNK + 16) * 4 + 64, st))
    RB_FAIL();
  if (mg_check_cuda(cudaMemsetAsync(R->depth.p, 0, 4, st), "memset", __FILE__, __LINE__)) RB_FAIL();
  if (modgpuTableExport(t, nullptr, nullptr, nullptr, (uint32_t *)R->depth.p + 1, st)) RB_FAIL();
  { LocScan f; f.depth = (const uint32_t *)R->depth.p; f.loc = (uint32_t *)R->loc.p;
    unsigned long long *dTotal = (unsigned long long *)R->misc.p;
    if (mg_ordered_scan(f, m, (uint32_t *)((char *)R->misc.p + 64), dTotal, st)) RB_FAIL();
  }
  if (nHits)
    { // stable sort of hit ordinals by modset index == the counting sort of modmap.c:88-90
      if (R->tmp.ensure(nHits * 16, st) || R->sortTmp.ensure(rs_scratch_words(nHits) * 4, st)) RB_FAIL();
      int endBit = 1;
      while (endBit < 32 && (m >> endBit)) ++endBit;
      uint32_t *w = (uint32_t *)R->tmp.p;
      if (rs_sort_ordinals((const uint32_t *)R->index.p, nHits, endBit, w, w + nHits, w + 2 * nHits, w + 3 * nHits,
                           (uint32_t *)R->sortTmp.p, (uint32_t *)R->rev.p, st))
        RB_FAIL();
    }
  if (mg_check_cuda(cudaStreamSynchronize(st), "sync", __FILE__, __LINE__)) RB_FAIL();
  R->tmp.release(); R->tmp2.release(); R->sortTmp.release(); R->slot.release();
#undef RB_FAIL
  return R;
}

extern "C" ModgpuModset *modgpuReferenceModset(ModgpuReference *R) { return R->ms; }
extern "C" uint32_t modgpuReferenceMax(ModgpuReference *R) { return R->max; }

extern "C" int modgpuReferenceExport(ModgpuReference *R, uint32_t *index, uint32_t *offset, uint32_t *id,
                                     uint32_t *depth, uint32_t *rev, uint32_t *loc)
{
  cudaStream_t st = mg_modset_stream(R->ms);
  const uint64_t n = R->max, m = mg_table_numbered(modgpuModsetTable(R->ms)) + 1;
  if (n)
    { if (index) MG_CUDA(cudaMemcpyAsync(index, R->index.p, n * 4, cudaMemcpyDeviceToHost, st));
      if (offset) MG_CUDA(cudaMemcpyAsync(offset, R->offset.p, n * 4, cudaMemcpyDeviceToHost, st));
      if (id) MG_CUDA(cudaMemcpyAsync(id, R->id.p, n * 4, cudaMemcpyDeviceToHost, st));
      if (rev) MG_CUDA(cudaMemcpyAsync(rev, R->rev.p, n * 4, cudaMemcpyDeviceToHost, st));
    }
  if (depth) MG_CUDA(cudaMemcpyAsync(depth, R->depth.p, m * 4, cudaMemcpyDeviceToHost, st));
  if (loc) MG_CUDA(cudaMemcpyAsync(loc, R->loc.p, m * 4, cudaMemcpyDeviceToHost, st));
  MG_CUDA(cudaStreamSynchronize(st));
  return MODGPU_OK;
}

// ------------------------------------------------------------------- query
extern "C" uint64_t modgpuReferenceQuery(ModgpuReference *R, const char *bases, const uint64_t *offs,
                                         uint64_t nSeq, int isAscii, uint64_t *seedOff,
                                         uint32_t *seedIndex, uint32_t *seedPos,
                                         uint32_t *hitId, uint32_t *hitOffset,
                                         int32_t *counters, uint64_t cap)
{
  const uint64_t FAIL = 0xFFFFFFFFFFFFFFFFull;
  ModgpuModset *ms = R->ms;
  cudaStream_t st = mg_modset_stream(ms);
  ModgpuTable *t = modgpuModsetTable(ms);
  if (!offs || offs[0] != 0) { mg_set_error("offsets must start at 0"); return FAIL; }
  for (uint64_t r = 0; r < nSeq; ++r)
    if (offs[r + 1] < offs[r] || offs[r + 1] - offs[r] > 0x7FFFFFFFull)
      { mg_set_error("bad length of query sequence %llu", (unsigned long long)r); return FAIL; }
  size_t nChunks = 0;
  MgFeed *feed = mg_feed_begin(ms, bases, offs, nSeq, MG_QUERY_CHUNK, &nChunks);
  if (!feed) return FAIL;
#define RQ_FAIL() do { mg_feed_end(ms, feed); return FAIL; } while (0)
  uint64_t total = 0;
  seedOff[0] = 0;
  for (size_t c = 0; c < nChunks; ++c)
    { // the reads of chunk c+1 cross PCIe while chunk c is looked up and its seeds travel back
      const uint8_t *dBases; const uint64_t *dOffs; uint64_t r0, r1;
      if (mg_feed_chunk(ms, feed, c, &dBases, &dOffs, &r0, &r1)) RQ_FAIL();
      const uint64_t nb = offs[r1] - offs[r0], ns = r1 - r0;
      if (nb >= (1ull << 32)) { mg_set_error("query sequence group exceeds 2^32-1 bases"); RQ_FAIL(); }
      uint64_t n = 0;
      if (mg_modset_select_chunk(ms, dBases, dOffs, ns, nb, isAscii, true, MODGPU_SEL_ORDERED, &n)) RQ_FAIL();
      // device scratch: aux, readId, pos, seedIndex, hitId[2], hitOffset[2] per seed; counters, seedOff per read
      const size_t perSeed = 4 * 8, need = n * perSeed + (ns + 1) * (16 + 8) + 256;
      if (R->tmp.ensure(need, st)) RQ_FAIL();
      uint32_t *dAux = (uint32_t *)R->tmp.p, *dRead = dAux + n, *dPos = dRead + n, *dIdx = dPos + n;
      uint32_t *dHitId = dIdx + n, *dHitOff = dHitId + 2 * n;
      uint64_t *dSeedOff = (uint64_t *)(((uintptr_t)(dHitOff + 2 * n) + 15) & ~(uintptr_t)15);
      int32_t *dCtr = (int32_t *)(dSeedOff + ns + 1);
      if (mg_check_cuda(cudaMemsetAsync(dCtr, 0, ns * 16, st), "memset", __FILE__, __LINE__)) RQ_FAIL();
      if (n)
        { if (mg_table_lookup_dev(t, (const uint64_t *)mg_modset_kmers(ms), nullptr, n, dAux, st)) RQ_FAIL();
          if (modgpuLocate((const uint32_t *)mg_modset_gpos(ms), n, dOffs, ns, dRead, dPos, st)) RQ_FAIL();
          query_resolve_kernel<<<rgrid(n), 256, 0, st>>>(dAux, n, dRead, (const uint32_t *)R->loc.p, (const uint32_t *)R->rev.p,
                                                         (const uint32_t *)R->id.p, (const uint32_t *)R->offset.p, dIdx, dHitId, dHitOff, dCtr);
          if (mg_check_cuda(cudaGetLastError(), "query_resolve", __FILE__, __LINE__)) RQ_FAIL();
        }
      seed_offsets_kernel<<<rgrid(ns + 1), 256, 0, st>>>(dRead, n, ns, total, dSeedOff);
      if (mg_check_cuda(cudaGetLastError(), "seed_offsets", __FILE__, __LINE__)) RQ_FAIL();
      if (mg_feed_release(ms, c)) RQ_FAIL();             // the chunk's bases are no longer needed
      // results back to the caller (clipped to cap)
      uint64_t room = total < cap ? cap - total : 0, take = n < room ? n : room;
      if (take)
        { if (mg_check_cuda(cudaMemcpyAsync(seedIndex + total, dIdx, take * 4, cudaMemcpyDeviceToHost, st), "d2h", __FILE__, __LINE__) ||
              mg_check_cuda(cudaMemcpyAsync(seedPos + total, dPos, take * 4, cudaMemcpyDeviceToHost, st), "d2h", __FILE__, __LINE__) ||
              mg_check_cuda(cudaMemcpyAsync(hitId + 2 * total, dHitId, take * 8, cudaMemcpyDeviceToHost, st), "d2h", __FILE__, __LINE__) ||
              mg_check_cuda(cudaMemcpyAsync(hitOffset + 2 * total, dHitOff, take * 8, cudaMemcpyDeviceToHost, st), "d2h", __FILE__, __LINE__))
            RQ_FAIL();
        }
      if (mg_check_cuda(cudaMemcpyAsync(seedOff + r0, dSeedOff, (ns + 1) * 8, cudaMemcpyDeviceToHost, st), "d2h", __FILE__, __LINE__) ||
          mg_check_cuda(cudaMemcpyAsync(counters + 4 * r0, dCtr, ns * 16, cudaMemcpyDeviceToHost, st), "d2h", __FILE__, __LINE__) ||
          mg_check_cuda(cudaStreamSynchronize(st), "sync", __FILE__, __LINE__))
        RQ_FAIL();
      total += n;
    }
  mg_feed_end(ms, feed);
#undef RQ_FAIL
  return total;
}
