// mg_api.h - private declarations shared by the host-level translation units of
// libmodgpu (api.cu, refmap.cu, setops.cu): the ModgpuModset object, its scratch
// buffers, per-kernel timing scopes, and the internal entry points of the
// kernel files.  Not part of the public ABI (include/modgpu.h).
#pragma once
#include <vector>
#include "mg_device.cuh"

MgKHasher mg_khasher_from(const ModgpuHasher *h);
int mg_ends_sparse(const uint64_t *d_offs, uint64_t nSeq, uint32_t *d_ends, uint8_t *d_tileFlags, int set, cudaStream_t st);
bool mg_table_clear_pending(const ModgpuTable *t);
void mg_table_bulk_abort(ModgpuTable *t, bool wasPending);
int mg_select_launches(const ModgpuHasher *h, int flags);      // kernels one hash/select call launches
int mg_table_insert_dev(ModgpuTable *t, const uint64_t *d_kmers, const uint64_t *d_n, uint64_t nMax,
                        uint32_t *d_slot, int exactOrder, cudaStream_t st);
int mg_table_lookup_dev(const ModgpuTable *t, const uint64_t *d_kmers, const uint64_t *d_n, uint64_t nMax,
                        uint32_t *d_out, cudaStream_t st);
uint64_t mg_table_numbered(const ModgpuTable *t);
int mg_table_insert_bulk(ModgpuTable *t, const uint64_t *d_kmers, uint64_t n, cudaStream_t st);
struct MgBulk { uint32_t slotBits, regionBits, nRegions, cap; uint32_t *cursors; uint64_t *buckets, *overflow; uint64_t overflowCap; uint64_t expected; };
int mg_table_bulk_begin(ModgpuTable *t, uint64_t expectedN, uint64_t maxN, MgBulk *b, cudaStream_t st);
const uint32_t *mg_table_bulk_overflow_count(const ModgpuTable *t);
int mg_table_bulk_finish(ModgpuTable *t, const MgBulk *b, cudaStream_t st);
// deferred build: buckets kept open over several chunks (table.cu)
int mg_table_bulk_open(ModgpuTable *t, uint64_t expected, uint32_t accum, MgBulk *b, cudaStream_t st);
void mg_table_bulk_commit(ModgpuTable *t, uint64_t expected, uint64_t ovfSeen);
int mg_table_bulk_rollback(ModgpuTable *t, cudaStream_t st);
int mg_table_bulk_close(ModgpuTable *t, cudaStream_t st);
bool mg_table_bulk_is_open(const ModgpuTable *t);
int mg_hash_select_peer(const ModgpuHasher *h, const uint64_t *d_packed, const uint32_t *d_ends, uint64_t nBases,
                        uint64_t *d_count, void *d_workspace, int flags, uint32_t slotBits, uint32_t regionBits,
                        uint32_t nOwners, uint32_t bucketCap, uint32_t *d_cursors, uint64_t *d_buckets,
                        uint64_t *d_overflow, uint64_t overflowCap, uint32_t *d_ovfCounts,
                        const uint8_t *d_raw, int rawAscii, const uint8_t *d_tileFlags, cudaStream_t st);
int mg_table_build_from_buckets(ModgpuTable *t, const uint64_t *d_buckets, const uint32_t *d_cursors, uint32_t cap, uint32_t nSrc,
                                const uint64_t *d_overflow, uint64_t overflowCap, const uint32_t *d_ovfCounts, cudaStream_t st);
int mg_table_build_from_peers(ModgpuTable *t, const uint64_t *const *d_buckets, const uint32_t *d_cursors, uint32_t cap, uint32_t nSrc,
                              const uint64_t *const *d_overflow, uint64_t overflowCap, const uint32_t *d_ovfCounts, cudaStream_t st);
uint32_t mg_table_regions(const ModgpuTable *t);
uint32_t mg_table_slot_bits(const ModgpuTable *t);
int mg_hash_select_owners(const ModgpuHasher *h, const uint64_t *d_packed, const uint32_t *d_ends, uint64_t nBases,
                          void *d_workspace, int flags, uint32_t nOwners, uint32_t *d_cursors, uint64_t *d_buf,
                          uint64_t ownerCap, cudaStream_t st);
int mg_table_insert_segments(ModgpuTable *t, const uint64_t *d_segs, uint32_t nSegs, uint64_t segCap,
                             const uint32_t *d_counts, uint64_t expectedN, cudaStream_t st);
int mg_hash_select_scatter(const ModgpuHasher *h, const uint64_t *d_packed, const uint32_t *d_ends, uint64_t nBases,
                           uint64_t *d_count, void *d_workspace, int flags, uint32_t slotBits, uint32_t regionBits,
                           uint32_t bucketCap, uint32_t *d_cursors, uint64_t *d_buckets, uint64_t *d_overflow,
                           uint64_t overflowCap, const uint8_t *d_raw, int rawAscii, const uint8_t *d_tileFlags, cudaStream_t st);
uint64_t mg_table_bulk_threshold(const ModgpuTable *t);
int mg_slot_partition(const uint64_t *d_kmers, uint64_t n, uint32_t slotBits, uint32_t bucketBits,
                      uint64_t *d_out, uint64_t *d_scratch, cudaStream_t st);
int mg_table_classify(ModgpuTable *t, int mode, int c1, int c2, int cM, int zeroDepth, uint32_t *d_classCounts, cudaStream_t st);

// bases per pipelined chunk when the batch comes from host memory / is resident
static const uint64_t MG_HOST_CHUNK = 1ull << 28;
static const uint64_t MG_DEV_CHUNK = (1ull << 32) - (1ull << 20);   // global offsets are 32-bit

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes)
  {
    if (bytes <= cap) return MODGPU_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    MG_CUDA(cudaMalloc(&p, want));
    cap = want;
    return MODGPU_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinBuf {
  void *p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes)
  {
    if (bytes <= cap) return MODGPU_OK;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    MG_CUDA(cudaMallocHost(&p, want));
    cap = want;
    return MODGPU_OK;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct TimedSpan { cudaEvent_t a, b; int cat; };

struct ModgpuModset {
  ModgpuHasher hasher;
  ModgpuTable *table = nullptr;
  int bits = 0;
  int device = 0;                   // the CUDA device the set lives on (threads other than the creator select it first)
  cudaStream_t stream = nullptr, copyStream = nullptr;
  bool ownStream = false;
  int selFlags = 0;
  int exactOrder = 0;
  bool depthIsZero = false;         // modmap-built sets keep ms->depth at 0
  bool dirty = false;               // entries inserted since the last numbering
  DevBuf bases[2], offs[2], pk[2], packed, ends, kmers, kmers2, gpos, slot, work, misc, expo;
  size_t endsCleanCap = 0;          // ms->ends is all zero over this capacity (0: unknown / flags of a batch still set)
  uint8_t *tileFlagsAt = nullptr;
  uint8_t *tileFlags = nullptr;     // the current batch's per-tile "a sequence ends here" bytes (tail of ms->ends), or null:
                                    // many short sequences - the count kernel stages the per-base flags with every tile
  int regionBits = -1;              // -1 auto: partition inserts by table region when the table exceeds L2
  PinBuf hOffs[2], hMisc;
  cudaEvent_t evCopied[2] = { nullptr, nullptr }, evFree[2] = { nullptr, nullptr };
  uint64_t totalHashes = 0;
  uint32_t accumulate = 1;          // > 1: deferred build, up to this many chunks share one region build
  // profiling
  bool profile = false;
  std::vector<TimedSpan> spans;
  std::vector<cudaEvent_t> evPool;
  double ms[MODGPU_T_N] = { 0, 0, 0, 0 };
  uint64_t launches[MODGPU_T_N] = { 0, 0, 0, 0 };
};

// ------------------------------------------------------------- profiling --
static inline cudaEvent_t prof_event(ModgpuModset *ms)
{
  if (!ms->evPool.empty()) { cudaEvent_t e = ms->evPool.back(); ms->evPool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

struct ProfScope {
  ModgpuModset *ms; TimedSpan s; bool on;
  ProfScope(ModgpuModset *m, int cat, int nLaunch = 1) : ms(m), on(m->profile)
  {
    ms->launches[cat] += (uint64_t)nLaunch;
    if (!on) return;
    s.cat = cat; s.a = prof_event(ms); s.b = prof_event(ms);
    cudaEventRecord(s.a, ms->stream);
  }
  ~ProfScope()
  {
    if (!on) return;
    cudaEventRecord(s.b, ms->stream);
    ms->spans.push_back(s);
  }
};

static inline void prof_collect(ModgpuModset *ms)
{
  for (TimedSpan &s : ms->spans)
    { float t = 0.f;
      if (cudaEventSynchronize(s.b) == cudaSuccess && cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess) ms->ms[s.cat] += t;
      ms->evPool.push_back(s.a); ms->evPool.push_back(s.b);
    }
  ms->spans.clear();
}


// api.cu
int mg_modset_select_chunk(ModgpuModset *ms, const uint8_t *d_bases, const uint64_t *d_offs,
                           uint64_t nSeq, uint64_t nBases, int isAscii, bool wantPos, int extraFlags,
                           uint64_t *nSelected);
int mg_modset_ensure_numbered(ModgpuModset *ms);
int mg_modset_classify(ModgpuModset *ms, int mode, int c1, int c2, int cM, uint32_t classCounts[4]);
