// setops.cu - whole-set operations and file formats around the modset table.
//
//   modgpuModsetPrune    == modsetDepthPrune                  reference modset.c:64-77
//   modgpuModsetMerge    == modsetMerge                       reference modset.c:106-128
//   modgpuModsetWriteMod == modsetWrite ("MSHSTv2")            reference modset.c:79-88, seqhash.c:41-44
//   modgpuModsetReadMod  == modsetRead                        reference modset.c:90-104
//   modgpuModsetReadset  == the hot loop of readsetFileRead   reference modasm.c:151-191
//
// SURVEY 8(f) "next" rows 1, 3, 4.  Set arithmetic runs on the device (ordered
// compaction, find-or-insert with first-occurrence numbering, the reference's
// index[] layout: hostsync.cu); the file functions are host I/O: the only
// host-side work is the byte layout of the files.
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <vector>
#include <zlib.h>
#include "mg_api.h"
#include "mg_table.cuh"
#include "mg_scan.cuh"

void mg_table_counters(ModgpuTable *t, unsigned long long **entries, uint32_t **error);
uint8_t *mg_table_info_hi(ModgpuTable *t);

static unsigned sgrid(uint64_t n)
{
  uint64_t blocks = (n + 255) / 256;
  uint64_t maxBlocks = (uint64_t)mg_num_sms() * 16;
  if (blocks > maxBlocks) blocks = maxBlocks;
  if (!blocks) blocks = 1;
  return (unsigned)blocks;
}

MgSlot *mg_table_slots_on(ModgpuTable *t, cudaStream_t st);
static MgSlot *table_slots(ModgpuModset *ms) { return mg_table_slots_on(ms->table, ms->stream); }

// dense export of a numbered set into one device buffer: value[n] | depth[n] | info[n]
static int export_dense(ModgpuModset *ms, DevBuf &buf, uint64_t n, uint64_t **dV, uint16_t **dD, uint8_t **dI)
{
  int rc = buf.ensure(n * 11 + 64);
  if (rc) return rc;
  *dV = (uint64_t *)buf.p; *dD = (uint16_t *)(*dV + n); *dI = (uint8_t *)(*dD + n);
  if ((rc = modgpuTableExport(ms->table, *dV, *dD, *dI, nullptr, ms->stream))) return rc;
  if (ms->depthIsZero) MG_CUDA(cudaMemsetAsync(*dD, 0, n * 2, ms->stream));
  return MODGPU_OK;
}

// ------------------------------------------------------------------ prune --
struct PruneScan {
  const uint64_t *v; const uint16_t *d; const uint8_t *i;
  uint64_t *ov; uint16_t *od; uint8_t *oi;
  int min, max;
  __device__ uint32_t value(uint64_t j) const
  { int dep = (int)d[j]; return (dep >= min && (!max || dep < max)) ? 1u : 0u; }       // modset.c:70
  __device__ void emit(uint64_t j, uint32_t rank, uint32_t keep) const
  { if (keep) { ov[rank] = v[j]; od[rank] = d[j]; oi[rank] = i[j]; } }
};

extern "C" int modgpuModsetPrune(ModgpuModset *ms, int min, int max)
{
  int rc = mg_modset_ensure_numbered(ms);
  if (rc) return rc;
  const uint64_t n = mg_table_numbered(ms->table);
  if (!n) return MODGPU_OK;
  cudaStream_t st = ms->stream;
  uint64_t *dV; uint16_t *dD; uint8_t *dI;
  if ((rc = export_dense(ms, ms->expo, n, &dV, &dD, &dI))) return rc;
  if ((rc = ms->kmers2.ensure(n * 11 + 64)) || (rc = ms->work.ensure((n / MG_CP_CHUNK + 16) * 4 + 64))) return rc;
  PruneScan f;
  f.v = dV; f.d = dD; f.i = dI; f.min = min; f.max = max;
  f.ov = (uint64_t *)ms->kmers2.p; f.od = (uint16_t *)(f.ov + n); f.oi = (uint8_t *)(f.od + n);
  unsigned long long *dTotal = (unsigned long long *)ms->work.p;
  { ProfScope p(ms, MODGPU_T_OTHER, 3);
    if ((rc = mg_ordered_scan(f, n, (uint32_t *)((char *)ms->work.p + 64), dTotal, st))) return rc;
  }
  volatile uint64_t *h = (volatile uint64_t *)ms->hMisc.p;
  MG_CUDA(cudaMemcpyAsync((void *)h, dTotal, 8, cudaMemcpyDeviceToHost, st));
  MG_CUDA(cudaStreamSynchronize(st));
  const uint64_t kept = h[0];
  // survivors are re-inserted in their old index order: new indices 1..kept (modset.c:69-74)
  if ((rc = modgpuTableClear(ms->table, st))) return rc;
  ms->dirty = false;
  if (kept && (rc = modgpuTableImport(ms->table, f.ov, ms->depthIsZero ? nullptr : f.od, f.oi, kept, st))) return rc;
  if (modgpuTableEntries(ms->table, st) == 0xFFFFFFFFFFFFFFFFull) return MODGPU_EFULL;
  return MODGPU_OK;
}

// ------------------------------------------------------------------ merge --
__global__ void __launch_bounds__(256) merge_insert_kernel(MgSlot *slots, uint32_t slotBits, const uint64_t *__restrict__ v2,
                                                           const uint16_t *__restrict__ d2, uint64_t n, uint32_t *slotOf,
                                                           unsigned long long *entries, uint32_t *error)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint32_t fresh = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    { bool isNew;
      uint64_t s = probe_insert(slots, slotBits, v2[i], &isNew);
      if (s == 0xFFFFFFFFFFFFFFFFull) { atomicExch(error, 1u); slotOf[i] = 0xFFFFFFFFu; continue; }
      fresh += isNew ? 1u : 0u;
      atomicAdd(&slots[s].count, (uint32_t)d2[i]);                 // depths add, clamp to 65535 on export (modset.c:121-122)
      atomicMin(&slots[s].aux, MG_AUX_ORD + (uint32_t)i);          // new entries are numbered in ms2 index order (modset.c:120)
      slotOf[i] = (uint32_t)s;
    }
  fresh = mg_warp_sum(fresh);
  if (mg_lane() == 0 && fresh) atomicAdd(entries, (unsigned long long)fresh);
}

// copy numbers: c = min(c1 + c2, 3), info = (info & 3) | c      (modset.c:124-125): every entry of ms1 that ms2 touches
// loses its flags beyond the copy bits (`&= 0x3`), the others keep theirs
__global__ void __launch_bounds__(256) merge_copy_kernel(MgSlot *slots, const uint32_t *__restrict__ slotOf,
                                                         const uint8_t *__restrict__ i2, uint64_t n, uint8_t *__restrict__ infoHi)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    { uint32_t s = slotOf[i];
      if (s == 0xFFFFFFFFu) continue;
      uint32_t aux = slots[s].aux;
      uint32_t c1 = aux & 3u, c = c1 + (i2[i] & 3u);
      if (c > 3u) c = 3u;
      slots[s].aux = (aux & ~3u) | c1 | c;
      if (infoHi) infoHi[(aux >> 2) - 1] = 0;
    }
}

extern "C" int modgpuModsetMerge(ModgpuModset *a, ModgpuModset *b)
{
  if (a->hasher.w != b->hasher.w || a->hasher.k != b->hasher.k || a->hasher.factor1 != b->hasher.factor1) return 0;   // modset.c:111
  int rc;
  if ((rc = mg_modset_ensure_numbered(a)) || (rc = mg_modset_ensure_numbered(b))) return rc;
  const uint64_t n2 = mg_table_numbered(b->table);
  if (!n2) return 1;
  if (n2 >= (1ull << 30) - 2) { mg_set_error("modsetMerge: second set too large"); return MODGPU_EINVAL; }
  uint64_t *dV; uint16_t *dD; uint8_t *dI;
  if ((rc = export_dense(b, b->expo, n2, &dV, &dD, &dI))) return rc;
  MG_CUDA(cudaStreamSynchronize(b->stream));
  cudaStream_t st = a->stream;
  if ((rc = a->slot.ensure(n2 * 4))) return rc;
  uint32_t *dSlot = (uint32_t *)a->slot.p;
  unsigned long long *entries; uint32_t *error;
  mg_table_counters(a->table, &entries, &error);
  MgSlot *slots = table_slots(a);
  if (!slots) return MODGPU_ECUDA;
  { ProfScope p(a, MODGPU_T_INSERT, 1);
    merge_insert_kernel<<<sgrid(n2), 256, 0, st>>>(slots, mg_table_slot_bits(a->table), dV, dD, n2, dSlot, entries, error);
    MG_LAUNCH_CHECK("merge_insert");
  }
  { ProfScope p(a, MODGPU_T_OTHER, 4);
    if ((rc = modgpuTableNumber(a->table, dSlot, n2, nullptr, st))) return rc;
    merge_copy_kernel<<<sgrid(n2), 256, 0, st>>>(slots, dSlot, dI, n2, mg_table_info_hi(a->table));
    MG_LAUNCH_CHECK("merge_copy");
  }
  if (modgpuTableEntries(a->table, st) == 0xFFFFFFFFFFFFFFFFull) return MODGPU_EFULL;
  return 1;
}

// ------------------------------------------------------------- .mod files --
// Seqhash as the reference dumps it raw (seqhash.h:15-23, 80 bytes on LP64)
static void seqhash_bytes(const ModgpuHasher *h, unsigned char out[80])
{
  memset(out, 0, 80);
  int32_t shift2 = 2 * h->k;
  memcpy(out + 0, &h->seed, 4); memcpy(out + 4, &h->k, 4); memcpy(out + 8, &h->w, 4);
  memcpy(out + 16, &h->mask, 8); memcpy(out + 24, &h->shift1, 4); memcpy(out + 28, &shift2, 4);
  memcpy(out + 32, &h->factor1, 8); memcpy(out + 40, &h->factor2, 8);
  for (int i = 0; i < 4; ++i) { uint64_t p = ((uint64_t)(3 - i)) << (2 * (h->k - 1)); memcpy(out + 48 + 8 * i, &p, 8); }
}

struct OutFile {                   // plain or gzip'd, like fzopen(path, "w") (utils.c:108-127)
  FILE *f = nullptr; gzFile z = nullptr;
  bool open(const char *path, int gz) { if (gz) z = gzopen(path, "wb1"); else f = fopen(path, "wb"); return f || z; }
  bool write(const void *p, size_t n)
  {
    const char *c = (const char *)p;
    while (n)
      { size_t m = n > (1u << 30) ? (1u << 30) : n;
        if (z) { if (gzwrite(z, c, (unsigned)m) != (int)m) return false; }
        else if (fwrite(c, 1, m, f) != m) return false;
        c += m; n -= m;
      }
    return true;
  }
  void close() { if (z) gzclose(z); if (f) fclose(f); z = nullptr; f = nullptr; }
};

extern "C" int modgpuModsetWriteMod(ModgpuModset *ms, const char *path, int gzip)
{
  int rc = mg_modset_ensure_numbered(ms);
  if (rc) return rc;
  const uint64_t n = mg_table_numbered(ms->table);
  const uint32_t size = (uint32_t)(n + 1);
  std::vector<uint64_t> value(size, 0);
  std::vector<uint16_t> depth(size, 0);
  std::vector<uint8_t> info(size, 0);
  if (n && (rc = modgpuModsetExport(ms, value.data() + 1, depth.data() + 1, info.data() + 1))) return rc;
  // index[]: the reference's own open-addressing layout (entries inserted in index order, home slot hash & mask, odd
  // double-hashing stride: modset.c:48-57), built on the device (hostsync.cu) so that an unmodified `modutils -r` /
  // `modmap -r` can look k-mers up in a GPU-built set
  const int bits = ms->bits;
  const uint64_t tableSize = 1ull << bits;
  std::vector<uint32_t> index(tableSize, 0);
  if ((rc = modgpuModsetReferenceIndex(ms, index.data()))) return rc;
  OutFile out;
  if (!out.open(path, gzip)) { mg_set_error("failed to open mod file %s", path); return MODGPU_EINVAL; }
  unsigned char sh[80];
  seqhash_bytes(&ms->hasher, sh);
  bool ok = out.write("MSHSTv2", 8) && out.write(&bits, 4) && out.write(&size, 4) && out.write("SQHSHv2", 8) && out.write(sh, 80) &&
            out.write(index.data(), tableSize * 4) && out.write(value.data(), (size_t)size * 8) &&
            out.write(depth.data(), (size_t)size * 2) && out.write(info.data(), size);
  out.close();
  if (!ok) { mg_set_error("failed to write mod file %s", path); return MODGPU_EINVAL; }
  return MODGPU_OK;
}

static bool gz_read_all(gzFile z, void *p, size_t n)
{
  char *c = (char *)p;
  while (n)
    { unsigned m = n > (1u << 30) ? (1u << 30) : (unsigned)n;
      int r = gzread(z, c, m);
      if (r <= 0) return false;
      c += r; n -= (size_t)r;
    }
  return true;
}

extern "C" ModgpuModset *modgpuModsetCreateWithHasher(int bits, const ModgpuHasher *h);

extern "C" ModgpuModset *modgpuModsetReadMod(const char *path)
{
  gzFile z = gzopen(path, "rb");                       // reads gzip'd and plain files alike, as fzopen does
  if (!z) { mg_set_error("failed to open mod file %s", path); return nullptr; }
  char name[8]; int bits = 0; uint32_t size = 0; unsigned char sh[80];
  ModgpuModset *ms = nullptr;
  std::vector<uint64_t> value; std::vector<uint16_t> depth; std::vector<uint8_t> info;
  ModgpuHasher h;
  if (!gz_read_all(z, name, 8) || memcmp(name, "MSHSTv2", 8)) { mg_set_error("bad modset header in %s", path); goto fail; }   // modset.c:92-93
  if (!gz_read_all(z, &bits, 4) || !gz_read_all(z, &size, 4) || size < 1) { mg_set_error("failed to read bits/size"); goto fail; }
  if (bits < 20 || bits > 34) { mg_set_error("table bits %d must be between 20 and 34 (%s)", bits, path); goto fail; }         // modset.c:17
  if ((uint64_t)size >= (1ull << (bits - 2))) { mg_set_error("Modset size %u is too big for %d bits (%s)", size, bits, path); goto fail; }   // modset.c:24
  if (!gz_read_all(z, name, 8) || memcmp(name, "SQHSHv2", 8) || !gz_read_all(z, sh, 80)) { mg_set_error("seqhash read mismatch"); goto fail; }
  if (modgpuHasherFromSeqhash(&h, sh)) goto fail;
  { // index[] is the reference's host-side table: skipped, ours is rebuilt from value[]
    std::vector<char> skip(1 << 24);
    uint64_t left = (1ull << bits) * 4;
    while (left) { size_t m = left > skip.size() ? skip.size() : (size_t)left; if (!gz_read_all(z, skip.data(), m)) { mg_set_error("failed read index"); goto fail; } left -= m; }
  }
  value.resize(size); depth.resize(size); info.resize(size);
  if (!gz_read_all(z, value.data(), (size_t)size * 8) || !gz_read_all(z, depth.data(), (size_t)size * 2) || !gz_read_all(z, info.data(), size))
    { mg_set_error("failed to read value/depth/info"); goto fail; }
  gzclose(z); z = nullptr;
  ms = modgpuModsetCreateWithHasher(bits, &h);
  if (!ms) return nullptr;
  if (size > 1 && modgpuModsetImport(ms, value.data() + 1, depth.data() + 1, info.data() + 1, size - 1)) { modgpuModsetDestroy(ms); return nullptr; }
  return ms;
fail:
  if (z) gzclose(z);
  return nullptr;
}

// ---------------------------------------------------------------- readset --
// per seed: slot (or none) of the k-mer
__global__ void __launch_bounds__(256) lookup_slot_kernel(const MgSlot *slots, uint32_t slotBits, const uint64_t *__restrict__ kmers,
                                                          uint64_t n, uint32_t *__restrict__ slotOut)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    { uint64_t s = probe_find(slots, slotBits, kmers[i] & 0x3FFFFFFFFFFFFFFFull);
      slotOut[i] = (s == 0xFFFFFFFFFFFFFFFFull) ? 0xFFFFFFFFu : (uint32_t)s;
    }
}

__global__ void __launch_bounds__(256) zero_counts_kernel(MgSlot *slots, uint64_t nSlots)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nSlots; i += stride) slots[i].count = 0;
}

// found seeds are compacted in order; misses are tallied per read; depth is re-counted (modasm.c:170-176)
struct ReadsetScan {
  MgSlot *slots; const uint32_t *slotOf; const uint64_t *kmers; const uint32_t *readId; const uint32_t *pos;
  uint32_t *hit; uint32_t *hitRead; uint32_t *hitPos; int32_t *nMiss;
  __device__ uint32_t value(uint64_t i) const { return slotOf[i] != 0xFFFFFFFFu ? 1u : 0u; }
  __device__ void emit(uint64_t i, uint32_t rank, uint32_t found) const
  {
    if (!found) { atomicAdd(&nMiss[readId[i]], 1); return; }
    const uint32_t s = slotOf[i];
    const uint32_t index = slots[s].aux >> 2;
    hit[rank] = (kmers[i] >> 63) ? (index | 0x80000000u) : index;        // TOPBIT = forward (modasm.c:22,171)
    hitRead[rank] = readId[i]; hitPos[rank] = pos[i];
    atomicAdd(&slots[s].count, 1u);
  }
};

__global__ void __launch_bounds__(256) readset_dx_kernel(const uint32_t *__restrict__ hitRead, const uint32_t *__restrict__ hitPos,
                                                         uint64_t n, uint16_t *__restrict__ dx)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
    { uint32_t last = (j && hitRead[j - 1] == hitRead[j]) ? hitPos[j - 1] : 0u;   // lastPos starts at 0 for every read
      dx[j] = (uint16_t)(hitPos[j] - last);                                     // U16 truncation as in the reference
    }
}

__global__ void __launch_bounds__(256) first_geq_kernel(const uint32_t *__restrict__ sortedIds, uint64_t n, uint64_t nSeq,
                                                        uint64_t base, uint64_t *out)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= nSeq; r += stride)
    { uint64_t lo = 0, hi = n;
      while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (sortedIds[mid] < r) lo = mid + 1; else hi = mid; }
      out[r] = base + lo;
    }
}

extern "C" uint64_t modgpuModsetReadset(ModgpuModset *ms, const char *bases, const uint64_t *offs, uint64_t nSeq, int isAscii,
                                        int resetDepth, uint64_t *hitOff, uint32_t *hit, uint16_t *dx, int32_t *nMiss, uint64_t cap)
{
  const uint64_t FAIL = 0xFFFFFFFFFFFFFFFFull;
  if (mg_modset_ensure_numbered(ms)) return FAIL;
  cudaStream_t st = ms->stream;
  if (!offs || offs[0] != 0) { mg_set_error("offsets must start at 0"); return FAIL; }
  for (uint64_t r = 0; r < nSeq; ++r)
    if (offs[r + 1] < offs[r] || offs[r + 1] - offs[r] > 0x7FFFFFFFull) { mg_set_error("bad length of read %llu", (unsigned long long)r); return FAIL; }
  MgSlot *slots = table_slots(ms);
  if (!slots) return FAIL;
  const uint64_t nSlots = modgpuTableSlots(ms->table);
  if (resetDepth)
    { zero_counts_kernel<<<sgrid(nSlots), 256, 0, st>>>(slots, nSlots);      // memset (rs->ms->depth, 0, ..), modasm.c:158
      if (mg_check_cuda(cudaGetLastError(), "zero_counts", __FILE__, __LINE__)) return FAIL;
      ms->depthIsZero = false;
    }
  hitOff[0] = 0;
  uint64_t total = 0, r0 = 0;
  while (r0 < nSeq)
    { uint64_t r1 = r0 + 1;
      while (r1 < nSeq && offs[r1 + 1] - offs[r0] <= (1ull << 30)) ++r1;
      const uint64_t nb = offs[r1] - offs[r0], ns = r1 - r0;
      if (nb >= (1ull << 32)) { mg_set_error("read group exceeds 2^32-1 bases"); return FAIL; }
      if (ms->bases[0].ensure(nb + 64) || ms->offs[0].ensure((ns + 1) * 8) || ms->hOffs[0].ensure((ns + 1) * 8)) return FAIL;
      if (mg_check_cuda(cudaStreamSynchronize(st), "sync", __FILE__, __LINE__)) return FAIL;
      uint64_t *ho = (uint64_t *)ms->hOffs[0].p;
      for (uint64_t r = 0; r <= ns; ++r) ho[r] = offs[r0 + r] - offs[r0];
      if (nb && mg_check_cuda(cudaMemcpyAsync(ms->bases[0].p, bases + offs[r0], nb, cudaMemcpyHostToDevice, st), "h2d", __FILE__, __LINE__)) return FAIL;
      if (mg_check_cuda(cudaMemcpyAsync(ms->offs[0].p, ho, (ns + 1) * 8, cudaMemcpyHostToDevice, st), "h2d", __FILE__, __LINE__)) return FAIL;
      uint64_t n = 0;
      if (mg_modset_select_chunk(ms, (const uint8_t *)ms->bases[0].p, (const uint64_t *)ms->offs[0].p, ns, nb, isAscii, true,
                                 MODGPU_SEL_ORDERED | MODGPU_SEL_STRAND, &n))
        return FAIL;
      // scratch: slotOf, readId, pos (per seed) | hit, hitRead, hitPos (per hit) | dx | nMiss, hitOff (per read)
      const size_t need = n * 4 * 6 + n * 2 + 64 + (ns + 1) * 12 + 64 + (n / MG_CP_CHUNK + 16) * 4 + 256;
      if (ms->kmers2.ensure(need)) return FAIL;
      uint32_t *dSlot = (uint32_t *)ms->kmers2.p, *dRead = dSlot + n, *dPos = dRead + n, *dHit = dPos + n, *dHitRead = dHit + n, *dHitPos = dHitRead + n;
      uint16_t *dDx = (uint16_t *)(dHitPos + n);
      uint64_t *dOff = (uint64_t *)(((uintptr_t)(dDx + n) + 63) & ~(uintptr_t)63);
      int32_t *dMiss = (int32_t *)(dOff + ns + 1);
      unsigned long long *dTotal = (unsigned long long *)(((uintptr_t)(dMiss + ns) + 63) & ~(uintptr_t)63);
      uint32_t *dScr = (uint32_t *)(dTotal + 8);
      if (mg_check_cuda(cudaMemsetAsync(dMiss, 0, ns * 4, st), "memset", __FILE__, __LINE__)) return FAIL;
      uint64_t nHit = 0;
      if (n)
        { lookup_slot_kernel<<<sgrid(n), 256, 0, st>>>(slots, mg_table_slot_bits(ms->table), (const uint64_t *)ms->kmers.p, n, dSlot);
          if (mg_check_cuda(cudaGetLastError(), "lookup_slot", __FILE__, __LINE__)) return FAIL;
          if (modgpuLocate((const uint32_t *)ms->gpos.p, n, (const uint64_t *)ms->offs[0].p, ns, dRead, dPos, st)) return FAIL;
          ReadsetScan f;
          f.slots = slots; f.slotOf = dSlot; f.kmers = (const uint64_t *)ms->kmers.p; f.readId = dRead; f.pos = dPos;
          f.hit = dHit; f.hitRead = dHitRead; f.hitPos = dHitPos; f.nMiss = dMiss;
          if (mg_ordered_scan(f, n, dScr, dTotal, st)) return FAIL;
          volatile uint64_t *h = (volatile uint64_t *)ms->hMisc.p;
          if (mg_check_cuda(cudaMemcpyAsync((void *)h, dTotal, 8, cudaMemcpyDeviceToHost, st), "d2h", __FILE__, __LINE__) ||
              mg_check_cuda(cudaStreamSynchronize(st), "sync", __FILE__, __LINE__))
            return FAIL;
          nHit = h[0];
          if (nHit)
            { readset_dx_kernel<<<sgrid(nHit), 256, 0, st>>>(dHitRead, dHitPos, nHit, dDx);
              if (mg_check_cuda(cudaGetLastError(), "readset_dx", __FILE__, __LINE__)) return FAIL;
            }
        }
      first_geq_kernel<<<sgrid(ns + 1), 256, 0, st>>>(dHitRead, nHit, ns, total, dOff);
      if (mg_check_cuda(cudaGetLastError(), "first_geq", __FILE__, __LINE__)) return FAIL;
      uint64_t room = total < cap ? cap - total : 0, take = nHit < room ? nHit : room;
      if (take)
        { if (mg_check_cuda(cudaMemcpyAsync(hit + total, dHit, take * 4, cudaMemcpyDeviceToHost, st), "d2h", __FILE__, __LINE__) ||
              mg_check_cuda(cudaMemcpyAsync(dx + total, dDx, take * 2, cudaMemcpyDeviceToHost, st), "d2h", __FILE__, __LINE__))
            return FAIL;
        }
      if (mg_check_cuda(cudaMemcpyAsync(hitOff + r0, dOff, (ns + 1) * 8, cudaMemcpyDeviceToHost, st), "d2h", __FILE__, __LINE__) ||
          mg_check_cuda(cudaMemcpyAsync(nMiss + r0, dMiss, ns * 4, cudaMemcpyDeviceToHost, st), "d2h", __FILE__, __LINE__) ||
          mg_check_cuda(cudaStreamSynchronize(st), "sync", __FILE__, __LINE__))
        return FAIL;
      total += nHit;
      r0 = r1;
    }
  return total;
}
