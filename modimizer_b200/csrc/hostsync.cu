// hostsync.cu - keeping a HOST Modset (reference modset.h:17-28) and its device twin in step.
//
// The reference's callers read and write the public struct fields directly (ms->depth[index] modutils.c:26,
// ms->value[i] / ms->max modutils.c:57-69, ms->info through the macros of modset.h:53-69), so a drop-in has to
// hand the arrays back whenever control returns to caller code.  libmodshim.so (csrc/shim) does
// that with the entry points below; everything that computes runs on the device:
//
//   modgpuModsetIndexFindBatch   batched modsetIndexFind (modset.c:45-62): lookup, or find-or-insert with the
//                                reference's numbering (index = ++max in input order), no depth change
//   modgpuModsetSetDepthInfo     host depth[] / info[] -> device (the caller changed them: ++depth, msSetCopy*)
//   modgpuModsetReferenceIndex   the reference's own index[] table (home slot hash & mask, odd double-hashing
//                                stride, entries inserted in index order: modset.c:48-57) built on the device
#include <string.h>
#include "mg_api.h"
#include "mg_table.cuh"

uint8_t *mg_table_info_hi(ModgpuTable *t);
void mg_table_set_track_info(ModgpuTable *t, bool on);
MgSlot *mg_table_slots_on(ModgpuTable *t, cudaStream_t st);
int mg_table_ensure_info(ModgpuTable *t, cudaStream_t st);

static unsigned hgrid(uint64_t n)
{
  uint64_t blocks = (n + 255) / 256;
  const uint64_t maxBlocks = (uint64_t)mg_num_sms() * 16;
  if (blocks > maxBlocks) blocks = maxBlocks;
  if (!blocks) blocks = 1;
  return (unsigned)blocks;
}

// ------------------------------------------------------- batched IndexFind --
extern "C" int modgpuModsetIndexFindBatch(ModgpuModset *ms, const uint64_t *kmers, uint64_t n, int isAdd, uint32_t *index)
{
  if (!ms || (n && (!kmers || !index))) { mg_set_error("modgpuModsetIndexFindBatch: null argument"); return MODGPU_EINVAL; }
  int rc = mg_modset_ensure_numbered(ms);
  if (rc) return rc;
  if (!n) return MODGPU_OK;
  if (n >= (1ull << 30) - 2) { mg_set_error("modgpuModsetIndexFindBatch: batch of %llu exceeds 2^30", (unsigned long long)n); return MODGPU_EINVAL; }
  cudaStream_t st = ms->stream;
  if ((rc = ms->kmers.ensure(n * 8)) || (rc = ms->slot.ensure(n * 4)) || (rc = ms->gpos.ensure(n * 4))) return rc;
  MG_CUDA(cudaMemcpyAsync(ms->kmers.p, kmers, n * 8, cudaMemcpyHostToDevice, st));
  uint32_t *dOut = (uint32_t *)ms->gpos.p;
  if (!isAdd)
    { ProfScope p(ms, MODGPU_T_OTHER, 1);
      if ((rc = mg_table_lookup_dev(ms->table, (const uint64_t *)ms->kmers.p, nullptr, n, dOut, st))) return rc;
    }
  else
    { ProfScope p(ms, MODGPU_T_INSERT, 3);
      // exactOrder 2: first-occurrence ordinals recorded, counts untouched; the numbering hands out ++max in input order
      if ((rc = mg_table_insert_dev(ms->table, (const uint64_t *)ms->kmers.p, nullptr, n, (uint32_t *)ms->slot.p, 2, st))) return rc;
      if ((rc = modgpuTableNumber(ms->table, (const uint32_t *)ms->slot.p, n, dOut, st))) return rc;
    }
  MG_CUDA(cudaMemcpyAsync(index, dOut, n * 4, cudaMemcpyDeviceToHost, st));
  MG_CUDA(cudaStreamSynchronize(st));
  if (!isAdd) for (uint64_t i = 0; i < n; ++i) index[i] >>= 2;          // the lookup returns index << 2 | copy
  else if (modgpuTableEntries(ms->table, st) == 0xFFFFFFFFFFFFFFFFull) return MODGPU_EFULL;   // modset.c:58
  return MODGPU_OK;
}

// ---------------------------------------------------------- depth / info --
__global__ void __launch_bounds__(256) set_depth_info_kernel(MgSlot *slots, uint64_t nSlots, const uint16_t *__restrict__ depth,
                                                             const uint8_t *__restrict__ info, uint8_t *__restrict__ infoHi, uint64_t n)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nSlots; i += stride)
    { uint4 v = reinterpret_cast<const uint4 *>(slots)[i];
      if ((v.x & v.y) == 0xFFFFFFFFu || v.w >= MG_AUX_ORD) continue;
      const uint64_t ix = (uint64_t)(v.w >> 2) - 1;
      if (ix >= n) continue;
      if (depth) v.z = depth[ix];
      if (info) { v.w = (v.w & ~3u) | (info[ix] & 3u); if (infoHi) infoHi[ix] = (uint8_t)(info[ix] >> 2); }
      reinterpret_cast<uint4 *>(slots)[i] = v;
    }
}

extern "C" int modgpuModsetSetDepthInfo(ModgpuModset *ms, const uint16_t *depth, const uint8_t *info, uint64_t n)
{
  if (!ms) { mg_set_error("modgpuModsetSetDepthInfo: null modset"); return MODGPU_EINVAL; }
  int rc = mg_modset_ensure_numbered(ms);
  if (rc) return rc;
  const uint64_t have = mg_table_numbered(ms->table);
  if (n > have) { mg_set_error("modgpuModsetSetDepthInfo: %llu entries given, the set has %llu", (unsigned long long)n, (unsigned long long)have); return MODGPU_EINVAL; }
  if (!n || (!depth && !info)) return MODGPU_OK;
  cudaStream_t st = ms->stream;
  if ((rc = ms->expo.ensure(n * 3 + 64))) return rc;
  uint16_t *dD = (uint16_t *)ms->expo.p;
  uint8_t *dI = (uint8_t *)(dD + n);
  if (depth) MG_CUDA(cudaMemcpyAsync(dD, depth, n * 2, cudaMemcpyHostToDevice, st));
  if (info)
    { MG_CUDA(cudaMemcpyAsync(dI, info, n, cudaMemcpyHostToDevice, st));
      if ((rc = mg_table_ensure_info(ms->table, st))) return rc;       // the whole byte is kept from here on
    }
  MgSlot *slots = mg_table_slots_on(ms->table, st);
  if (!slots) return MODGPU_ECUDA;
  const uint64_t nSlots = modgpuTableSlots(ms->table);
  { ProfScope p(ms, MODGPU_T_OTHER, 1);
    set_depth_info_kernel<<<hgrid(nSlots), 256, 0, st>>>(slots, nSlots, depth ? dD : nullptr, info ? dI : nullptr,
                                                        info ? mg_table_info_hi(ms->table) : nullptr, n);
    MG_LAUNCH_CHECK("set_depth_info");
  }
  if (depth) ms->depthIsZero = false;
  MG_CUDA(cudaStreamSynchronize(st));                   // the host arrays may change again as soon as we return
  return MODGPU_OK;
}

// ----------------------------------------------------- reference index[] --
// Sequential insertion with open addressing and no deletions puts entry i into the first slot of ITS probe sequence
// that no entry j < i holds in the final table.  That layout is the unique fixed point of "the smallest index wins a
// slot": every entry walks its probe sequence past slots held by smaller indices and claims the first other one with
// atomicMin; an entry that finds itself evicted walks on.  Slot values only decrease, so a few rounds converge
// (load <= 1/4 by the reference's own capacity rule, modset.c:24-26).  No host-side probing anywhere.
#define MG_RIX_EMPTY 0xFFFFFFFFu

__global__ void __launch_bounds__(256) ref_index_round_kernel(const uint64_t *__restrict__ value, uint64_t n, uint64_t factor, uint32_t shift,
                                                              uint32_t bits, uint32_t *tab, uint32_t *cur, int first, uint32_t *changed)
{
  const uint64_t mask = (((uint64_t)1) << bits) - 1;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  bool any = false;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
    { const uint32_t i = (uint32_t)j + 1;                              // the entry's index, 1-based
      const uint64_t hash = (value[j] * factor) >> shift;              // seqhash(), seqhash.h:58
      const uint64_t diff = ((hash >> bits) & mask) | 1;               // modset.c:52
      uint64_t s = first ? (hash & mask) : (uint64_t)cur[j];
      for (;;)
        { const uint32_t c = tab[s];
          if (c == i) break;                                           // still mine
          if (c > i)
            { const uint32_t old = atomicMin(&tab[s], i);
              if (old > i) { any = true; break; }                      // claimed (whoever held it moves on next round)
            }
          s = (s + diff) & mask;
        }
      cur[j] = (uint32_t)s;
    }
  if (__syncthreads_or(any) && threadIdx.x == 0) atomicExch(changed, 1u);
}

__global__ void __launch_bounds__(256) ref_index_finish_kernel(uint32_t *tab, uint64_t n)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) if (tab[i] == MG_RIX_EMPTY) tab[i] = 0u;
}

// d_index: device, 2^bits words.  Leaves the table on the device (mg_modset_reference_index_device) or copies it out.
int mg_modset_reference_index_device(ModgpuModset *ms, uint32_t **d_index_out)
{
  int rc = mg_modset_ensure_numbered(ms);
  if (rc) return rc;
  const uint64_t n = mg_table_numbered(ms->table);
  const uint32_t bits = (uint32_t)ms->bits;
  const uint64_t tableSize = 1ull << bits;
  cudaStream_t st = ms->stream;
  if ((rc = ms->kmers2.ensure(tableSize * 4)) || (rc = ms->expo.ensure(n * 8 + 64)) || (rc = ms->slot.ensure(n * 4 + 64))) return rc;
  uint32_t *tab = (uint32_t *)ms->kmers2.p, *cur = (uint32_t *)ms->slot.p;
  uint64_t *dV = (uint64_t *)ms->expo.p;
  MG_CUDA(cudaMemsetAsync(tab, 0xFF, tableSize * 4, st));
  if (n)
    { if ((rc = modgpuTableExport(ms->table, dV, nullptr, nullptr, nullptr, st))) return rc;
      uint32_t *dChanged = (uint32_t *)((char *)ms->misc.p + 1024);
      volatile uint32_t *hChanged = (volatile uint32_t *)((char *)ms->hMisc.p + 1024);
      ProfScope p(ms, MODGPU_T_OTHER, 0);
      for (int round = 0;; ++round)
        { MG_CUDA(cudaMemsetAsync(dChanged, 0, 4, st));
          for (int sub = 0; sub < 4; ++sub)              // a few rounds between readbacks
            { ref_index_round_kernel<<<hgrid(n), 256, 0, st>>>(dV, n, ms->hasher.factor1, (uint32_t)ms->hasher.shift1, bits, tab, cur,
                                                               (round == 0 && sub == 0) ? 1 : 0, dChanged);
              MG_LAUNCH_CHECK("ref_index_round");
              ++ms->launches[MODGPU_T_OTHER];
              if (round == 0 && sub == 0) MG_CUDA(cudaMemsetAsync(dChanged, 0, 4, st));   // the first round always claims
            }
          MG_CUDA(cudaMemcpyAsync((void *)hChanged, dChanged, 4, cudaMemcpyDeviceToHost, st));
          MG_CUDA(cudaStreamSynchronize(st));
          if (!*hChanged) break;
          if (round > 4096) { mg_set_error("reference index build did not converge"); return MODGPU_ECUDA; }
        }
    }
  ref_index_finish_kernel<<<hgrid(tableSize), 256, 0, st>>>(tab, tableSize);
  MG_LAUNCH_CHECK("ref_index_finish");
  *d_index_out = tab;
  return MODGPU_OK;
}

extern "C" int modgpuModsetReferenceIndex(ModgpuModset *ms, uint32_t *index)
{
  if (!ms || !index) { mg_set_error("modgpuModsetReferenceIndex: null argument"); return MODGPU_EINVAL; }
  uint32_t *tab = nullptr;
  int rc = mg_modset_reference_index_device(ms, &tab);
  if (rc) return rc;
  MG_CUDA(cudaMemcpyAsync(index, tab, (1ull << ms->bits) * 4, cudaMemcpyDeviceToHost, ms->stream));
  MG_CUDA(cudaStreamSynchronize(ms->stream));
  return MODGPU_OK;
}
