// mg_table.cuh - device-side probing rules of the modset table, shared by the
// kernels of table.cu and setops.cu.  Linear probing confined to aligned regions
// of MG_REGION_SLOTS slots, so that a region is a self-contained sub-table that
// one block can build in shared memory (table.cu, region_build_kernel).
#pragma once
#include "mg_device.cuh"

#define MG_REGION_BITS 11
#define MG_REGION_SLOTS (1u << MG_REGION_BITS)
__device__ __forceinline__ uint64_t next_slot(uint64_t s)
{ return (s & ~(uint64_t)(MG_REGION_SLOTS - 1)) | ((s + 1) & (MG_REGION_SLOTS - 1)); }

// find-or-insert; returns the slot or UINT64_MAX when the table is full
__device__ __forceinline__ uint64_t probe_insert(MgSlot *slots, uint32_t slotBits, uint64_t key, bool *isNew)
{
  uint64_t s = mg_slot_hash(key, slotBits);
  *isNew = false;
  for (uint32_t probes = 0; probes < MG_REGION_SLOTS; ++probes, s = next_slot(s))
    { unsigned long long *kp = reinterpret_cast<unsigned long long *>(&slots[s].key);
      unsigned long long cur = __ldcg(kp);
      if (cur == key) return s;
      if (cur == MG_EMPTY)
        { unsigned long long old = atomicCAS(kp, MG_EMPTY, (unsigned long long)key);
          if (old == MG_EMPTY) { *isNew = true; return s; }
          if (old == key) return s;
        }
    }
  return 0xFFFFFFFFFFFFFFFFull;
}

__device__ __forceinline__ uint64_t probe_find(const MgSlot *slots, uint32_t slotBits, uint64_t key)
{
  uint64_t s = mg_slot_hash(key, slotBits);
  for (uint32_t probes = 0; probes < MG_REGION_SLOTS; ++probes, s = next_slot(s))
    { unsigned long long cur = __ldcg(reinterpret_cast<const unsigned long long *>(&slots[s].key));
      if (cur == key) return s;
      if (cur == MG_EMPTY) break;
    }
  return 0xFFFFFFFFFFFFFFFFull;
}

