// mg_select.cuh - parameters and workspace layout shared by the hash/select kernels (hash_select.cu, hash_count2.cu)
#pragma once
#include "mg_device.cuh"

struct SelectParams {
  MgKHasher H;
  const uint64_t *packed;
  const uint32_t *ends;
  uint64_t nBases;
  uint32_t nTiles;
  uint32_t strandBit;          // MODGPU_SEL_STRAND
  uint64_t *outKmer;
  uint32_t *outPos;            // nullable
  uint64_t cap;
  unsigned long long *count;   // device total
  uint64_t *status;            // look-back descriptors [nTiles]
  uint32_t *ticket;            // tile ticket
  // SCATTER: selected k-mers go straight into the table's per-region buckets
  // (table.cu bulk insert) instead of a list
  uint32_t slotBits, regionBits, nRegions, bucketCap;
  uint32_t *cursors;           // [nRegions] fill counts, [nRegions] = overflow count
  uint64_t *buckets;
  uint64_t *overflow;
  uint64_t overflowCap;
  // OUT == 2: selected k-mers go into nOwners contiguous segments of a send
  // buffer (multi-GPU: one segment per owner GPU), ownerCursor[o] counts them
  uint32_t nOwners;
  uint32_t *ownerCursor;
  uint64_t *ownerBuf;
  uint64_t ownerCap;
  // LOAD == 2: the batch as bytes (16-byte aligned), K1 fused into the tile loader
  const uint8_t *raw;
  uint32_t rawAscii;
  // LUTK != 0: the 16 KiB candidate table of mg_lut_entry (built per launch into the workspace)
  const uint8_t *lut;
  uint32_t keepBuckets;        // 1: bucket stores ask L2 to keep the line (evict-last)
  // second-generation count kernel (hash_count2.cu): one byte per 2048-base warp tile, non-zero when a sequence ends in
  // (or just before the end of) the tile - the per-base end flags are then only read for those tiles
  const uint8_t *tileFlags;
  MgEval32 E;                  // mg_eval32_prepare(H), computed once on the host: operands straight from the constant bank
  uint32_t regionShift;        // region of a k-mer = high word of its slot-hash product >> regionShift
};

// count mode, second generation (hash_count2.cu); returns MODGPU_OK, or 1 when the configuration is not covered
// (the caller then launches the first-generation kernel)
int mg_count2_launch(const SelectParams &P, int out, int flags, cudaStream_t st);

// workspace layout: [0, 64) ticket and scratch counters, [64, 64 + MG_LUT_SIZE) candidate table, then the
// look-back descriptors of the ordered kernel
#define MG_WS_LUT 64
#define MG_WS_STATUS (64 + MG_LUT_SIZE)


#ifdef __CUDACC__
// the 16 KiB candidate table of the table-driven scan (mg_lut_entry), built per launch
static __global__ void __launch_bounds__(256) lut_build_kernel(const MgKHasher H, uint8_t *lut)
{
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x < MG_LUT_SIZE) lut[x] = (uint8_t)mg_lut_entry(H, x);
}

// ---- geometry of the count kernels: every warp owns tiles of 64 runs = 2048 window starts
#ifndef MG_CNT_WARPS
#define MG_CNT_WARPS 16
#endif
#define MG_CNT_THREADS (MG_CNT_WARPS * 32)
#ifndef MG_CNT_CHUNK
#define MG_CNT_CHUNK 4                                          // warp tiles per scheduling chunk (power of two)
#endif
#define MG_WT_RUNS 64                                          // runs per warp tile
#define MG_WT_BASES (MG_WT_RUNS * MG_RUN)                      // 2048
#define MG_WQ_CAP 256                                          // queue entries per warp (of its 2048 windows)
#define MG_WS_RAW_BYTES (MG_WT_BASES + 32)                     // the tile + the overlap word's 32 bases
#define MG_WS_PACK_BYTES (MG_WT_RUNS * 8 + 16)                 // 64 words + overlap word (+ pad to 16 B)
#define MG_WS_ENDS_BYTES (MG_WT_RUNS * 4 + 16)                 // 64 flag words + 2 (+ pad)

// 16 bytes -> 16 two-bit codes, first base in the top bits (K1 arithmetic, mg_pack4;
// the four gathered bytes are merged with three byte permutes instead of shifts and masks)
template <bool ASCII>
__device__ __forceinline__ uint32_t pack16_dev(const uint4 q)
{
  uint32_t c0, c1, c2, c3;
  if (ASCII)
    { c0 = ((q.x >> 1) ^ (q.x >> 2)) & 0x03030303u; c1 = ((q.y >> 1) ^ (q.y >> 2)) & 0x03030303u;
      c2 = ((q.z >> 1) ^ (q.z >> 2)) & 0x03030303u; c3 = ((q.w >> 1) ^ (q.w >> 2)) & 0x03030303u;
    }
  else
    { // reference codes are 0..3 by contract (include/modgpu.h; seqIOread dies on anything dna2indexConv does not map,
      // seqio.c:643-652, and patternRC[] has four entries, seqhash.h:22): no mask - four instructions per 16 bases of a
      // pass that is bound by instruction issue.  Other byte values give undefined k-mers, nothing else
      c0 = q.x; c1 = q.y; c2 = q.z; c3 = q.w;
    }
  const uint32_t p0 = c0 * 0x40100401u, p1 = c1 * 0x40100401u, p2 = c2 * 0x40100401u, p3 = c3 * 0x40100401u;
  const uint32_t t = __byte_perm(p0, p1, 0x3700), u = __byte_perm(p2, p3, 0x0037);
  return __byte_perm(t, u, 0x3254);
}

// 32 bases starting at b0 of the raw batch, with bounds (the ragged last tile only)
template <bool ASCII>
__device__ __forceinline__ uint64_t pack32_raw(const uint8_t *raw, uint64_t b0, uint64_t nBases)
{
  if (b0 + 32 <= nBases)
    { const uint4 *src = reinterpret_cast<const uint4 *>(raw + b0);
      return ((uint64_t)pack16_dev<ASCII>(__ldg(src)) << 32) | pack16_dev<ASCII>(__ldg(src + 1));
    }
  uint64_t w = 0;
  for (uint32_t j = 0; j < 32 && b0 + j < nBases; ++j)
    w |= (uint64_t)mg_code_of(raw[b0 + j], ASCII) << (62 - 2 * j);
  return w;
}

#endif  // __CUDACC__
