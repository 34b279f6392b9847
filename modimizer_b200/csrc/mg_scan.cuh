// mg_scan.cuh - ordered exclusive scan / compaction in two passes over chunks
// of MG_CP_ROWS*256 items: pass A sums f.value(i) per chunk, a one-block scan
// turns the chunk sums into chunk offsets, pass B gives every item its
// exclusive prefix in item order through f.emit(i, prefix, value).
// Streaming, HBM-bound helpers for the table numbering (first-occurrence
// ranks, reference modset.c:57) and for referencePack's loc[] (modmap.c:84-86).
#pragma once
#include "mg_device.cuh"

#define MG_CP_ROWS 16
#define MG_CP_CHUNK (MG_CP_ROWS * 256)

template <class F>
__global__ void __launch_bounds__(256) chunk_sum_kernel(F f, uint64_t n, uint32_t *chunkSums)
{
  const uint64_t base = (uint64_t)blockIdx.x * MG_CP_CHUNK;
  uint32_t c = 0;
#pragma unroll 4
  for (int r = 0; r < MG_CP_ROWS; ++r)
    { uint64_t i = base + (uint64_t)r * 256 + threadIdx.x;
      if (i < n) c += f.value(i);
    }
  c = mg_warp_sum(c);
  __shared__ uint32_t sW[8];
  if (mg_lane() == 0) sW[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0)
    { uint32_t t = 0;
      for (int w = 0; w < 8; ++w) t += sW[w];
      chunkSums[blockIdx.x] = t;
    }
}

// exclusive scan of m chunk sums in place by ONE block of 1024 threads
// (offsets are kept modulo 2^32; the 64-bit grand total goes to *total)
static __global__ void __launch_bounds__(1024) scan_chunks_kernel(uint32_t *sums, uint64_t m, unsigned long long *total)
{
  __shared__ uint32_t sW[32];
  __shared__ unsigned long long sCarry;
  if (threadIdx.x == 0) sCarry = 0;
  __syncthreads();
  for (uint64_t base = 0; base < m; base += 1024)
    { uint64_t i = base + threadIdx.x;
      uint32_t v = (i < m) ? sums[i] : 0u;
      uint32_t incl = mg_warp_incl_scan(v);
      if (mg_lane() == 31) sW[threadIdx.x >> 5] = incl;
      __syncthreads();
      uint32_t pre = 0, tot = 0;
      for (int w = 0; w < 32; ++w) { uint32_t t = sW[w]; if (w < (int)(threadIdx.x >> 5)) pre += t; tot += t; }
      unsigned long long carry = sCarry;
      if (i < m) sums[i] = (uint32_t)(carry + pre + incl - v);
      __syncthreads();
      if (threadIdx.x == 0) sCarry = carry + tot;
      __syncthreads();
    }
  if (threadIdx.x == 0) *total = sCarry;
}

template <class F>
__global__ void __launch_bounds__(256) chunk_emit_kernel(F f, uint64_t n, const uint32_t *chunkOffsets)
{
  __shared__ uint32_t sWarp[8];
  const uint64_t base = (uint64_t)blockIdx.x * MG_CP_CHUNK;
  uint32_t run = chunkOffsets[blockIdx.x];
  for (int r = 0; r < MG_CP_ROWS; ++r)
    { uint64_t i = base + (uint64_t)r * 256 + threadIdx.x;
      uint32_t v = (i < n) ? f.value(i) : 0u;
      uint32_t total;
      uint32_t off = mg_block_excl_scan256(v, sWarp, &total);
      if (i < n) f.emit(i, run + off, v);
      run += total;
      __syncthreads();                                  // sWarp is reused by the next row
    }
}

// scratch: at least n / MG_CP_CHUNK + 1 words; dTotal: device u64
template <class F>
static int mg_ordered_scan(F f, uint64_t n, uint32_t *dScratch, unsigned long long *dTotal, cudaStream_t st)
{
  if (!n) { MG_CUDA(cudaMemsetAsync(dTotal, 0, 8, st)); return MODGPU_OK; }
  uint64_t chunks = (n + MG_CP_CHUNK - 1) / MG_CP_CHUNK;
  chunk_sum_kernel<F><<<(unsigned)chunks, 256, 0, st>>>(f, n, dScratch);
  MG_LAUNCH_CHECK("chunk_sum");
  scan_chunks_kernel<<<1, 1024, 0, st>>>(dScratch, chunks, dTotal);
  MG_LAUNCH_CHECK("scan_chunks");
  chunk_emit_kernel<F><<<(unsigned)chunks, 256, 0, st>>>(f, n, dScratch);
  MG_LAUNCH_CHECK("chunk_emit");
  return MODGPU_OK;
}
