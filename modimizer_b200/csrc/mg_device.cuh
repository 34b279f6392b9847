// mg_device.cuh - device-only helpers: error plumbing, warp/block scans, the
// mbarrier + TMA bulk-copy primitives used to stage packed sequence in shared
// memory (sm_100a; SASS: UBLKCP / SYNCS), and a generic two-pass ordered
// compaction used by the table kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "mg_common.cuh"
#include "../../include/modgpu.h"

// ---------------------------------------------------------------- host side
void mg_set_error(const char *fmt, ...);
int mg_check_cuda(cudaError_t e, const char *what, const char *file, int line);
int mg_num_sms();

#define MG_CUDA(call)                                                            \
  do { int mg_rc_ = mg_check_cuda((call), #call, __FILE__, __LINE__);            \
       if (mg_rc_) return mg_rc_; } while (0)
#define MG_CUDA_PTR(call)                                                        \
  do { if (mg_check_cuda((call), #call, __FILE__, __LINE__)) return nullptr; } while (0)
#define MG_LAUNCH_CHECK(name) MG_CUDA(cudaGetLastError())

#ifdef __CUDACC__
// -------------------------------------------------------------- warp / block
__device__ __forceinline__ uint32_t mg_lane() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t mg_warp_incl_scan(uint32_t v)
{
  uint32_t lane = mg_lane();
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
    { uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= (uint32_t)o) v += n;
    }
  return v;
}

__device__ __forceinline__ uint32_t mg_warp_sum(uint32_t v)
{
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// exclusive scan of one value per thread over a block of NWARPS warps.
// sWarp: NWARPS words of shared memory.  Returns the thread's exclusive prefix
// and the block total.  Contains one __syncthreads.
template <int NWARPS>
__device__ __forceinline__ uint32_t mg_block_excl_scan(uint32_t v, uint32_t *sWarp, uint32_t *total)
{
  uint32_t incl = mg_warp_incl_scan(v);
  uint32_t w = threadIdx.x >> 5;
  if (mg_lane() == 31) sWarp[w] = incl;
  __syncthreads();
  uint32_t pre = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < NWARPS; ++i)
    { uint32_t t = sWarp[i];
      if ((uint32_t)i < w) pre += t;
      tot += t;
    }
  *total = tot;
  return pre + incl - v;
}

__device__ __forceinline__ uint32_t mg_block_excl_scan256(uint32_t v, uint32_t *sWarp, uint32_t *total)
{ return mg_block_excl_scan<MG_TILE_THREADS / 32>(v, sWarp, total); }

// ------------------------------------------------ decoupled look-back scan --
// status word: flag (high 32 bits: 0 none, 1 tile aggregate, 2 inclusive
// prefix) | value (low 32 bits), written with one 64-bit store so no fence is
// needed between flag and value.  Called by all 32 lanes of one warp; tiles are
// claimed through an atomic ticket so every predecessor is owned by a running
// block (no deadlock whatever the residency).
#define MG_ST_AGG 1ull
#define MG_ST_INC 2ull

__device__ __forceinline__ uint64_t mg_ld_volatile64(const uint64_t *p)
{
  uint64_t v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void mg_st_volatile64(uint64_t *p, uint64_t v)
{
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint32_t mg_lookback(uint64_t *status, uint32_t tile, uint32_t total)
{
  const uint32_t lane = mg_lane();
  if (tile == 0)
    { if (lane == 0) mg_st_volatile64(status, (MG_ST_INC << 32) | total);
      return 0;
    }
  if (lane == 0) mg_st_volatile64(status + tile, (MG_ST_AGG << 32) | total);
  uint32_t excl = 0;
  int64_t look = (int64_t)tile - 1;
  for (;;)
    { int64_t idx = look - lane;
      uint64_t v = (idx >= 0) ? mg_ld_volatile64(status + idx) : (MG_ST_INC << 32);
      uint32_t flag = (uint32_t)(v >> 32);
      uint32_t pending = __ballot_sync(0xffffffffu, flag == 0);
      uint32_t incs = __ballot_sync(0xffffffffu, flag == (uint32_t)MG_ST_INC);
      int first = incs ? (__ffs(incs) - 1) : 32;
      uint32_t need = (first >= 31) ? 0xffffffffu : ((2u << first) - 1u);
      if (pending & need) continue;                         // predecessors not published yet
      uint32_t val = ((int)lane <= first) ? (uint32_t)v : 0u;
      excl += mg_warp_sum(val);
      if (first < 32) break;
      look -= 32;
    }
  if (lane == 0) mg_st_volatile64(status + tile, (MG_ST_INC << 32) | (uint64_t)(excl + total));
  return excl;
}

// -------------------------------------------------------- mbarrier and TMA --
__device__ __forceinline__ uint32_t mg_smem_addr(const void *p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mg_mbar_init(uint64_t *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mg_smem_addr(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mg_fence_barrier_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mg_fence_proxy_async()
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mg_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mg_smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mg_mbar_wait(uint64_t *bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MG_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MG_DONE_%=;\n"
      "bra MG_WAIT_%=;\n"
      "MG_DONE_%=:\n"
      "}\n" ::"r"(mg_smem_addr(bar)), "r"(parity) : "memory");
}

// 1-D bulk copy global -> shared, completion signalled on an mbarrier
// (cp.async.bulk = the TMA engine without a tensor map; SASS UBLKCP).
// dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void mg_tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(mg_smem_addr(dst)),
      "l"(src), "r"(bytes), "r"(mg_smem_addr(bar))
      : "memory");
}

// L2 eviction-priority policies (the createpolicy.fractional encodings CUTLASS ships as TMA::CacheHintSm90)
#define MG_L2_EVICT_FIRST 0x12F0000000000000ull
#define MG_L2_EVICT_LAST  0x14F0000000000000ull

// the same bulk copy with an L2 cache hint: streaming input is marked evict-first so that it does not push
// the half-filled tail sectors of the output buckets out of L2
__device__ __forceinline__ void mg_tma_load_1d_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy)
{
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(mg_smem_addr(dst)),
      "l"(src), "r"(bytes), "r"(mg_smem_addr(bar)), "l"(policy)
      : "memory");
}

// 8-byte store that asks L2 to keep the line (bucket tails are written 8 bytes at a time: a sector evicted
// before its four k-mers arrived costs DRAM a read-modify-write)
__device__ __forceinline__ void mg_st_keep(uint64_t *p, uint64_t v)
{
  asm volatile("st.global.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(MG_L2_EVICT_LAST) : "memory");
}

#endif  // __CUDACC__
