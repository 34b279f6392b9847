// partition.cu - route selected modimizers to their owner GPU (multi-GPU modset).
//
// The reference has no distributed mode; its only parallel recipe
// is one modset per input merged offline with modsetMerge (modset.c:106-128).
// Here the table is sharded by an independent hash of the k-mer (mg_owner) and
// each rank's selected list is bucketed by owner before the all-to-all exchange.
// Two streaming passes over 8 B per selected k-mer: count per owner, then
// scatter into contiguous per-owner segments (order inside a segment is
// irrelevant: counting is a commutative sum).
#include "mg_device.cuh"

#define MG_MAX_OWNERS 64

__global__ void __launch_bounds__(256) owner_count_kernel(const uint64_t *__restrict__ kmers, uint64_t n,
                                                          uint32_t nOwners, unsigned long long *counts)
{
  __shared__ uint32_t sC[MG_MAX_OWNERS];
  if (threadIdx.x < MG_MAX_OWNERS) sC[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    atomicAdd(&sC[mg_owner(kmers[i] & 0x3FFFFFFFFFFFFFFFull, nOwners)], 1u);
  __syncthreads();
  if (threadIdx.x < nOwners && sC[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)sC[threadIdx.x]);
}

// cursors[o] must start at the exclusive prefix of counts (segment starts)
__global__ void __launch_bounds__(256) owner_scatter_kernel(const uint64_t *__restrict__ kmers, uint64_t n,
                                                            uint32_t nOwners, unsigned long long *cursors,
                                                            uint64_t *__restrict__ out)
{
  __shared__ uint32_t sC[MG_MAX_OWNERS];
  __shared__ unsigned long long sBase[MG_MAX_OWNERS];
  const uint64_t chunk = (uint64_t)blockDim.x * 8;
  for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < n; base += (uint64_t)gridDim.x * chunk)
    { if (threadIdx.x < MG_MAX_OWNERS) sC[threadIdx.x] = 0;
      __syncthreads();
      uint64_t km[8]; uint32_t ow[8], rk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        { uint64_t i = base + (uint64_t)j * blockDim.x + threadIdx.x;
          ow[j] = 0xFFFFFFFFu;
          if (i < n)
            { km[j] = kmers[i];
              ow[j] = mg_owner(km[j] & 0x3FFFFFFFFFFFFFFFull, nOwners);
              rk[j] = atomicAdd(&sC[ow[j]], 1u);
            }
        }
      __syncthreads();
      if (threadIdx.x < nOwners) sBase[threadIdx.x] = sC[threadIdx.x] ? atomicAdd(&cursors[threadIdx.x], (unsigned long long)sC[threadIdx.x]) : 0ull;
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (ow[j] != 0xFFFFFFFFu) out[sBase[ow[j]] + rk[j]] = km[j];
      __syncthreads();
    }
}

extern "C" int modgpuOwnerCount(const uint64_t *d_kmers, uint64_t n, uint32_t nOwners, uint64_t *d_counts, void *stream)
{
  cudaStream_t st = (cudaStream_t)stream;
  if (nOwners < 1 || nOwners > MG_MAX_OWNERS) { mg_set_error("nOwners %u out of range 1..%d", nOwners, MG_MAX_OWNERS); return MODGPU_EINVAL; }
  MG_CUDA(cudaMemsetAsync(d_counts, 0, nOwners * sizeof(uint64_t), st));
  if (!n) return MODGPU_OK;
  uint64_t blocks = (n + 2047) / 2048;
  uint64_t maxBlocks = (uint64_t)mg_num_sms() * 8;
  if (blocks > maxBlocks) blocks = maxBlocks;
  owner_count_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_kmers, n, nOwners, (unsigned long long *)d_counts);
  MG_LAUNCH_CHECK("owner_count");
  return MODGPU_OK;
}

extern "C" int modgpuOwnerScatter(const uint64_t *d_kmers, uint64_t n, uint32_t nOwners, uint64_t *d_cursors,
                                  uint64_t *d_out, void *stream)
{
  cudaStream_t st = (cudaStream_t)stream;
  if (nOwners < 1 || nOwners > MG_MAX_OWNERS) { mg_set_error("nOwners %u out of range 1..%d", nOwners, MG_MAX_OWNERS); return MODGPU_EINVAL; }
  if (!n) return MODGPU_OK;
  uint64_t blocks = (n + 2047) / 2048;
  uint64_t maxBlocks = (uint64_t)mg_num_sms() * 8;
  if (blocks > maxBlocks) blocks = maxBlocks;
  owner_scatter_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_kmers, n, nOwners, (unsigned long long *)d_cursors, d_out);
  MG_LAUNCH_CHECK("owner_scatter");
  return MODGPU_OK;
}

extern "C" uint32_t modgpuOwnerOf(uint64_t kmer, uint32_t nOwners) { return mg_owner(kmer & 0x3FFFFFFFFFFFFFFFull, nOwners); }
