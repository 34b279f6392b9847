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

#define MG_MAX_OWNERS 256

// MODE 0: owner GPU of the k-mer (a = nOwners)
// MODE 1: table region of the k-mer: top b bits of its slot hash (a = slotBits) -
//         inserting region after region keeps the touched part of the table
//         L2-resident instead of paying one DRAM line per random probe
template <int MODE>
__device__ __forceinline__ uint32_t bucket_of(uint64_t kmer, uint32_t a, uint32_t b)
{
  kmer &= 0x3FFFFFFFFFFFFFFFull;
  if (MODE == 0) return mg_owner(kmer, a);
  return (uint32_t)(mg_slot_hash(kmer, a) >> (a - b));
}

template <int MODE>
__global__ void __launch_bounds__(256) owner_count_kernel(const uint64_t *__restrict__ kmers, uint64_t n,
                                                          uint32_t nOwners, uint32_t a, uint32_t b, unsigned long long *counts)
{
  __shared__ uint32_t sC[MG_MAX_OWNERS];
  if (threadIdx.x < MG_MAX_OWNERS) sC[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    { const uint32_t o = bucket_of<MODE>(kmers[i], a, b);
      if (o < MG_MAX_OWNERS) atomicAdd(&sC[o], 1u);
    }
  __syncthreads();
  if (threadIdx.x < nOwners && sC[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)sC[threadIdx.x]);
}

// cursors[o] must start at the exclusive prefix of counts (segment starts)
template <int MODE>
__global__ void __launch_bounds__(256) owner_scatter_kernel(const uint64_t *__restrict__ kmers, uint64_t n,
                                                            uint32_t nOwners, uint32_t a, uint32_t b,
                                                            unsigned long long *cursors, uint64_t *__restrict__ out)
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
              ow[j] = bucket_of<MODE>(km[j], a, b);
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
  owner_count_kernel<0><<<(unsigned)blocks, 256, 0, st>>>(d_kmers, n, nOwners, nOwners, 0, (unsigned long long *)d_counts);
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
  owner_scatter_kernel<0><<<(unsigned)blocks, 256, 0, st>>>(d_kmers, n, nOwners, nOwners, 0, (unsigned long long *)d_cursors, d_out);
  MG_LAUNCH_CHECK("owner_scatter");
  return MODGPU_OK;
}

extern "C" uint32_t modgpuOwnerOf(uint64_t kmer, uint32_t nOwners) { return mg_owner(kmer & 0x3FFFFFFFFFFFFFFFull, nOwners); }

// counts -> exclusive prefix (cursors), one small block
__global__ void bucket_prefix_kernel(const unsigned long long *counts, unsigned long long *cursors, uint32_t nb)
{
  if (threadIdx.x == 0)
    { unsigned long long run = 0;
      for (uint32_t i = 0; i < nb; ++i) { cursors[i] = run; run += counts[i]; }
    }
}

// reorder a selected list by table region (top bucketBits bits of the slot hash).
// d_scratch: 2 * 2^bucketBits uint64.  d_out: n k-mers, grouped by region.
int mg_slot_partition(const uint64_t *d_kmers, uint64_t n, uint32_t slotBits, uint32_t bucketBits,
                      uint64_t *d_out, uint64_t *d_scratch, cudaStream_t st)
{
  const uint32_t nb = 1u << bucketBits;
  if (nb > MG_MAX_OWNERS || bucketBits > slotBits) { mg_set_error("bad bucketBits %u", bucketBits); return MODGPU_EINVAL; }
  if (!n) return MODGPU_OK;
  unsigned long long *counts = (unsigned long long *)d_scratch, *cursors = counts + nb;
  MG_CUDA(cudaMemsetAsync(counts, 0, nb * sizeof(uint64_t), st));
  uint64_t blocks = (n + 2047) / 2048;
  uint64_t maxBlocks = (uint64_t)mg_num_sms() * 8;
  if (blocks > maxBlocks) blocks = maxBlocks;
  owner_count_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(d_kmers, n, nb, slotBits, bucketBits, counts);
  MG_LAUNCH_CHECK("bucket_count");
  bucket_prefix_kernel<<<1, 32, 0, st>>>(counts, cursors, nb);
  MG_LAUNCH_CHECK("bucket_prefix");
  owner_scatter_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(d_kmers, n, nb, slotBits, bucketBits, cursors, d_out);
  MG_LAUNCH_CHECK("bucket_scatter");
  return MODGPU_OK;
}
