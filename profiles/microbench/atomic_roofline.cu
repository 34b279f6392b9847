// atomic_roofline.cu - standalone microbenchmark: what can a B200 sustain for the
// access pattern of the modset insert (random 16-byte slots in a multi-GB table)?
// This measures the "atomic / L2 roofline" SURVEY 8(d) asks the insert kernel to
// be reported against.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomic_roofline atomic_roofline.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x)
{
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}

struct Slot { unsigned long long key; unsigned int count, aux; };

// mode 0: 8-byte load   1: 16-byte load   2: CAS(EMPTY->key)   3: RED.ADD on count
// mode 4: load + CAS + RED (the insert)   5: load + RED (existing key)
template <int MODE, int ILP>
__global__ void __launch_bounds__(256) probe(Slot *t, uint64_t mask, uint64_t n, uint64_t seed, unsigned long long *sink, uint32_t regionBits = 0)
{
  const uint64_t perRegion = regionBits ? (n >> regionBits) + 1 : 0;
  const uint32_t slotBits = 64 - __clzll(mask);
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long acc = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * ILP)
    { uint64_t s[ILP]; unsigned long long v[ILP];
#pragma unroll
      for (int j = 0; j < ILP; ++j)
        { uint64_t e = i + j * stride;
          s[j] = mix(seed + e) & mask;
          if (regionBits) { uint64_t rg = e / perRegion; if (rg >> regionBits) rg = (1ull << regionBits) - 1; s[j] = (s[j] >> regionBits) | (rg << (slotBits - regionBits)); }
        }
#pragma unroll
      for (int j = 0; j < ILP; ++j)
        { if (MODE == 0 || MODE == 4 || MODE == 5) v[j] = __ldcg(&t[s[j]].key);
          else if (MODE == 1) { uint4 q = __ldcg((const uint4 *)&t[s[j]]); v[j] = q.x + q.w; }
          else v[j] = 0;
        }
#pragma unroll
      for (int j = 0; j < ILP; ++j)
        { if (MODE == 2) v[j] = atomicCAS(&t[s[j]].key, 0xFFFFFFFFFFFFFFFFull, (unsigned long long)s[j]);
          if (MODE == 4 && v[j] == 0xFFFFFFFFFFFFFFFFull) v[j] = atomicCAS(&t[s[j]].key, 0xFFFFFFFFFFFFFFFFull, (unsigned long long)s[j]);
          if (MODE == 3 || MODE == 4 || MODE == 5) atomicAdd(&t[s[j]].count, 1u);
          acc += v[j];
        }
    }
  if (acc == 0x123456789ull) *sink = acc;
}

template <int MODE, int ILP>
static void run(const char *name, Slot *t, uint64_t slots, uint64_t n, unsigned long long *sink, int blocksPerSm, uint32_t regionBits = 0)
{
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep)
    { CK(cudaMemset(t, 0xFF, slots * sizeof(Slot)));
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(a));
      probe<MODE, ILP><<<148 * blocksPerSm, 256>>>(t, slots - 1, n, 1234 + rep, sink, regionBits);
      CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      if (rep && ms < best) best = ms;
    }
  printf("%-34s slots 2^%2d (%6.0f MB)  ILP %d  blocks/SM %2d regions 2^%u : %8.3f ms  %7.2f G ops/s\n", name,
         (int)__builtin_ctzll(slots), slots * 16.0 / 1048576, ILP, blocksPerSm, regionBits, best, n / best / 1e6);
}

int main()
{
  const uint64_t n = 48ull << 20;
  Slot *t; unsigned long long *sink;
  CK(cudaMalloc(&t, (1ull << 27) * sizeof(Slot))); CK(cudaMalloc(&sink, 8));
  for (uint64_t bits : { 22, 27 })
    { uint64_t slots = 1ull << bits;
      run<0, 1>("load 8B", t, slots, n, sink, 8);
      run<0, 4>("load 8B", t, slots, n, sink, 8);
      run<1, 1>("load 16B", t, slots, n, sink, 8);
      run<1, 4>("load 16B", t, slots, n, sink, 8);
      run<2, 1>("CAS", t, slots, n, sink, 8);
      run<2, 4>("CAS", t, slots, n, sink, 8);
      run<3, 1>("RED add", t, slots, n, sink, 8);
      run<3, 4>("RED add", t, slots, n, sink, 8);
      run<4, 1>("load+CAS+RED (insert, new keys)", t, slots, n, sink, 8);
      run<4, 2>("load+CAS+RED (insert, new keys)", t, slots, n, sink, 8);
      run<4, 4>("load+CAS+RED (insert, new keys)", t, slots, n, sink, 8);
      run<4, 4>("load+CAS+RED (insert, new keys)", t, slots, n, sink, 16);
      run<5, 1>("load+RED", t, slots, n, sink, 8);
      run<5, 4>("load+RED", t, slots, n, sink, 8);
    }
  printf("-- region-ordered inserts into the 2 GB table (emulates the partitioned list)\n");
  for (uint32_t rb : { 4u, 5u, 6u, 7u, 8u, 10u })
    for (int bps : { 8, 16 })
      { run<4, 1>("insert, region-ordered", t, 1ull << 27, n, sink, bps, rb);
        run<4, 4>("insert, region-ordered", t, 1ull << 27, n, sink, bps, rb);
      }
  return 0;
}
