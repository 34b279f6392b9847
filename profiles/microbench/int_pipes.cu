// int_pipes.cu - standalone microbenchmark: sustained per-SM throughput of the integer
// instructions hash_select is made of (sm_100a), alone and mixed, so that the kernel's
// instruction budget can be planned against measured pipe rates instead of guesses.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipes int_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

#define NCHAIN 8
#define ITERS 4096

// OP: one dependent step on chain value x with loop-invariant operands a, b
template <int OP>
__device__ __forceinline__ uint32_t step(uint32_t x, uint32_t a, uint32_t b)
{
  uint32_t r;
  if (OP == 0) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(a), "r"(b));                 // IMAD
  else if (OP == 1) asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(a));                          // IMAD.HI
  else if (OP == 2) { uint64_t w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(x), "r"(a)); r = (uint32_t)(w >> 32) ^ (uint32_t)w; }  // IMAD.WIDE + LOP3
  else if (OP == 3) asm volatile("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(a), "r"(b));           // IDP.4A
  else if (OP == 4) asm volatile("shf.l.wrap.b32 %0, %1, %2, 7;" : "=r"(r) : "r"(x), "r"(a));                  // SHF
  else if (OP == 5) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(x), "r"(a), "r"(b));         // LOP3
  else if (OP == 6) asm volatile("min.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(a));                            // VIMNMX
  else if (OP == 7) asm volatile("prmt.b32 %0, %1, %2, 0x2143;" : "=r"(r) : "r"(x), "r"(a));                   // PRMT
  else if (OP == 8) asm volatile("brev.b32 %0, %1;" : "=r"(r) : "r"(x));                                       // BREV
  else if (OP == 9) asm volatile("popc.b32 %0, %1;" : "=r"(r) : "r"(x));                                       // POPC
  else if (OP == 10) asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(a));                           // IADD3
  else if (OP == 11) asm volatile("{.reg .pred p; setp.lt.u32 p, %1, %2; @p or.b32 %0, %1, %3; @!p mov.b32 %0, %1;}" : "=r"(r) : "r"(x), "r"(a), "r"(b)); // ISETP + sel
  else if (OP == 12) asm volatile("shl.b32 %0, %1, 3;" : "=r"(r) : "r"(x));                                    // SHL (may become IMAD.SHL)
  else if (OP == 13) asm volatile("vmin2.u32.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(a), "r"(b));     // video SIMD min
  else if (OP == 14) asm volatile("bfe.u32 %0, %1, 5, 8;" : "=r"(r) : "r"(x));                                 // BFE
  else if (OP == 15) asm volatile("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));                                     // FLO
  else r = x;
  return r;
}

// MIX kernels: A ops of OP1 + B ops of OP2 per chain step
template <int OP1, int N1, int OP2, int N2>
__global__ void __launch_bounds__(256) bench(uint32_t *out, uint32_t a, uint32_t b, long long *cycles)
{
  uint32_t x[NCHAIN];
#pragma unroll
  for (int j = 0; j < NCHAIN; ++j) x[j] = threadIdx.x * 2654435761u + j * 40503u + blockIdx.x;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it)
    {
#pragma unroll
      for (int j = 0; j < NCHAIN; ++j)
        {
#pragma unroll
          for (int u = 0; u < N1; ++u) x[j] = step<OP1>(x[j], a, b);
#pragma unroll
          for (int u = 0; u < N2; ++u) x[j] = step<OP2>(x[j], a, b);
        }
    }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < NCHAIN; ++j) s ^= x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP1, int N1, int OP2, int N2>
static void run(const char *name, uint32_t *out, long long *cyc, int sms)
{
  const int blocksPerSm = 4;                        // 1024 threads per SM = 8 warps per SMSP
  const int grid = sms * blocksPerSm;
  bench<OP1, N1, OP2, N2><<<grid, 256>>>(out, 0x9E3779B1u, 0x85EBCA77u, cyc);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  bench<OP1, N1, OP2, N2><<<grid, 256>>>(out, 0x9E3779B1u, 0x85EBCA77u, cyc);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  long long *h = (long long *)malloc(grid * sizeof(long long));
  CK(cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
  free(h);
  const double opsPerSm = (double)blocksPerSm * 256 * NCHAIN * ITERS * (N1 + N2);
  printf("%-34s %7.1f thread-ops/clk/SM   (%.3f ms, %.0f cycles)\n", name, opsPerSm / avg, ms, avg);
}

int main()
{
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  printf("%s, %d SMs\n", p.name, sms);
  uint32_t *out; long long *cyc;
  CK(cudaMalloc(&out, (size_t)sms * 4 * 256 * 4)); CK(cudaMalloc(&cyc, (size_t)sms * 4 * 8));
  run<0, 1, 16, 0>("IMAD (mad.lo)", out, cyc, sms);
  run<1, 1, 16, 0>("IMAD.HI (mul.hi.u32)", out, cyc, sms);
  run<2, 1, 16, 0>("IMAD.WIDE + LOP3", out, cyc, sms);
  run<3, 1, 16, 0>("IDP.4A (dp4a)", out, cyc, sms);
  run<4, 1, 16, 0>("SHF (funnel)", out, cyc, sms);
  run<5, 1, 16, 0>("LOP3", out, cyc, sms);
  run<6, 1, 16, 0>("VIMNMX (min.u32)", out, cyc, sms);
  run<7, 1, 16, 0>("PRMT", out, cyc, sms);
  run<8, 1, 16, 0>("BREV", out, cyc, sms);
  run<9, 1, 16, 0>("POPC", out, cyc, sms);
  run<10, 1, 16, 0>("IADD3 (add)", out, cyc, sms);
  run<11, 1, 16, 0>("ISETP + 2 predicated", out, cyc, sms);
  run<12, 1, 16, 0>("SHL imm", out, cyc, sms);
  run<13, 1, 16, 0>("vmin2 (video SIMD)", out, cyc, sms);
  run<14, 1, 16, 0>("BFE", out, cyc, sms);
  run<15, 1, 16, 0>("FLO (bfind)", out, cyc, sms);
  printf("-- mixes (does the second op ride a different pipe?)\n");
  run<0, 1, 5, 1>("IMAD + LOP3", out, cyc, sms);
  run<0, 1, 4, 1>("IMAD + SHF", out, cyc, sms);
  run<4, 1, 5, 1>("SHF + LOP3", out, cyc, sms);
  run<3, 1, 5, 1>("IDP.4A + LOP3", out, cyc, sms);
  run<3, 1, 0, 1>("IDP.4A + IMAD", out, cyc, sms);
  run<1, 1, 5, 1>("IMAD.HI + LOP3", out, cyc, sms);
  run<1, 1, 0, 1>("IMAD.HI + IMAD", out, cyc, sms);
  run<0, 2, 5, 3>("2 IMAD + 3 LOP3", out, cyc, sms);
  run<0, 2, 5, 5>("2 IMAD + 5 LOP3", out, cyc, sms);
  run<7, 1, 0, 1>("PRMT + IMAD", out, cyc, sms);
  run<6, 1, 0, 1>("VIMNMX + IMAD", out, cyc, sms);
  run<8, 1, 5, 1>("BREV + LOP3", out, cyc, sms);
  run<9, 1, 5, 1>("POPC + LOP3", out, cyc, sms);
  return 0;
}
