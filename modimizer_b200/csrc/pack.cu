// pack.cu - K1: bytes -> 2 bits/base, and the read-end flag stream.
//
// replaces: the per-byte conversion loop of seqIOread (reference seqio.c:322
// FASTA, :328-331 FASTQ) with dna2indexConv (seqio.c:643-652) patched N,n -> 0
// (modutils.c:39, modmap.c:97,193).  Inputs containing other letters are out
// of contract (the reference would index patternRC[-2]).
//
// HBM-bound streaming kernel: 1 B/base in, 0.25 B/base out.  One thread packs
// one 64-bit word from 32 bytes (two 16-byte loads; a warp reads 1 KiB
// contiguous), SWAR converts four bytes at a time.
#include "mg_device.cuh"

// nFull: words whose 32 bytes are all inside the input; nWords: words to write
// (the tail is zero = 'a' padding, never selected thanks to the end flags).
template <bool ASCII, bool ALIGNED>
__global__ void __launch_bounds__(256) pack2bit_kernel(const uint8_t *__restrict__ in, uint64_t nBases,
                                                       uint64_t *__restrict__ out, uint64_t nWords)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nWords; w += stride)
    { const uint64_t b0 = w * 32;
      uint64_t word = 0;
      if (b0 + 32 <= nBases && ALIGNED)
        { const uint4 *p = reinterpret_cast<const uint4 *>(in + b0);
          uint4 a = __ldg(p), b = __ldg(p + 1);
          const uint32_t v[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
          word = mg_pack32(v, ASCII);
        }
      else if (b0 < nBases)
        { uint32_t n = (uint32_t)((nBases - b0 < 32) ? (nBases - b0) : 32);
          for (uint32_t j = 0; j < n; ++j)
            word |= (uint64_t)mg_code_of(in[b0 + j], ASCII) << (62 - 2 * j);
        }
      out[w] = word;
    }
}

// one thread per sequence: flag its last base.  d_ends must be zeroed first.
__global__ void __launch_bounds__(256) mark_ends_kernel(const uint64_t *__restrict__ offs, uint64_t nSeq,
                                                        uint32_t *__restrict__ ends)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nSeq; r += stride)
    { uint64_t a = offs[r], b = offs[r + 1];
      if (b > a)
        { uint64_t g = b - 1;
          atomicOr(ends + (g >> 5), 1u << (g & 31));
        }
    }
}

extern "C" uint64_t modgpuPackedWords(uint64_t nBases)
{
  uint64_t words = (nBases + 31) / 32;
  uint64_t tiles = (words + MG_TILE_THREADS - 1) / MG_TILE_THREADS;
  if (!tiles) tiles = 1;
  return tiles * MG_TILE_THREADS + MG_PACK_SLACK_WORDS;
}

extern "C" uint64_t modgpuEndsWords(uint64_t nBases) { return modgpuPackedWords(nBases); }

extern "C" int modgpuPack2bit(const uint8_t *d_bases, uint64_t nBases, int isAscii, uint64_t *d_packed, void *stream)
{
  cudaStream_t st = (cudaStream_t)stream;
  const uint64_t nWords = modgpuPackedWords(nBases);
  const bool aligned = (((uintptr_t)d_bases) & 15) == 0;
  uint64_t blocks = (nWords + 255) / 256;
  uint64_t maxBlocks = (uint64_t)mg_num_sms() * 16;
  if (blocks > maxBlocks) blocks = maxBlocks;
  dim3 grid((unsigned)blocks), block(256);
  if (isAscii)
    { if (aligned) pack2bit_kernel<true, true><<<grid, block, 0, st>>>(d_bases, nBases, d_packed, nWords);
      else pack2bit_kernel<true, false><<<grid, block, 0, st>>>(d_bases, nBases, d_packed, nWords);
    }
  else
    { if (aligned) pack2bit_kernel<false, true><<<grid, block, 0, st>>>(d_bases, nBases, d_packed, nWords);
      else pack2bit_kernel<false, false><<<grid, block, 0, st>>>(d_bases, nBases, d_packed, nWords);
    }
  MG_LAUNCH_CHECK("pack2bit");
  return MODGPU_OK;
}

extern "C" int modgpuMarkEnds(const uint64_t *d_offs, uint64_t nSeq, uint64_t nBases, uint32_t *d_ends, void *stream)
{
  cudaStream_t st = (cudaStream_t)stream;
  MG_CUDA(cudaMemsetAsync(d_ends, 0, modgpuEndsWords(nBases) * sizeof(uint32_t), st));
  if (!nSeq) return MODGPU_OK;
  uint64_t blocks = (nSeq + 255) / 256;
  uint64_t maxBlocks = (uint64_t)mg_num_sms() * 8;
  if (blocks > maxBlocks) blocks = maxBlocks;
  mark_ends_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_offs, nSeq, d_ends);
  MG_LAUNCH_CHECK("mark_ends");
  return MODGPU_OK;
}
