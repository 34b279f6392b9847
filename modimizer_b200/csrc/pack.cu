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

// nWords: words to write (the tail is zero = 'a' padding, never selected thanks
// to the end flags).  MIS = misalignment of `in` modulo 16 in 4-byte words
// (0..3) with `sh` = 8 * (misalignment modulo 4): a misaligned batch (a group
// of records inside a larger resident buffer) is still read with aligned
// 16-byte loads, three per thread, and realigned with funnel shifts.
template <bool ASCII, int MIS>
__global__ void __launch_bounds__(256) pack2bit_kernel(const uint8_t *__restrict__ in, uint64_t nBases,
                                                       uint64_t *__restrict__ out, uint64_t nWords, uint32_t sh)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const bool aligned = (MIS == 0 && sh == 0);
  for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nWords; w += stride)
    { const uint64_t b0 = w * 32;
      uint64_t word = 0;
      if (aligned && b0 + 32 <= nBases)
        { const uint4 *p = reinterpret_cast<const uint4 *>(in + b0);
          uint4 a = __ldg(p), b = __ldg(p + 1);
          const uint32_t v[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
          word = mg_pack32(v, ASCII);
        }
      else if (!aligned && b0 + 64 <= nBases)
        { const uint4 *p = reinterpret_cast<const uint4 *>(in + b0 - (MIS * 4 + sh / 8));
          uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
          const uint32_t q[12] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w };
          uint32_t v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __funnelshift_r(q[MIS + j], q[MIS + j + 1], sh);
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
// tileFlags (nullable): one byte per 2048-base warp tile of the count kernels, set for the tile that holds the flag and
// for the tile before it when the flag lies within that tile's 64-base lookahead (hash_count2.cu reads the per-base
// flags only for tiles whose byte is set)
__global__ void __launch_bounds__(256) mark_ends_kernel(const uint64_t *__restrict__ offs, uint64_t nSeq,
                                                        uint32_t *__restrict__ ends, uint8_t *__restrict__ tileFlags)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nSeq; r += stride)
    { uint64_t a = offs[r], b = offs[r + 1];
      if (b > a)
        { uint64_t g = b - 1;
          atomicOr(ends + (g >> 5), 1u << (g & 31));
          if (tileFlags)
            { const uint64_t t = g >> 11;
              tileFlags[t] = 1;
              if (t && (g & 2047) < 64) tileFlags[t - 1] = 1;
            }
        }
    }
}

// the same flags cleared again (every touched word becomes 0: with the clean-buffer invariant of api.cu the
// whole array is zero again afterwards, without a 1-bit-per-base memset per batch)
__global__ void __launch_bounds__(256) unmark_ends_kernel(const uint64_t *__restrict__ offs, uint64_t nSeq,
                                                          uint32_t *__restrict__ ends, uint8_t *__restrict__ tileFlags)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nSeq; r += stride)
    { uint64_t a = offs[r], b = offs[r + 1];
      if (b > a)
        { const uint64_t g = b - 1;
          ends[g >> 5] = 0u;
          if (tileFlags)
            { const uint64_t t = g >> 11;
              tileFlags[t] = 0;
              if (t && (g & 2047) < 64) tileFlags[t - 1] = 0;
            }
        }
    }
}

// set (1) or clear (0) the end flags of nSeq sequences in a buffer that is otherwise all zero
int mg_ends_sparse(const uint64_t *d_offs, uint64_t nSeq, uint32_t *d_ends, uint8_t *d_tileFlags, int set, cudaStream_t st)
{
  if (!nSeq) return MODGPU_OK;
  uint64_t blocks = (nSeq + 255) / 256;
  uint64_t maxBlocks = (uint64_t)mg_num_sms() * 8;
  if (blocks > maxBlocks) blocks = maxBlocks;
  if (set) mark_ends_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_offs, nSeq, d_ends, d_tileFlags);
  else unmark_ends_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_offs, nSeq, d_ends, d_tileFlags);
  MG_LAUNCH_CHECK("mark_ends");
  return MODGPU_OK;
}

extern "C" uint64_t modgpuPackedWords(uint64_t nBases)
{
  uint64_t words = (nBases + 31) / 32;
  uint64_t tiles = (words + MG_TILE_THREADS - 1) / MG_TILE_THREADS;
  if (!tiles) tiles = 1;
  return tiles * MG_TILE_THREADS + MG_PACK_SLACK_WORDS;
}

extern "C" uint64_t modgpuEndsWords(uint64_t nBases) { return modgpuPackedWords(nBases); }

template <bool ASCII>
static void launch_pack(const uint8_t *d_bases, uint64_t nBases, uint64_t *d_packed, uint64_t nWords, dim3 grid, cudaStream_t st)
{
  const uint32_t mis = (uint32_t)(((uintptr_t)d_bases) & 15), sh = 8 * (mis & 3);
  switch (mis >> 2)
    { case 0: pack2bit_kernel<ASCII, 0><<<grid, 256, 0, st>>>(d_bases, nBases, d_packed, nWords, sh); break;
      case 1: pack2bit_kernel<ASCII, 1><<<grid, 256, 0, st>>>(d_bases, nBases, d_packed, nWords, sh); break;
      case 2: pack2bit_kernel<ASCII, 2><<<grid, 256, 0, st>>>(d_bases, nBases, d_packed, nWords, sh); break;
      default: pack2bit_kernel<ASCII, 3><<<grid, 256, 0, st>>>(d_bases, nBases, d_packed, nWords, sh); break;
    }
}

extern "C" int modgpuPack2bit(const uint8_t *d_bases, uint64_t nBases, int isAscii, uint64_t *d_packed, void *stream)
{
  cudaStream_t st = (cudaStream_t)stream;
  const uint64_t nWords = modgpuPackedWords(nBases);
  uint64_t blocks = (nWords + 255) / 256;
  uint64_t maxBlocks = (uint64_t)mg_num_sms() * 16;
  if (blocks > maxBlocks) blocks = maxBlocks;
  dim3 grid((unsigned)blocks);
  if (isAscii) launch_pack<true>(d_bases, nBases, d_packed, nWords, grid, st);
  else launch_pack<false>(d_bases, nBases, d_packed, nWords, grid, st);
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
  mark_ends_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_offs, nSeq, d_ends, nullptr);
  MG_LAUNCH_CHECK("mark_ends");
  return MODGPU_OK;
}
