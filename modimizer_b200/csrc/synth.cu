// synth.cu - device generators of the synthetic genomes / readsets of
// include/modgpu_synth.h (same integer code as the host oracle uses), so that
// benchmark inputs are born in HBM and the CPU checkers see identical bytes.
#include "mg_device.cuh"
#include "../../include/modgpu_synth.h"

__global__ void __launch_bounds__(256) synth_genome_kernel(uint64_t seed, uint64_t start, uint64_t n, int dupMode, uint8_t *out)
{
  // one thread = 32 consecutive bases = one generator word when start is aligned
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t nChunks = (n + 31) / 32;
  for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nChunks; c += stride)
    { uint64_t g0 = start + c * 32;
      uint32_t m = (uint32_t)((n - c * 32 < 32) ? (n - c * 32) : 32);
      if ((g0 & 31) == 0 && m == 32 && ((((uintptr_t)out) + c * 32) & 15) == 0)
        { uint64_t w = mg_genome_word(seed, g0 >> 5, dupMode);
          uint32_t v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q)
            { uint32_t x = (uint32_t)(w >> (8 * q)) & 0xFFu;     // 4 bases
              v[q] = (x & 3u) | (((x >> 2) & 3u) << 8) | (((x >> 4) & 3u) << 16) | (((x >> 6) & 3u) << 24);
            }
          uint4 *p = reinterpret_cast<uint4 *>(out + c * 32);
          p[0] = make_uint4(v[0], v[1], v[2], v[3]);
          p[1] = make_uint4(v[4], v[5], v[6], v[7]);
        }
      else
        for (uint32_t j = 0; j < m; ++j) out[c * 32 + j] = mg_genome_base(seed, g0 + j, dupMode);
    }
}

// substitution-only reads: one thread per 16 bases of a read
__global__ void __launch_bounds__(256) synth_reads_kernel(MgReadSpec sp, uint64_t firstRead, uint64_t nReads, uint8_t *out)
{
  const uint32_t L = sp.readLen;
  const uint32_t perRead = (L + 15) / 16;
  const uint64_t total = nReads * perRead;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride)
    { uint64_t rr = t / perRead;
      uint32_t j0 = (uint32_t)(t % perRead) * 16;
      uint64_t r = firstRead + rr;
      uint64_t start; int rev;
      mg_read_layout(&sp, r, &start, &rev);
      uint32_t j1 = j0 + 16 < L ? j0 + 16 : L;
      for (uint32_t j = j0; j < j1; ++j) out[rr * L + j] = mg_read_base(&sp, r, j, start, rev);
    }
}

// ONT-like reads with indels: the walk is sequential, one thread per read
__global__ void __launch_bounds__(128) synth_ont_kernel(MgReadSpec sp, uint64_t firstRead, uint64_t nReads, uint8_t *out)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t rr = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; rr < nReads; rr += stride)
    mg_ont_read(&sp, firstRead + rr, out + rr * sp.readLen);
}

extern "C" int modgpuSynthGenome(uint64_t seed, uint64_t start, uint64_t n, int dupMode, uint8_t *d_codes, void *stream)
{
  if (!n) return MODGPU_OK;
  uint64_t blocks = ((n + 31) / 32 + 255) / 256;
  uint64_t maxBlocks = (uint64_t)mg_num_sms() * 16;
  if (blocks > maxBlocks) blocks = maxBlocks;
  synth_genome_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(seed, start, n, dupMode, d_codes);
  MG_LAUNCH_CHECK("synth_genome");
  return MODGPU_OK;
}

extern "C" int modgpuSynthReads(const void *spec, uint64_t firstRead, uint64_t nReads, int ont, uint8_t *d_codes, void *stream)
{
  if (!nReads) return MODGPU_OK;
  MgReadSpec sp = *(const MgReadSpec *)spec;
  if (sp.readLen == 0 || sp.genomeLen < (uint64_t)sp.readLen * (ont ? 2 : 1) ||
      (sp.pairMode && (sp.fragLen < sp.readLen || sp.genomeLen < sp.fragLen)))
    { mg_set_error("modgpuSynthReads: inconsistent read spec"); return MODGPU_EINVAL; }
  uint64_t maxBlocks = (uint64_t)mg_num_sms() * 16;
  if (ont)
    { uint64_t blocks = (nReads + 127) / 128;
      if (blocks > maxBlocks) blocks = maxBlocks;
      synth_ont_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(sp, firstRead, nReads, d_codes);
    }
  else
    { uint64_t threads = nReads * ((sp.readLen + 15) / 16);
      uint64_t blocks = (threads + 255) / 256;
      if (blocks > maxBlocks) blocks = maxBlocks;
      synth_reads_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(sp, firstRead, nReads, d_codes);
    }
  MG_LAUNCH_CHECK("synth_reads");
  return MODGPU_OK;
}
