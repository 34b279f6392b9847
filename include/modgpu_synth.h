/* modgpu_synth.h - deterministic integer-only synthetic genomes / readsets.
 *
 * One definition compiled three ways (gcc for the oracle and the CPU baseline,
 * g++ for host tests, nvcc for the device generators in csrc/synth.cu) so that the
 * CPU oracle and the GPU path see byte-identical inputs without shipping files.
 * No floating point anywhere: every decision is an integer compare on a
 * counter-based 64-bit mixer, so host and device agree bit for bit.
 *
 * Shapes follow SURVEY.md section 8(d) (configs C1..C5 of BASELINE.json).
 * Bases are produced as reference byte codes a=0 c=1 g=2 t=3, i.e. what the
 * reference's seqIOread hands to addSequence after dna2indexConv
 * (reference seqio.c:643-652 with the N->0 patch of modutils.c:39).
 */
#ifndef MODGPU_SYNTH_H
#define MODGPU_SYNTH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define MG_HD __host__ __device__ __forceinline__
#else
#define MG_HD static inline
#endif

/* splitmix64 finaliser: a bijective 64-bit mixer */
MG_HD uint64_t mg_mix64(uint64_t x)
{
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

/* counter-based generator: independent streams of 64-bit values */
MG_HD uint64_t mg_rng(uint64_t seed, uint64_t stream, uint64_t idx)
{
  return mg_mix64(mg_mix64(seed ^ (stream * 0xD6E8FEB86659FD93ull)) + idx * 0x9E3779B97F4A7C15ull);
}

/* ---------------------------------------------------------------- genome --
 * The genome is cut into segments of 2^16 bases.  Each segment draws its
 * 2048 words of 32 bases from a "content id".  Most segments own a unique
 * content; with dupMode != 0
 *   - one pair of neighbouring segments in 32 shares one content  (-> copy 2)
 *   - one segment in 64 is taken from a pool of 8 repeat families (-> copy M)
 * which gives modmap's copy1/copy2/multi classes (reference modmap.c:125-129)
 * non-trivial populations.
 */
#define MG_SEG_SHIFT 16
#define MG_SEG_MASK  0xFFFFull

MG_HD uint64_t mg_genome_content(uint64_t seed, uint64_t seg, int dupMode)
{
  if (dupMode)
    { uint64_t r = mg_rng(seed, 1, seg);
      if ((r & 63) == 0) return (r >> 8) & 7;                 /* repeat family 0..7 */
      uint64_t pr = mg_rng(seed, 2, seg >> 1);
      if ((pr & 31) == 0) return 8 + ((seg >> 1) << 1);       /* both halves of the pair */
    }
  return 8 + seg;
}

MG_HD uint64_t mg_genome_word(uint64_t seed, uint64_t g32 /* = g >> 5 */, int dupMode)
{
  uint64_t seg = g32 >> (MG_SEG_SHIFT - 5);
  uint64_t c = mg_genome_content(seed, seg, dupMode);
  return mg_rng(seed, 3, (c << (MG_SEG_SHIFT - 5)) | (g32 & (MG_SEG_MASK >> 5)));
}

MG_HD uint8_t mg_genome_base(uint64_t seed, uint64_t g, int dupMode)
{
  uint64_t w = mg_genome_word(seed, g >> 5, dupMode);
  return (uint8_t)((w >> (2 * (g & 31))) & 3);
}

/* ----------------------------------------------------------------- reads --
 * Fixed-length reads sampled from a genome.
 *   pairMode 0: read r starts uniformly in [0, G-L], strand = one random bit;
 *               reverse strand = reverse of (3-b)              (C1, C3)
 *   pairMode 1: reads 2f, 2f+1 are the two ends of fragment f of length
 *               fragLen, mate 1 being the reverse complement of the far end (C4)
 * Substitution errors: position j of read r is in error iff
 *   e % 1000000 < subPPM, and is replaced by (b + 1 + (e>>32) % 3) & 3.
 */
typedef struct {
  uint64_t genomeSeed;
  uint64_t genomeLen;
  uint64_t readSeed;
  uint32_t readLen;
  uint32_t subPPM;      /* substitutions per million bases */
  uint32_t insPPM;      /* insertions per million (ONT mode only) */
  uint32_t delPPM;      /* deletions per million (ONT mode only) */
  uint32_t fragLen;     /* pairMode 1 */
  int32_t  pairMode;
  int32_t  dupMode;
  int32_t  pad_;
} MgReadSpec;

MG_HD void mg_read_layout(const MgReadSpec *sp, uint64_t r, uint64_t *start, int *rev)
{
  if (sp->pairMode)
    { uint64_t f = r >> 1;
      uint64_t x = mg_rng(sp->readSeed, 10, f);
      uint64_t span = sp->genomeLen - sp->fragLen + 1;
      uint64_t fs = (x >> 1) % span;
      int frev = (int)(x & 1);
      int mate = (int)(r & 1);
      /* mate 0 reads into the fragment from its 5' end on the fragment strand,
         mate 1 from the other end on the opposite strand */
      *rev = mate ^ frev;
      *start = *rev ? (fs + sp->fragLen - sp->readLen) : fs;
    }
  else
    { uint64_t x = mg_rng(sp->readSeed, 10, r);
      uint64_t span = sp->genomeLen - sp->readLen + 1;
      *start = (x >> 1) % span;
      *rev = (int)(x & 1);
    }
}

/* base j of (substitution-only) read r; start/rev from mg_read_layout */
MG_HD uint8_t mg_read_base(const MgReadSpec *sp, uint64_t r, uint32_t j, uint64_t start, int rev)
{
  uint64_t g = rev ? (start + sp->readLen - 1 - j) : (start + j);
  uint8_t b = mg_genome_base(sp->genomeSeed, g, sp->dupMode);
  if (rev) b = (uint8_t)(3 - b);
  if (sp->subPPM)
    { uint64_t e = mg_rng(sp->readSeed, 11, r * (uint64_t)sp->readLen + j);
      if ((e & 0xFFFFFFFFull) % 1000000u < sp->subPPM) b = (uint8_t)((b + 1 + (e >> 32) % 3) & 3);
    }
  return b;
}

/* ONT-like read with substitutions, insertions and deletions: a sequential
 * walk along the template, emitting exactly readLen bases.  The template
 * window is 2*readLen long so that deletions cannot run off the genome. */
MG_HD void mg_ont_read(const MgReadSpec *sp, uint64_t r, uint8_t *out)
{
  uint64_t x = mg_rng(sp->readSeed, 10, r);
  uint64_t span = sp->genomeLen - 2ull * sp->readLen + 1;
  uint64_t start = (x >> 1) % span;
  int rev = (int)(x & 1);
  uint64_t tlen = 2ull * sp->readLen;
  uint64_t t = 0;                    /* template cursor */
  uint64_t step = 0;                 /* decision counter */
  uint32_t n = 0;
  uint32_t tSub = sp->subPPM, tIns = tSub + sp->insPPM, tDel = tIns + sp->delPPM;
  while (n < sp->readLen)
    { uint64_t e = mg_rng(sp->readSeed, 12, r * 4ull * sp->readLen + step); ++step;
      uint32_t u = (uint32_t)((e & 0xFFFFFFFFull) % 1000000u);
      if (u >= tIns && u < tDel && t + 1 < tlen) { ++t; continue; }      /* deletion */
      if (u >= tSub && u < tIns) { out[n++] = (uint8_t)((e >> 32) & 3); continue; }  /* insertion */
      uint64_t tt = (t < tlen) ? t : tlen - 1; ++t;
      uint64_t g = rev ? (start + tlen - 1 - tt) : (start + tt);
      uint8_t b = mg_genome_base(sp->genomeSeed, g, sp->dupMode);
      if (rev) b = (uint8_t)(3 - b);
      if (u < tSub) b = (uint8_t)((b + 1 + (e >> 32) % 3) & 3);          /* substitution */
      out[n++] = b;
    }
}

#endif /* MODGPU_SYNTH_H */
