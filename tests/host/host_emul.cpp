// host_emul.cpp - TEST INFRASTRUCTURE: drives the __host__ __device__ arithmetic
// of modimizer_b200/csrc/mg_common.cuh on the CPU, run by run, exactly as the
// kernels pack.cu / hash_select.cu do per thread, so that the per-base math is
// checked against the oracle on boxes without a GPU (tests/test_math_host.py).
// It is compiled only by the test suite and is not part of libmodgpu.
#include <stdint.h>
#include <string.h>
#include <vector>
#include "../../modimizer_b200/csrc/mg_common.cuh"

static long long g_superset_extra = 0;        // windows the paired superset scan flags beyond the exact ones

extern "C" {

long long hm_superset_extra(void) { return g_superset_extra; }

void hm_khasher(int k, int d, uint64_t factor1, uint64_t out[8])
{
  MgKHasher H = mg_make_khasher(k, d, factor1);
  out[0] = H.mask; out[1] = H.shift; out[2] = H.tz; out[3] = H.oddInv; out[4] = H.oddLim;
  out[5] = H.prefilter; out[6] = H.pfMul; out[7] = H.pfLim;
}

int hm_divisible(int d, uint64_t hash)
{
  MgKHasher H = mg_make_khasher(19, d, 1);
  return mg_divisible(H, hash) ? 1 : 0;
}

// K1: same word layout and SWAR path as pack2bit_kernel
void hm_pack(const uint8_t *bytes, uint64_t nBases, int ascii, uint64_t *words, uint64_t nWords)
{
  for (uint64_t w = 0; w < nWords; ++w)
    { uint64_t b0 = w * 32, word = 0;
      if (b0 + 32 <= nBases)
        { uint32_t v[8];
          memcpy(v, bytes + b0, 32);
          word = mg_pack32(v, ascii != 0);
        }
      else if (b0 < nBases)
        for (uint64_t j = 0; b0 + j < nBases; ++j) word |= (uint64_t)mg_code_of(bytes[b0 + j], ascii != 0) << (62 - 2 * j);
      words[w] = word;
    }
}

// K2 per-thread logic over every run; usePrefilter: -1 = as the kernel decides
int64_t hm_select(int k, int d, uint64_t factor1, const uint8_t *bytes, uint64_t nBases, int ascii,
                  const uint64_t *offs, uint64_t nSeq, int usePrefilter,
                  uint64_t *kmer, uint32_t *gpos, uint8_t *isF, int64_t cap)
{
  MgKHasher H = mg_make_khasher(k, d, factor1);
  bool pf = usePrefilter < 0 ? (H.prefilter != 0) : (usePrefilter != 0 && H.prefilter);
  uint64_t nWords = (nBases + 31) / 32 + 2;
  std::vector<uint64_t> words(nWords, 0);
  std::vector<uint32_t> ends(nWords + 1, 0);
  hm_pack(bytes, nBases, ascii, words.data(), nWords);
  for (uint64_t r = 0; r < nSeq; ++r)
    if (offs[r + 1] > offs[r]) { uint64_t g = offs[r + 1] - 1; ends[g >> 5] |= 1u << (g & 31); }
  int64_t n = 0;
  for (uint64_t T = 0; T * 32 < nBases; ++T)
    { uint64_t eflags = (uint64_t)ends[T] | ((uint64_t)ends[T + 1] << 32);
      uint64_t p0 = T * 32;
      uint32_t usable = mg_run_usable(eflags, H.k, p0, nBases);
      MgRun R = mg_run_prepare(words[T], words[T + 1], H.k);
      uint32_t sel = 0;
      if (pf)
        { uint32_t cand = 0;
          for (uint32_t i = 0; i < 32; ++i) cand |= (mg_prefilter_candidate(H, R, i) ? 1u : 0u) << i;
          cand &= usable;
          for (uint32_t i = 0; i < 32; ++i)
            if (cand >> i & 1)
              { uint64_t km, km2; bool f, f2;
                bool a = mg_eval_window(H, R, i, &km, &f), b = mg_eval_single(H, words[T], words[T + 1], i, &km2, &f2);
                if (a != b || (a && (km != km2 || f != f2))) return -2000000 - (int64_t)(p0 + i);
                if (a) sel |= 1u << i;
              }
        }
      else
        { for (uint32_t i = 0; i < 32; ++i) { uint64_t km; bool f; if (mg_eval_window(H, R, i, &km, &f)) sel |= 1u << i; }
          if (H.shift <= 32)                 // the kernel's 32-bit evaluation of the full scan (k >= 16) must agree everywhere
            { const MgEval32 E = mg_eval32_prepare(H);
              const MgRun32 Q = mg_run32(R);
              uint32_t s32 = 0, s32b = 0;
              for (uint32_t i = 0; i < 32; ++i)
                { if (mg_selected32<false>(E, Q, i)) s32 |= 1u << i;
                  if (H.tz == 0 && mg_selected32<true>(E, Q, i)) s32b |= 1u << i;
                }
              if (s32 != sel || (H.tz == 0 && s32b != sel)) return -3000000 - (int64_t)p0;
              // the paired superset form of the second-generation kernel: never misses a selected window, and what it
              // adds must be rare (the high-word comparison of the exact-division test)
              uint32_t sp = 0, spb = 0;
              for (uint32_t i = 0; i < 16; ++i)
                { bool a, b;
                  mg_selected32_pair<false>(E, Q, i, &a, &b);
                  sp |= (a ? 1u : 0u) << i; sp |= (b ? 1u : 0u) << (i + 16);
                  if (H.tz == 0) { mg_selected32_pair<true>(E, Q, i, &a, &b); spb |= (a ? 1u : 0u) << i; spb |= (b ? 1u : 0u) << (i + 16); }
                }
              if ((sp & sel) != sel || (H.tz == 0 && (spb & sel) != sel)) return -6000000 - (int64_t)p0;
              if (sp != sel || (H.tz == 0 && spb != sel)) ++g_superset_extra;
            }
          sel &= usable;
        }
      if (H.shift <= 32)                   // the count kernel's per-entry evaluation in 32-bit pieces: every window, both forms
        { const MgEval32 E = mg_eval32_prepare(H);
          const uint64_t w0 = words[T], w1 = words[T + 1];
          for (uint32_t i = 0; i < 32; ++i)
            { uint64_t km; bool f; uint32_t kl, kh; bool f3;
              const bool a = mg_eval_single(H, w0, w1, i, &km, &f);
              const bool b3 = mg_eval32_single<false>(E, H.shift, (uint32_t)(w0 >> 32), (uint32_t)w0, (uint32_t)(w1 >> 32), (uint32_t)w1, i, &kl, &kh, &f3);
              if (a != b3 || km != (((uint64_t)kh << 32) | kl) || f != f3) return -4000000 - (int64_t)(p0 + i);
              if (H.oddInv == 1)
                { const bool b4 = mg_eval32_single<true>(E, H.shift, (uint32_t)(w0 >> 32), (uint32_t)w0, (uint32_t)(w1 >> 32), (uint32_t)w1, i, &kl, &kh, &f3);
                  if (a != b4 || km != (((uint64_t)kh << 32) | kl) || f != f3) return -5000000 - (int64_t)(p0 + i);
                }
            }
        }
      for (uint32_t i = 0; i < 32; ++i)
        if (sel >> i & 1)
          { uint64_t km; bool f;
            uint64_t km2; bool f2;
            bool ok1 = mg_eval_window(H, R, i, &km2, &f2);
            bool ok2 = mg_eval_single(H, words[T], words[T + 1], i, &km, &f);     // what the kernel's phase 3 uses
            if (!ok1 || !ok2 || km != km2 || f != f2) return -1000000 - (int64_t)(p0 + i);
            if (n < cap) { kmer[n] = km; gpos[n] = (uint32_t)(p0 + i); isF[n] = f ? 1 : 0; }
            ++n;
          }
    }
  return n;
}

// table-driven prefilter (mg_lut_scan) against the arithmetic prefilter, every window of the batch;
// returns the number of candidate-mask mismatches (-1: the table scan does not apply to these parameters)
int64_t hm_lut_check(int k, int d, uint64_t factor1, const uint8_t *bytes, uint64_t nBases, int ascii, uint64_t *nCand)
{
  MgKHasher H = mg_make_khasher(k, d, factor1);
  if (!H.lut) return -1;
  std::vector<uint8_t> lut(MG_LUT_SIZE);
  for (uint32_t x = 0; x < MG_LUT_SIZE; ++x) lut[x] = (uint8_t)mg_lut_entry(H, x);
  uint64_t nWords = (nBases + 31) / 32 + 4;
  std::vector<uint64_t> words(nWords, 0);
  hm_pack(bytes, nBases, ascii, words.data(), nWords);
  int64_t bad = 0;
  *nCand = 0;
  for (uint64_t T = 0; T * 32 < nBases; T += 2)
    { uint32_t lo, hi, c[2] = { 0, 0 };
      if (k == 31) mg_lut_scan<31>(lut.data(), words[T], words[T + 1], words[T + 2], &lo, &hi);
      else mg_lut_scan<30>(lut.data(), words[T], words[T + 1], words[T + 2], &lo, &hi);
      for (int h = 0; h < 2; ++h)
        { MgRun R = mg_run_prepare(words[T + h], words[T + h + 1], H.k);
          for (uint32_t i = 0; i < 32; ++i) c[h] |= (mg_prefilter_candidate(H, R, i) ? 1u : 0u) << i;
        }
      if (lo != c[0]) ++bad;
      if (hi != c[1]) ++bad;
      if (H.oddInv == 1)                                    // d a power of two: the specialised evaluation == the generic one
        for (int h = 0; h < 2; ++h)
          for (uint32_t i = 0; i < 32; ++i)
            if (c[h] >> i & 1)
              { uint64_t k1, k2; bool f1, f2;
                bool a = mg_eval_single(H, words[T + h], words[T + h + 1], i, &k1, &f1);
                bool b = (k == 31) ? mg_eval_single_pow2<31>(H, words[T + h], words[T + h + 1], i, &k2, &f2)
                                   : mg_eval_single_pow2<30>(H, words[T + h], words[T + h + 1], i, &k2, &f2);
                if (a != b || k1 != k2 || f1 != f2) ++bad;
              }
      *nCand += __builtin_popcount(c[0]) + __builtin_popcount(c[1]);
    }
  return bad;
}

uint64_t hm_slot_hash(uint64_t kmer, uint32_t bits) { return mg_slot_hash(kmer, bits); }
uint32_t hm_owner(uint64_t kmer, uint32_t n) { return mg_owner(kmer, n); }

}  // extern "C"

// ---- synthetic inputs on the host: the same header the device generators use
#include "../../include/modgpu_synth.h"
extern "C" {
void hs_genome(uint64_t seed, uint64_t start, uint64_t n, int dupMode, uint8_t *out)
{ for (uint64_t i = 0; i < n; ++i) out[i] = mg_genome_base(seed, start + i, dupMode); }

void hs_reads(const MgReadSpec *sp, uint64_t firstRead, uint64_t nReads, int ont, uint8_t *out)
{
  for (uint64_t rr = 0; rr < nReads; ++rr)
    { uint64_t r = firstRead + rr;
      if (ont) { mg_ont_read(sp, r, out + rr * sp->readLen); continue; }
      uint64_t start; int rev;
      mg_read_layout(sp, r, &start, &rev);
      for (uint32_t j = 0; j < sp->readLen; ++j) out[rr * sp->readLen + j] = mg_read_base(sp, r, j, start, rev);
    }
}
}
