// mg_common.cuh - shared types and arithmetic of libmodgpu (sm_100a).
//
// The arithmetic in this header is __host__ __device__ on purpose: tests/ compiles
// it with the host compiler and checks it against the oracle position by
// position (tests/test_math_host.py), so that what the kernels compute per base
// is verified even on a box without a GPU.  The product never runs it on the
// CPU: every public entry point launches kernels or fails.
#pragma once

#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define MGHD __host__ __device__ __forceinline__
#define MGD __device__ __forceinline__
#else
#define MGHD static inline
#define MGD static inline
#endif

// ----------------------------------------------------------------- geometry
// One thread owns a RUN of 32 consecutive window starts (= one packed word),
// one warp 1024, one 256-thread tile 8192 bases = 2 KiB of packed sequence.
#define MG_RUN 32
#define MG_TILE_THREADS 256
#define MG_TILE_BASES (MG_RUN * MG_TILE_THREADS)
// slack words readable past the last tile (thread t also reads word t+1)
#define MG_PACK_SLACK_WORDS 8

// ------------------------------------------------------------ kernel hasher
// Device view of the reference Seqhash (seqhash.h:15-23) plus precomputed
// constants of the divisibility test and of the power-of-two prefilter.
struct MgKHasher {
  uint64_t factor;      // factor1
  uint64_t mask;        // 2k low bits
  uint64_t oddInv;      // inverse of the odd part of d modulo 2^64
  uint64_t oddLim;      // floor((2^64-1) / odd part)
  uint32_t k;
  uint32_t shift;       // 64 - 2k
  uint32_t d;           // modulus (reference w)
  uint32_t tz;          // d = 2^tz * odd
  uint32_t prefilter;   // 1: low-word candidate filter usable (tz >= 3, shift + tz <= 32)
  uint32_t pfMul;       // (uint32) factor << (32 - shift - tz)
  uint32_t pfLim;       // 1 << (32 - tz)
  uint32_t lut;         // 1: the prefilter depends on <= 4 bases per strand (shift + tz <= 8, k >= 30): table-driven scan
};

MGHD uint64_t mg_mulinv64(uint64_t a)   // a odd: Newton iteration, 5 steps double the bits 5->64
{
  uint64_t x = a;                       // correct to 3 bits (a*a = 1 mod 8)
  for (int i = 0; i < 6; ++i) x *= 2 - a * x;
  return x;
}

MGHD MgKHasher mg_make_khasher(int k, int d, uint64_t factor1)
{
  MgKHasher H;
  H.factor = factor1;
  H.k = (uint32_t)k;
  H.shift = (uint32_t)(64 - 2 * k);
  H.mask = (((uint64_t)1) << (2 * k)) - 1;
  H.d = (uint32_t)d;
  uint32_t tz = 0, odd = (uint32_t)d;
  while (!(odd & 1)) { odd >>= 1; ++tz; }
  H.tz = tz;
  H.oddInv = mg_mulinv64(odd);
  H.oddLim = 0xFFFFFFFFFFFFFFFFull / odd;
  H.prefilter = (tz >= 3 && H.shift + tz <= 32) ? 1u : 0u;
  H.pfMul = H.prefilter ? (uint32_t)(factor1 << (32 - H.shift - tz)) : 0u;
  H.pfLim = H.prefilter ? (1u << (32 - tz)) : 0u;
  H.lut = (H.prefilter && H.shift + tz <= 8 && k >= 30) ? 1u : 0u;
  return H;
}

// seqhash() of the reference (seqhash.h:58)
MGHD uint64_t mg_hash(const MgKHasher &H, uint64_t kmer) { return (kmer * H.factor) >> H.shift; }

// hash % d == 0 without a division (the reference spends most of its 10 ns/base
// in this `%`, seqhash.c:171,190): low tz bits zero, and exact-division test
// for the odd part: n divisible by odd  <=>  n * inv(odd) mod 2^64 <= (2^64-1)/odd
MGHD bool mg_divisible(const MgKHasher &H, uint64_t hash)
{
  if (hash & ((1ull << H.tz) - 1)) return false;
  return (hash >> H.tz) * H.oddInv <= H.oddLim;
}

// canonical choice of hashRC (seqhash.c:60-68): forward only if strictly smaller
MGHD bool mg_canonical_select(const MgKHasher &H, uint64_t fwd, uint64_t rc, bool *isF)
{
  uint64_t hf = mg_hash(H, fwd), hr = mg_hash(H, rc);
  bool f = hf < hr;
  *isF = f;
  return mg_divisible(H, f ? hf : hr);
}

// ------------------------------------------------------------- 2-bit windows
// A run's view of the packed stream: two consecutive words w0:w1 = 64 bases,
// base j at bits [126-2j, 127-2j] of the 128-bit value (first base on top).
// The reverse-complement stream keeps base j at bits [2j, 2j+1], complemented,
// so that the rc k-mer of window i is again a plain bit field (seqhash.c:74:
// rc rolls in at the top, i.e. base p+m of the window sits at bits 2m).
struct MgRun {
  uint64_t yhi, ylo;    // (w0:w1) >> (64 - 2k): forward k-mer i = (y >> (64-2i)) & mask
  uint64_t rhi, rlo;    // ~pairreverse(w0:w1):  rc k-mer i      = (r >> 2i) & mask
};

MGHD uint64_t mg_pairrev64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
  uint64_t y = __brevll(x);
#else
  uint64_t y = x;
  y = ((y >> 1) & 0x5555555555555555ull) | ((y & 0x5555555555555555ull) << 1);
  y = ((y >> 2) & 0x3333333333333333ull) | ((y & 0x3333333333333333ull) << 2);
  y = ((y >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((y & 0x0F0F0F0F0F0F0F0Full) << 4);
  y = ((y >> 8) & 0x00FF00FF00FF00FFull) | ((y & 0x00FF00FF00FF00FFull) << 8);
  y = ((y >> 16) & 0x0000FFFF0000FFFFull) | ((y & 0x0000FFFF0000FFFFull) << 16);
  y = (y >> 32) | (y << 32);
#endif
  // full bit reversal also swapped the two bits inside every base: swap back
  return ((y >> 1) & 0x5555555555555555ull) | ((y & 0x5555555555555555ull) << 1);
}

MGHD MgRun mg_run_prepare(uint64_t w0, uint64_t w1, uint32_t k)
{
  MgRun R;
  uint32_t s = 64 - 2 * k;                   // 2..62
  R.yhi = w0 >> s;
  R.ylo = (w0 << (64 - s)) | (w1 >> s);
  R.rlo = ~mg_pairrev64(w0);
  R.rhi = ~mg_pairrev64(w1);
  return R;
}

// forward / reverse-complement k-mer of window i (0..31) of a run
MGHD uint64_t mg_run_fwd(const MgRun &R, uint32_t i, uint64_t mask)
{
  uint64_t v = i ? ((R.yhi << (2 * i)) | (R.ylo >> (64 - 2 * i))) : R.yhi;
  return v & mask;
}

MGHD uint64_t mg_run_rc(const MgRun &R, uint32_t i, uint64_t mask)
{
  uint64_t v = i ? ((R.rlo >> (2 * i)) | (R.rhi << (64 - 2 * i))) : R.rlo;
  return v & mask;
}

// Window starts of a run that may NOT be used: a k-mer must not span two
// sequences (modRCiterator is per sequence, seqhash.c:154-177).  `ends` = end
// flags of the 64 bases starting at the run (bit j = base j is the last base of
// a sequence).  Start i is blocked iff a flag sits on one of the window's first
// k-1 bases, i.e. bits i .. i+k-2.  Returns the 32-bit blocked mask.
MGHD uint32_t mg_blocked_mask(uint64_t ends, uint32_t k)
{
  uint32_t L = k - 1;                       // window of flags to OR, 0..30
  if (L == 0 || ends == 0) return 0u;       // no sequence ends in sight (the common case)
  uint64_t s = ends;                        // OR over 1 flag
  uint32_t have = 1;
  while (have * 2 <= L) { s |= s >> have; have *= 2; }   // OR over `have` flags, have = 2^m <= L
  s |= s >> (L - have);                     // OR over exactly L flags
  return (uint32_t)s;
}

// full canonical evaluation of window i of a run: hashRC (seqhash.c:60-68)
// followed by the `% w` of modRCiterator/modRCnext (seqhash.c:171,190)
MGHD bool mg_eval_window(const MgKHasher &H, const MgRun &R, uint32_t i, uint64_t *kmer, bool *isF)
{
  // The hashes are never shifted into place: h = P >> shift is monotone in the masked product P & ~(2^shift - 1)
  // = h << shift, so the canonical choice compares masked products; "the tz low bits of h are zero" is one AND
  // with a constant; and the odd part of d divides h exactly when it divides h << shift (it is coprime to 2), which
  // the exact-division test answers on the masked product itself.  9 instructions fewer per window than
  // shift - compare - shift - multiply (the full scan of a generic d evaluates every window).
  const uint64_t f = mg_run_fwd(R, i, H.mask), r = mg_run_rc(R, i, H.mask);
  const uint64_t keep = ~((((uint64_t)1) << H.shift) - 1);
  const uint64_t pf = (f * H.factor) & keep, pr = (r * H.factor) & keep;
  const bool fw = pf < pr;                                     // hashF < hashR (ties go reverse, seqhash.c:66-67)
  const uint64_t pm = fw ? pf : pr;
  *kmer = fw ? f : r;
  *isF = fw;
  const bool lowOk = (pm & (((((uint64_t)1) << H.tz) - 1) << H.shift)) == 0;
  const bool oddOk = pm * H.oddInv <= H.oddLim;
  return lowOk && oddOk;
}

// ---- the full scan of a generic d (tz < 3: odd d such as the reference's default 31, 2*odd, 4*odd) evaluates EVERY
// window, so its instruction count is the kernel's.  For k >= 16 (shift <= 32) the same test in explicit 32-bit
// pieces: the k-mer's low word needs no mask (2k >= 32), the masked product keeps its whole high word, one funnel
// shift per k-mer word, three multiply-adds per 64-bit product, and for an odd d no low-bit test at all.
// 24-26 instructions per window instead of 38 (ncu prof_r01_generic).  Same result as mg_eval_window (checked
// position by position on the host, tests/test_math_host.py).
struct MgRun32 { uint32_t y[4], r[4]; };   // y[3]:y[2] = R.yhi, y[1]:y[0] = R.ylo; r[1]:r[0] = R.rlo, r[3]:r[2] = R.rhi

MGHD MgRun32 mg_run32(const MgRun &R)
{
  MgRun32 Q;
  Q.y[3] = (uint32_t)(R.yhi >> 32); Q.y[2] = (uint32_t)R.yhi; Q.y[1] = (uint32_t)(R.ylo >> 32); Q.y[0] = (uint32_t)R.ylo;
  Q.r[0] = (uint32_t)R.rlo; Q.r[1] = (uint32_t)(R.rlo >> 32); Q.r[2] = (uint32_t)R.rhi; Q.r[3] = (uint32_t)(R.rhi >> 32);
  return Q;
}

MGHD uint32_t mg_fl32(uint32_t lo, uint32_t hi, uint32_t s)   // high word of (hi:lo) << s, 0 <= s < 32
{
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(lo, hi, s);
#else
  return s ? ((hi << s) | (lo >> (32 - s))) : hi;
#endif
}

MGHD uint32_t mg_fr32(uint32_t lo, uint32_t hi, uint32_t s)   // low word of (hi:lo) >> s, 0 <= s < 32
{
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, s);
#else
  return s ? ((lo >> s) | (hi << (32 - s))) : lo;
#endif
}

// constants of the 32-bit evaluation (kernel-uniform; the compiler keeps them in registers across the unrolled scan)
struct MgEval32 { uint32_t fLo, fHi, maskHi, keepLo, lowLo, lowHi, invLo, invHi, limLo, limHi; };

MGHD MgEval32 mg_eval32_prepare(const MgKHasher &H)             // requires H.shift <= 32
{
  MgEval32 E;
  E.fLo = (uint32_t)H.factor; E.fHi = (uint32_t)(H.factor >> 32);
  E.maskHi = (uint32_t)(H.mask >> 32);
  const uint64_t keep = ~((((uint64_t)1) << H.shift) - 1);
  E.keepLo = (uint32_t)keep;
  const uint64_t low = ((((uint64_t)1) << H.tz) - 1) << H.shift;
  E.lowLo = (uint32_t)low; E.lowHi = (uint32_t)(low >> 32);
  E.invLo = (uint32_t)H.oddInv; E.invHi = (uint32_t)(H.oddInv >> 32);
  E.limLo = (uint32_t)H.oddLim; E.limHi = (uint32_t)(H.oddLim >> 32);
  return E;
}

// low 64 bits of (ah:al) * (bh:bl) as two words: one wide multiply and two multiply-adds into its high word.
// Spelled out in PTX on the device: from the C expression the compiler built 64-bit additions of zero-extended
// halves, four extra ALU-pipe instructions per product in a loop that is ALU-pipe bound.
MGHD void mg_mul64lo(uint32_t al, uint32_t ah, uint32_t bl, uint32_t bh, uint32_t *lo, uint32_t *hi)
{
#if defined(__CUDA_ARCH__)
  uint32_t l, h;
  asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %4;\n\tmov.b64 {%0, %1}, t;\n\t"
      "mad.lo.u32 %1, %2, %5, %1;\n\tmad.lo.u32 %1, %3, %4, %1;\n\t}"
      : "=&r"(l), "=&r"(h) : "r"(al), "r"(ah), "r"(bl), "r"(bh));
  *lo = l; *hi = h;
#else
  const uint64_t w = (uint64_t)al * bl;
  *lo = (uint32_t)w;
  *hi = (uint32_t)(w >> 32) + al * bh + ah * bl;
#endif
}

template <bool ODD>
MGHD bool mg_selected32(const MgEval32 &E, const MgRun32 &Q, uint32_t i)
{
  const uint32_t s = (2 * i) & 31u;
  const uint32_t o = (i < 16) ? 1u : 0u, p = 1u - o;
  const uint32_t fl = mg_fl32(Q.y[o], Q.y[o + 1], s), fh = mg_fl32(Q.y[o + 1], Q.y[o + 2], s) & E.maskHi;
  const uint32_t rl = mg_fr32(Q.r[p], Q.r[p + 1], s), rh = mg_fr32(Q.r[p + 1], Q.r[p + 2], s) & E.maskHi;
  // (branch-free on purpose: 64-bit compares are two ISETP, the short-circuit forms compiled to branches)
  uint32_t pfl, pfh, prl, prh;
  mg_mul64lo(fl, fh, E.fLo, E.fHi, &pfl, &pfh);
  mg_mul64lo(rl, rh, E.fLo, E.fHi, &prl, &prh);
  // only the DECISION is wanted here, and it depends on min(hashF, hashR) alone: the smaller whole product has the
  // smaller-or-equal hash, and on a tie of the hashes either product carries that hash - so the products are compared
  // unmasked and the bits below the hash are dropped once, after the choice (the strand, where ties matter -
  // seqhash.c:66-67 - is decided exactly by mg_eval_single when the window is extracted)
  const uint64_t pf = ((uint64_t)pfh << 32) | pfl, pr = ((uint64_t)prh << 32) | prl;
  const uint64_t pm = (pf < pr) ? pf : pr;
  const uint32_t ml = (uint32_t)pm & E.keepLo, mh = (uint32_t)(pm >> 32);
  uint32_t ql, qh;
  mg_mul64lo(ml, mh, E.invLo, E.invHi, &ql, &qh);
  const uint64_t q = ((uint64_t)qh << 32) | ql, lim = ((uint64_t)E.limHi << 32) | E.limLo;
  bool ok = q <= lim;
  if (!ODD) ok = ok & (((ml & E.lowLo) | (mh & E.lowHi)) == 0u);
  return ok;
}

// The same decision for TWO windows at once, i and i + 16 (i < 16), as a SUPERSET of the selected windows (the count
// kernel re-evaluates every flagged window exactly, mg_eval32_single):
//  - windows i and i + 16 lie exactly one 32-bit word apart, so the middle funnel shift of each strand serves both
//    (the low word of one window is the unmasked high word of the other): three shifts per strand for two windows;
//  - the exact-division test compares the HIGH product word only (q <= lim iff qh < limHi, or qh == limHi and
//    ql <= limLo: qh <= limHi can only add windows), which needs the high half of ONE partial product instead of a
//    64-bit compare.
template <bool ODD>
MGHD void mg_selected32_pair(const MgEval32 &E, const MgRun32 &Q, uint32_t i, bool *sel0, bool *sel16)
{
  const uint32_t s = 2 * i;                                   // 0 .. 30
  const uint32_t fA = mg_fl32(Q.y[0], Q.y[1], s), fB = mg_fl32(Q.y[1], Q.y[2], s), fC = mg_fl32(Q.y[2], Q.y[3], s);
  const uint32_t rA = mg_fr32(Q.r[0], Q.r[1], s), rB = mg_fr32(Q.r[1], Q.r[2], s), rC = mg_fr32(Q.r[2], Q.r[3], s);
  // window i: forward (fB, fC & m), reverse (rA, rB & m); window i + 16: forward (fA, fB & m), reverse (rB, rC & m)
  for (int h = 0; h < 2; ++h)
    { const uint32_t fl = h ? fA : fB, fh = (h ? fB : fC) & E.maskHi;
      const uint32_t rl = h ? rB : rA, rh = (h ? rC : rB) & E.maskHi;
      uint32_t pfl, pfh, prl, prh;
      mg_mul64lo(fl, fh, E.fLo, E.fHi, &pfl, &pfh);
      mg_mul64lo(rl, rh, E.fLo, E.fHi, &prl, &prh);
      const uint64_t pf = ((uint64_t)pfh << 32) | pfl, pr = ((uint64_t)prh << 32) | prl;
      const uint64_t pm = (pf < pr) ? pf : pr;                // the smaller product carries the smaller hash (see mg_selected32)
      const uint32_t ml = (uint32_t)pm & E.keepLo, mh = (uint32_t)(pm >> 32);
#if defined(__CUDA_ARCH__)
      const uint32_t qh = __umulhi(ml, E.invLo) + ml * E.invHi + mh * E.invLo;
#else
      const uint32_t qh = (uint32_t)(((uint64_t)ml * E.invLo) >> 32) + ml * E.invHi + mh * E.invLo;
#endif
      bool ok = qh <= E.limHi;
      if (!ODD) ok = ok & (((ml & E.lowLo) | (mh & E.lowHi)) == 0u);
      if (h) *sel16 = ok; else *sel0 = ok;
    }
}

// The same for ONE window given the run's two packed words (phase 3 of the
// kernel, where only the few queued windows are evaluated): the forward k-mer
// is a bit field of w0:w1 and the reverse complement is computed from the k-mer
// itself (reverse the 2-bit groups, complement, realign) instead of preparing
// the whole reverse-complement stream.
MGHD bool mg_eval_single(const MgKHasher &H, uint64_t w0, uint64_t w1, uint32_t i, uint64_t *kmer, bool *isF)
{
  const uint32_t sh = 128 - 2 * (i + H.k);              // 4..126: bits of w0:w1 below the window
  uint64_t f = (sh >= 64) ? (w0 >> (sh - 64)) : ((w0 << (64 - sh)) | (w1 >> sh));
  f &= H.mask;
  const uint64_t r = (~mg_pairrev64(f)) >> H.shift;
  // masked products instead of shifted hashes, as in mg_eval_window
  const uint64_t keep = ~((((uint64_t)1) << H.shift) - 1);
  const uint64_t pf = (f * H.factor) & keep, pr = (r * H.factor) & keep;
  const bool fw = pf < pr;
  const uint64_t pm = fw ? pf : pr;
  *kmer = fw ? f : r;
  *isF = fw;
  const bool lowOk = (pm & (((((uint64_t)1) << H.tz) - 1) << H.shift)) == 0;
  const bool oddOk = (H.oddInv == 1) || (pm * H.oddInv <= H.oddLim);    // odd part 1: d is a power of two
  return lowOk && oddOk;
}

// The same for the table-driven scan's parameter range (K = 30 or 31 known at compile time, d a power of two
// with 64-2K+tz <= 8): the shifts are constants, the two hashes are compared as masked products (h = P >> S is
// monotone in P & ~(2^S - 1)), "hash % d == 0" is one AND on the low product word, and no odd-part test is needed.
template <int K>
MGHD bool mg_eval_single_pow2(const MgKHasher &H, uint64_t w0, uint64_t w1, uint32_t i, uint64_t *kmer, bool *isF)
{
  constexpr uint32_t S = 64 - 2 * K;
  constexpr uint64_t MASK = (((uint64_t)1) << (2 * K)) - 1;
  const uint32_t sl = 2 * i;                                   // the window starts 2i bits below the top of w0:w1
  const uint64_t top = sl ? ((w0 << sl) | (w1 >> (64 - sl))) : w0;
  const uint64_t f = (top >> S) & MASK;
  const uint64_t r = (~mg_pairrev64(f)) >> S;
  const uint64_t pf = (f * H.factor) & ~(uint64_t)((1u << S) - 1u), pr = (r * H.factor) & ~(uint64_t)((1u << S) - 1u);
  const bool fw = pf < pr;                                     // hashF < hashR (ties go reverse, seqhash.c:66-67)
  const uint64_t pm = fw ? pf : pr;
  *kmer = fw ? f : r;
  *isF = fw;
  return ((uint32_t)pm & (((1u << H.tz) - 1u) << S)) == 0u;    // the tz low hash bits are zero
}

// ---- one window in 32-bit pieces (k >= 16): what the second-generation count kernel evaluates per queue entry.
// Same result as mg_eval_single (canonical k-mer, strand with ties going reverse, hash % d == 0), branch-free:
// the window is a funnel-shifted field of the run's two packed words held as four 32-bit halves, the reverse
// complement comes from two BREVs, the two 64-bit products are one wide multiply and two multiply-adds each, and
// the tests run on the masked product as in mg_eval_window.  POW2: d is a power of two (no odd-part test).
MGHD uint32_t mg_frc32(uint32_t lo, uint32_t hi, uint32_t s)   // low word of (hi:lo) >> s, 0 <= s <= 32 (32 gives hi)
{
#if defined(__CUDA_ARCH__)
  return __funnelshift_rc(lo, hi, s);
#else
  return s >= 32 ? hi : (s ? ((lo >> s) | (hi << (32 - s))) : lo);
#endif
}

MGHD uint32_t mg_pairrev32(uint32_t x)                          // reverse the order of the sixteen 2-bit groups
{
#if defined(__CUDA_ARCH__)
  const uint32_t y = __brev(x);
#else
  uint32_t y = x;
  y = ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
  y = ((y >> 2) & 0x33333333u) | ((y & 0x33333333u) << 2);
  y = ((y >> 4) & 0x0F0F0F0Fu) | ((y & 0x0F0F0F0Fu) << 4);
  y = ((y >> 8) & 0x00FF00FFu) | ((y & 0x00FF00FFu) << 8);
  y = (y >> 16) | (y << 16);
#endif
  return ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
}

// ~mg_pairrev32(x): the complement of the reversed groups (the reverse complement of sixteen bases).  On the device
// the swap of the bits of every pair and the complement are ONE three-input logic operation on (y >> 1, y << 1,
// 0x55555555): from C the compiler emits and / or-and / not
MGHD uint32_t mg_pairrev32_not(uint32_t x)
{
#if defined(__CUDA_ARCH__)
  const uint32_t y = __brev(x);
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, 0x55555555, 0x1B;" : "=r"(r) : "r"(y >> 1), "r"(y << 1));      // ~((a & c) | (b & ~c))
  return r;
#else
  return ~mg_pairrev32(x);
#endif
}

template <bool POW2>
MGHD bool mg_eval32_single(const MgEval32 &E, uint32_t shift, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t i,
                           uint32_t *kmLo, uint32_t *kmHi, bool *isF)
{ // the run's words w0 = a:b, w1 = c:d; the window starts 2i bits below the top of w0
  const uint32_t sl = 2 * i;
  const bool up = sl >= 32;
  const uint32_t x0 = up ? b : a, x1 = up ? c : b, x2 = up ? d : c;
  const uint32_t tHi = mg_fl32(x1, x0, sl & 31u), tLo = mg_fl32(x2, x1, sl & 31u);     // top 64 bits of (w0:w1) << 2i
  const uint32_t fLo = mg_frc32(tLo, tHi, shift), fHi = mg_frc32(tHi, 0u, shift);     // >> (64 - 2k): the forward k-mer
  const uint32_t qHi = mg_pairrev32_not(fLo), qLo = mg_pairrev32_not(fHi);                  // complement of the reversed k-mer, at the top
  const uint32_t rLo = mg_frc32(qLo, qHi, shift), rHi = mg_frc32(qHi, 0u, shift);
  uint32_t pfl, pfh, prl, prh;
  mg_mul64lo(fLo, fHi, E.fLo, E.fHi, &pfl, &pfh);
  mg_mul64lo(rLo, rHi, E.fLo, E.fHi, &prl, &prh);
  pfl &= E.keepLo; prl &= E.keepLo;                           // the hashes, still shifted up: compared as masked products
  const bool fw = (pfh < prh) || (pfh == prh && pfl < prl);   // hashF < hashR; ties go reverse (seqhash.c:66-67)
  const uint32_t ml = fw ? pfl : prl, mh = fw ? pfh : prh;
  *kmLo = fw ? fLo : rLo; *kmHi = fw ? fHi : rHi; *isF = fw;
  bool ok = ((ml & E.lowLo) | (mh & E.lowHi)) == 0u;          // the tz low hash bits are zero
  if (!POW2)
    { uint32_t ql, qh;
      mg_mul64lo(ml, mh, E.invLo, E.invHi, &ql, &qh);         // exact division by the odd part
      ok = ok && ((qh < E.limHi) || (qh == E.limHi && ql <= E.limLo));
    }
  return ok;
}

// power-of-two prefilter (H.prefilter): a window can only be selected if the tz
// low bits of hash(fwd) or of hash(rc) are zero, and with 64-2k+tz <= 32 those
// bits live in the LOW 32-bit word of the product, which depends only on the low
// word of the k-mer: (lo * factor) >> shift & (2^tz - 1) == 0
//   <=>  lo * (factor << (32-shift-tz))  <  2^(32-tz)        (mod 2^32)
// low 32 bits of the forward / reverse-complement window i, one funnel shift each
MGHD uint32_t mg_funnel_l(uint32_t lo, uint32_t hi, uint32_t s)   // high word of (hi:lo) << s, 0 < s < 32
{
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(lo, hi, s);
#else
  return (hi << s) | (lo >> (32 - s));
#endif
}

MGHD uint32_t mg_funnel_r(uint32_t lo, uint32_t hi, uint32_t s)   // low word of (hi:lo) >> s, 0 < s < 32
{
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, s);
#else
  return (lo >> s) | (hi << (32 - s));
#endif
}

MGHD uint32_t mg_run_fwd_lo(const MgRun &R, uint32_t i)
{
  const uint32_t yhl = (uint32_t)R.yhi, ylh = (uint32_t)(R.ylo >> 32), yll = (uint32_t)R.ylo;
  if (i == 0) return yhl;
  if (i < 16) return mg_funnel_l(ylh, yhl, 2 * i);
  if (i == 16) return ylh;
  return mg_funnel_r(yll, ylh, 64 - 2 * i);
}

MGHD uint32_t mg_run_rc_lo(const MgRun &R, uint32_t i)
{
  const uint32_t rll = (uint32_t)R.rlo, rlh = (uint32_t)(R.rlo >> 32), rhl = (uint32_t)R.rhi;
  if (i == 0) return rll;
  if (i < 16) return mg_funnel_r(rll, rlh, 2 * i);
  if (i == 16) return rlh;
  return mg_funnel_r(rlh, rhl, 2 * i - 32);
}

MGHD bool mg_prefilter_candidate(const MgKHasher &H, const MgRun &R, uint32_t i)
{
  const uint32_t pf = mg_run_fwd_lo(R, i) * H.pfMul, pr = mg_run_rc_lo(R, i) * H.pfMul;
  return (pf < pr ? pf : pr) < H.pfLim;
}

// ---- table-driven prefilter (H.lut): when the candidate test reads only the low byte of each
// strand's k-mer (64-2k+tz <= 8), "window p is a candidate" is a function of two groups of four
// bases: the LAST four of the window (forward k-mer, low byte = those bases) and the FIRST four
// (reverse complement: base p+m complemented at bits 2m).  A 16 KiB table indexed by 7 consecutive
// bases answers both questions for the 4 four-base groups starting at those bases with ONE
// shared-memory load: bit i = group i is a reverse-complement candidate, bit 4+i = a forward one.
// One lookup per 4 positions replaces 4 x (2 funnel shifts + 2 multiplies + min + compare + or).
#define MG_LUT_BITS 14
#define MG_LUT_SIZE (1u << MG_LUT_BITS)

MGHD uint32_t mg_lut_entry(const MgKHasher &H, uint32_t x)    // x = 7 bases, the first in the top two bits
{
  uint32_t e = 0;
  for (int i = 0; i < 4; ++i)
    { const uint32_t B = (x >> (6 - 2 * i)) & 0xFFu;          // bases i..i+3, the low byte of a forward k-mer ending here
      uint32_t RB = 0;                                        // the low byte of a reverse-complement k-mer starting here
      for (int m = 0; m < 4; ++m) RB |= (3u - ((B >> (6 - 2 * m)) & 3u)) << (2 * m);
      if (B * H.pfMul < H.pfLim) e |= 1u << (4 + i);
      if (RB * H.pfMul < H.pfLim) e |= 1u << i;
    }
  return e;
}

// candidate masks of two consecutive runs (64 window starts) from the 96 bases w0:w1:w2;
// lut = the table above (shared memory on the device).  K = 30 or 31.
template <int K>
MGHD void mg_lut_scan(const uint8_t *lut, uint64_t w0, uint64_t w1, uint64_t w2, uint32_t *candLo, uint32_t *candHi)
{
  const uint32_t S[6] = { (uint32_t)(w0 >> 32), (uint32_t)w0, (uint32_t)(w1 >> 32), (uint32_t)w1, (uint32_t)(w2 >> 32), (uint32_t)w2 };
  uint32_t e[23];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < 23; ++j)                                // group j = bases 4j .. 4j+6
    { const int r = j >> 2, c = j & 3;
      uint32_t x;
      if (c == 0) x = S[r] >> 18;
      else if (c == 1) x = (S[r] >> 10) & 0x3FFFu;
      else if (c == 2) x = (S[r] >> 2) & 0x3FFFu;
      else x = mg_funnel_r(S[r + 1], S[r], 26) & 0x3FFFu;
      e[j] = lut[x];
    }
  // reverse-complement bits of position q -> bit q of (rLo, rHi)
  uint32_t rLo = 0, rHi = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 7; j >= 0; --j) { rLo = (rLo << 4) | (e[j] & 15u); rHi = (rHi << 4) | (e[j + 8] & 15u); }
  // forward bits of the four-base group at q belong to the window starting at q - OFF; three chains of
  // seven lookups keep the high nibble in place ((chain << 4) | (e & 0xF0) = the bits shifted left by 4)
  constexpr int OFF = K - 4, J0 = OFF / 4;
  uint32_t fa = 0, fb = 0, fc = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 6; j >= 0; --j)
    { fa = (fa << 4) | (e[J0 + j] & 0xF0u);
      fb = (fb << 4) | (e[J0 + 7 + j] & 0xF0u);
      if (J0 + 14 + j <= 22) fc = (fc << 4) | (e[J0 + 14 + j] & 0xF0u);
    }
  // chain a holds q = 4 J0 .. at bit q - 4 J0 + 4, i.e. window p at bit p + OFF - 4 J0 + 4
  constexpr int SA = OFF - 4 * J0 + 4, TB = 4 * J0 + 24 - OFF, TC = 4 * J0 + 52 - OFF;
  *candLo = rLo | (fa >> SA) | (fb << TB);
  *candHi = rHi | (fb >> (32 - TB)) | (fc << (TC - 32));
}

// window starts of the run at global offset p0 that may be selected at all
MGHD uint32_t mg_run_usable(uint64_t eflags, uint32_t k, uint64_t p0, uint64_t nBases)
{
  uint32_t usable = ~mg_blocked_mask(eflags, k);
  if (p0 + MG_RUN > nBases)
    usable &= (p0 >= nBases) ? 0u : ((1u << (uint32_t)(nBases - p0)) - 1u);
  return usable;
}

// ---- K1 arithmetic: four bytes -> four 2-bit codes, first byte in the top two
// bits of the 8-bit result.  ASCII: ((c>>1)^(c>>2))&3 maps A,a->0 C,c->1 G,g->2
// T,t->3 and N,n->0 (dna2indexConv with the N patch, seqio.c:643-652, modutils.c:39)
MGHD uint32_t mg_pack4(uint32_t v, bool ascii)
{
  uint32_t c = ascii ? (((v >> 1) ^ (v >> 2)) & 0x03030303u) : (v & 0x03030303u);
  return (c * 0x40100401u) >> 24;      // gather: byte0 -> bits 7:6 ... byte3 -> bits 1:0
}

MGHD uint64_t mg_pack32(const uint32_t v[8], bool ascii)
{
  uint32_t hi = (mg_pack4(v[0], ascii) << 24) | (mg_pack4(v[1], ascii) << 16) | (mg_pack4(v[2], ascii) << 8) | mg_pack4(v[3], ascii);
  uint32_t lo = (mg_pack4(v[4], ascii) << 24) | (mg_pack4(v[5], ascii) << 16) | (mg_pack4(v[6], ascii) << 8) | mg_pack4(v[7], ascii);
  return ((uint64_t)hi << 32) | lo;
}

MGHD uint32_t mg_code_of(uint32_t ch, bool ascii) { return ascii ? (((ch >> 1) ^ (ch >> 2)) & 3u) : (ch & 3u); }

// ------------------------------------------------------------------- table
// 16-byte slot, one 32-byte sector holds two: a probe costs one sector.
struct __attribute__((aligned(16))) MgSlot {
  uint64_t key;        // k-mer (< 2^62) or MG_EMPTY
  uint32_t count;      // occurrences (reference depth, clamped to 65535 on export)
  uint32_t aux;        // see MG_AUX_*
};

#define MG_EMPTY 0xFFFFFFFFFFFFFFFFull
// aux word:  < MG_AUX_ORD : (dense index << 2) | copy class      (numbered entry)
//           >= MG_AUX_ORD : not numbered yet; MG_AUX_ORD + ordinal of the first
//                           occurrence in the current exact-order batch, or
//                           MG_AUX_FRESH when no ordinal was recorded
#define MG_AUX_ORD 0xC0000000u
#define MG_AUX_FRESH 0xFFFFFFFFu
#define MG_MAX_INDEX ((MG_AUX_ORD >> 2) - 1)

MGHD uint64_t mg_slot_hash(uint64_t kmer, uint32_t slotBits)
{
  return (kmer * 0x9E3779B97F4A7C15ull) >> (64 - slotBits);
}

// owner GPU of a k-mer in the hash-sharded multi-GPU table: an independent
// multiplicative hash mapped uniformly onto [0, nOwners)
MGHD uint32_t mg_owner(uint64_t kmer, uint32_t nOwners)
{
  const uint32_t h = (uint32_t)((kmer * 0xD6E8FEB86659FD93ull) >> 32);
#if defined(__CUDA_ARCH__)
  return __umulhi(h, nOwners);               // floor(h * nOwners / 2^32), always < nOwners
#else
  return (uint32_t)(((uint64_t)h * nOwners) >> 32);
#endif
}
