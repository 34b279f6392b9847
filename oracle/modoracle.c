/* modoracle.c - CPU restatement of modimizer's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; the product (modimizer_b200/) never
 * does and has no CPU fallback.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors
 * (SURVEY.md section 4), so this restatement is pinned against the reference itself:
 * oracle/Makefile compiles the unmodified reference sources where they lie
 * into oracle/_ref/libmodref.so (+ the stock modutils/modmap CLIs) and
 * tests/test_oracle_vs_ref.py compares every function below with it; the
 * known-answer vectors in tests/golden/ were generated from that build by
 * tests/golden/make_golden.py.  factor1 depends on glibc random()
 * (seqhash.c:30-31); both checkers call the same libc, so same-box parity holds.
 *
 * Each function cites the reference file:line it restates.  The code is
 * written from the algorithm, not transcribed: one flat scanner instead of
 * the iterator object, explicit structs instead of macros.
 */
#define _DEFAULT_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#define HX(name) orc_##name
#include "harness_api.h"

static void orc_die(const char *msg, long a, long b)
{
  fprintf(stderr, "FATAL ERROR (oracle): %s (%ld, %ld)\n", msg, a, b);
  exit(-1);                                        /* reference: die(), utils.c:19-30 */
}

static void *xcalloc(size_t n, size_t sz)
{
  void *p = calloc(n ? n : 1, sz);
  if (!p) orc_die("out of memory", (long)n, (long)sz);
  return p;
}

/* ------------------------------------------------------------------ hasher */

typedef struct {
  int k, w, seed, shift;          /* shift = 64 - 2k                      seqhash.c:32 */
  uint64_t mask;                  /* 2k low bits                          seqhash.c:27 */
  uint64_t factor, factor2;       /* odd multipliers from libc random()   seqhash.c:30-33 */
  uint64_t rcTop[4];              /* (3-b) << 2(k-1)                      seqhash.c:35 */
} OrcHasher;

static void hasher_init(OrcHasher *h, int k, int w, int seed)
{
  if (k < 1 || k >= 32) orc_die("seqhash k must be between 1 and 32", k, 0);   /* seqhash.c:24 */
  if (w < 1) orc_die("seqhash w must be positive", w, 0);                      /* seqhash.c:25 */
  h->k = k; h->w = w; h->seed = seed;
  h->mask = (((uint64_t)1) << (2 * k)) - 1;
  h->shift = 64 - 2 * k;
  srandom((unsigned)seed);
  /* seqhash.c:31 evaluates (random() << 32) | random() | 1; gcc takes the left
     call first - sequenced explicitly here */
  uint64_t hi = (uint64_t)random(), lo = (uint64_t)random();
  h->factor = (hi << 32) | lo | 1;
  hi = (uint64_t)random(); lo = (uint64_t)random();
  h->factor2 = (hi << 32) | lo | 1;
  for (int b = 0; b < 4; ++b) h->rcTop[b] = ((uint64_t)(3 - b)) << (2 * (k - 1));
}

static inline uint64_t hash_of(const OrcHasher *h, uint64_t kmer)     /* seqhash.h:58 */
{ return (kmer * h->factor) >> h->shift; }

void orc_hasher(int k, int w, int seed, uint64_t out[4])
{
  OrcHasher h; hasher_init(&h, k, w, seed);
  out[0] = h.mask; out[1] = (uint64_t)h.shift; out[2] = h.factor; out[3] = h.factor2;
}

/* ------------------------------------------------------------ the scanner --
 * What modRCiterator + repeated modRCnext deliver (seqhash.c:154-196): every
 * window start p in [0, len-k], in order, whose canonical hash
 *     min(hash(fwd), hash(rc)), ties going to the reverse strand (seqhash.c:66-67)
 * is divisible by w; reported as (kmer of the winning strand, p, isForward).
 * fwd rolls in at the bottom, rc at the top (seqhash.c:72-74).
 */
typedef void (*orc_emit_fn)(void *ctx, uint64_t kmer, int pos, int isF);

static int64_t scan_sequence(const OrcHasher *h, const char *s, int len, orc_emit_fn emit, void *ctx)
{
  const int k = h->k;
  if (len < k) return 0;                                   /* seqhash.c:162 */
  uint64_t fwd = 0, rc = 0;
  int64_t n = 0;
  for (int i = 0; i < len; ++i)
    { unsigned b = (unsigned char)s[i];
      fwd = ((fwd << 2) & h->mask) | b;
      rc = (rc >> 2) | h->rcTop[b];
      if (i < k - 1) continue;                             /* still priming, seqhash.c:164-168 */
      uint64_t hf = hash_of(h, fwd), hr = hash_of(h, rc);
      int isF = hf < hr;
      uint64_t hv = isF ? hf : hr;
      if (hv % (uint64_t)h->w == 0)                        /* seqhash.c:171,190 */
        { emit(ctx, isF ? fwd : rc, i - k + 1, isF); ++n; }
    }
  return n;
}

typedef struct { uint64_t *kmer; int32_t *pos; uint8_t *isF; int64_t cap, n; } ScanOut;

static void emit_store(void *ctx, uint64_t kmer, int pos, int isF)
{
  ScanOut *o = (ScanOut *)ctx;
  if (o->n < o->cap)
    { if (o->kmer) o->kmer[o->n] = kmer;
      if (o->pos) o->pos[o->n] = pos;
      if (o->isF) o->isF[o->n] = (uint8_t)isF;
    }
  ++o->n;
}

int64_t orc_mod_scan(int k, int w, int seed, const char *codes, int len,
                     uint64_t *kmer, int32_t *pos, uint8_t *isF, int64_t cap)
{
  OrcHasher h; hasher_init(&h, k, w, seed);
  ScanOut o = { kmer, pos, isF, cap, 0 };
  return scan_sequence(&h, codes, len, emit_store, &o);
}

/* ------------------------------------------------------------------ modset */

struct HxModset {
  OrcHasher hasher;
  int bits;
  uint64_t tableSize, tableMask;
  uint32_t size;                 /* capacity of value/depth/info         modset.c:24-26 */
  uint32_t max;                  /* entries; valid indices are 1..max    modset.c:57 */
  uint32_t *index;               /* 2^bits slots, 0 = empty */
  uint64_t *value;
  uint16_t *depth;
  uint8_t *info;
};

static HxModset *modset_alloc(const OrcHasher *h, int bits, uint32_t size)
{
  if (bits < 20 || bits > 34) orc_die("table bits must be between 20 and 34", bits, 0);  /* modset.c:17 */
  HxModset *ms = (HxModset *)xcalloc(1, sizeof(HxModset));
  ms->hasher = *h;
  ms->bits = bits;
  ms->tableSize = ((uint64_t)1) << bits;
  ms->tableMask = ms->tableSize - 1;
  if (size >= (ms->tableSize >> 2)) orc_die("Modset size too big for bits", size, bits); /* modset.c:24 */
  ms->size = size ? size : (uint32_t)((ms->tableSize >> 2) - 1);                        /* modset.c:25-26 */
  ms->index = (uint32_t *)xcalloc(ms->tableSize, sizeof(uint32_t));
  ms->value = (uint64_t *)xcalloc(ms->size, sizeof(uint64_t));
  ms->depth = (uint16_t *)xcalloc(ms->size, sizeof(uint16_t));
  ms->info = (uint8_t *)xcalloc(ms->size, sizeof(uint8_t));
  return ms;
}

HxModset *orc_modset_new(int bits, int k, int w, int seed)
{
  OrcHasher h; hasher_init(&h, k, w, seed);
  return modset_alloc(&h, bits, 0);
}

void orc_modset_prefault(HxModset *ms) { memset(ms->index, 0, (size_t)(ms->tableMask + 1) * sizeof(uint32_t)); }

void orc_modset_free(HxModset *ms)
{
  if (!ms) return;
  free(ms->index); free(ms->value); free(ms->depth); free(ms->info); free(ms);
}

/* modsetIndexFind (modset.c:45-62): home slot = low `bits` bits of the k-mer's
   seqhash; on collision step by an odd stride taken from the next `bits` bits */
static uint32_t modset_find(HxModset *ms, uint64_t kmer, int add)
{
  uint64_t hv = hash_of(&ms->hasher, kmer);
  uint64_t slot = hv & ms->tableMask;
  uint64_t stride = ((hv >> ms->bits) & ms->tableMask) | 1;
  uint32_t ix;
  while ((ix = ms->index[slot]) != 0 && ms->value[ix] != kmer)
    slot = (slot + stride) & ms->tableMask;
  if (!ix && add)
    { ix = ++ms->max;
      ms->index[slot] = ix;
      if (ms->max >= ms->size) orc_die("hashTableSize is too small", ms->size, ms->max);  /* modset.c:58 */
      ms->value[ix] = kmer;
    }
  return ix;
}

static void emit_count(void *ctx, uint64_t kmer, int pos, int isF)
{
  (void)pos; (void)isF;
  HxModset *ms = (HxModset *)ctx;
  uint32_t ix = modset_find(ms, kmer, 1);
  if (ms->depth[ix] != 0xFFFF) ++ms->depth[ix];            /* saturating ++, modutils.c:26 */
}

uint64_t orc_modset_add(HxModset *ms, const char *codes, const uint64_t *offs, int64_t nseq)
{
  uint64_t tot = 0;
  for (int64_t r = 0; r < nseq; ++r)                       /* addSequence, modutils.c:19-31 */
    tot += (uint64_t)scan_sequence(&ms->hasher, codes + offs[r], (int)(offs[r + 1] - offs[r]), emit_count, ms);
  return tot;
}

uint32_t orc_modset_max(HxModset *ms) { return ms->max; }

void orc_modset_export(HxModset *ms, uint64_t *value, uint16_t *depth, uint8_t *info)
{
  for (uint32_t i = 1; i <= ms->max; ++i)
    { if (value) value[i - 1] = ms->value[i];
      if (depth) depth[i - 1] = ms->depth[i];
      if (info) info[i - 1] = ms->info[i];
    }
}

uint32_t orc_modset_find(HxModset *ms, uint64_t kmer) { return modset_find(ms, kmer, 0); }

/* copy classes live in the low two info bits; note the asymmetry of the
   reference setters: 0/1/2 clear-then-set, M only ORs (modset.h:53-56) */
static inline void set_copy(HxModset *ms, uint32_t i, int c)
{
  if (c == 3) ms->info[i] |= 3;
  else ms->info[i] = (uint8_t)((ms->info[i] & 0xfc) | c);
}

void orc_modset_setcopy(HxModset *ms, int c1, int c2, int cM)        /* modutils.c:205-214 */
{
  for (uint32_t i = 1; i <= ms->max; ++i)
    { int d = ms->depth[i];
      set_copy(ms, i, d < c1 ? 0 : d < c2 ? 1 : d < cM ? 2 : 3);
    }
}

void orc_modset_setcopyM(HxModset *ms, int cM)                       /* modutils.c:215-219 */
{
  for (uint32_t i = 1; i <= ms->max; ++i)
    if ((int)ms->depth[i] >= cM) set_copy(ms, i, 3);
}

void orc_modset_hist(HxModset *ms, uint32_t *bins)                   /* modutils.c:53-63 */
{
  memset(bins, 0, 65536 * sizeof(uint32_t));
  for (uint32_t i = 1; i <= ms->max; ++i) ++bins[ms->depth[i]];
}

/* modsetSummary (modset.c:130-153) including its 32-bit products: the
   reference multiplies U32 depth by U32 bin count before widening */
int orc_modset_summary(HxModset *ms, char *buf, int n)
{
  int o = 0;
  o += snprintf(buf + o, (size_t)(n - o), "SH k %d  w/m %d  s %d\n", ms->hasher.k, ms->hasher.w, ms->hasher.seed);
  o += snprintf(buf + o, (size_t)(n - o), "MS table bits %d size %llu number of entries %u",
                ms->bits, (unsigned long long)ms->tableSize, ms->max);
  if (!ms->max) { o += snprintf(buf + o, (size_t)(n - o), "\n"); return o; }
  uint32_t *h = (uint32_t *)xcalloc(65536, sizeof(uint32_t));
  uint32_t copy[4] = { 0, 0, 0, 0 }, top = 0;
  for (uint32_t i = 1; i <= ms->max; ++i)
    { ++h[ms->depth[i]];
      if ((uint32_t)ms->depth[i] + 1 > top) top = (uint32_t)ms->depth[i] + 1;   /* arrayMax(h) */
      ++copy[ms->info[i] & 3];
    }
  uint64_t sum = 0, tot = 0;
  for (uint32_t i = 0; i < top; ++i) { sum += h[i]; tot += (uint32_t)(i * h[i]); }
  int64_t half = (int64_t)(tot / 2);
  uint32_t n50;
  for (n50 = 0; n50 < top; ++n50) { half -= (uint32_t)(n50 * h[n50]); if (half < 0) break; }
  o += snprintf(buf + o, (size_t)(n - o), " total count %llu\nMS average depth %.1f N50 depth %u",
                (unsigned long long)tot, tot / (double)sum, n50);
  if (copy[0] < ms->max)
    o += snprintf(buf + o, (size_t)(n - o), " copy0 %u copy1 %u copy2 %u copyM %u", copy[0], copy[1], copy[2], copy[3]);
  o += snprintf(buf + o, (size_t)(n - o), "\n");
  free(h);
  return o;
}

/* modsetDepthPrune (modset.c:64-77): re-insert survivors in index order */
void orc_modset_prune(HxModset *ms, int min, int max)
{
  uint32_t n = ms->max;
  ms->max = 0;
  memset(ms->index, 0, ms->tableSize * sizeof(uint32_t));
  for (uint32_t i = 1; i <= n; ++i)
    if ((int)ms->depth[i] >= min && (!max || (int)ms->depth[i] < max))
      { modset_find(ms, ms->value[i], 1);
        ms->info[ms->max] = ms->info[i];
        ms->depth[ms->max] = ms->depth[i];
      }
}

/* modsetMerge (modset.c:106-128): depths add and clamp at 65535, copy numbers
   add and clamp at 3; other info bits of the target are dropped (&= 3) */
int orc_modset_merge(HxModset *a, HxModset *b)
{
  if (a->hasher.w != b->hasher.w || a->hasher.k != b->hasher.k || a->hasher.factor != b->hasher.factor) return 0;
  uint64_t want = (uint64_t)a->max + b->max + 1;
  if (want >= (a->tableSize >> 2)) want = (a->tableSize >> 2) - 1;
  a->value = (uint64_t *)realloc(a->value, want * sizeof(uint64_t));
  a->depth = (uint16_t *)realloc(a->depth, want * sizeof(uint16_t));
  a->info = (uint8_t *)realloc(a->info, want * sizeof(uint8_t));
  if (want > a->size)
    { memset(a->depth + a->size, 0, (want - a->size) * sizeof(uint16_t));   /* see note below */
      memset(a->info + a->size, 0, (want - a->size) * sizeof(uint8_t));
    }
  /* note: the reference's resize() leaves the new tail uninitialised (utils.h:54);
     on a fresh glibc heap that tail reads as zero, which is what we make explicit */
  a->size = (uint32_t)want;
  for (uint32_t i = 1; i <= b->max; ++i)
    { uint32_t ix = modset_find(a, b->value[i], 1);
      uint32_t d = (uint32_t)a->depth[ix] + b->depth[i];
      a->depth[ix] = (uint16_t)(d > 0xFFFF ? 0xFFFF : d);
      int c = (a->info[ix] & 3) + (b->info[i] & 3);
      if (c > 3) c = 3;
      a->info[ix] &= 0x3; a->info[ix] |= (uint8_t)c;
    }
  return 1;
}

/* ----------------------------------------------------------------- readset */
typedef struct { HxModset *ms; uint32_t *hit; uint16_t *dx; int64_t cap, n; int32_t *miss; int lastPos; } ReadsetCtx;

static void emit_readset(void *ctx, uint64_t kmer, int pos, int isF)   /* modasm.c:168-178 */
{
  ReadsetCtx *c = (ReadsetCtx *)ctx;
  uint32_t ix = modset_find(c->ms, kmer, 0);
  if (!ix) { ++*c->miss; return; }
  if (c->n < c->cap)
    { c->hit[c->n] = isF ? (ix | 0x80000000u) : ix;
      c->dx[c->n] = (uint16_t)(pos - c->lastPos);
    }
  c->lastPos = pos;
  ++c->n;
  if (c->ms->depth[ix] != 0xFFFF) ++c->ms->depth[ix];
}

int64_t orc_readset(HxModset *ms, const char *codes, const uint64_t *offs, int64_t nseq,
                    uint64_t *hitOff, uint32_t *hit, uint16_t *dx, int32_t *nMiss, int64_t cap)
{
  memset(ms->depth, 0, ((size_t)ms->max + 1) * sizeof(uint16_t));    /* modasm.c:158 */
  ReadsetCtx c = { ms, hit, dx, cap, 0, 0, 0 };
  for (int64_t r = 0; r < nseq; ++r)
    { hitOff[r] = (uint64_t)c.n;
      nMiss[r] = 0; c.miss = &nMiss[r]; c.lastPos = 0;
      scan_sequence(&ms->hasher, codes + offs[r], (int)(offs[r + 1] - offs[r]), emit_readset, &c);
    }
  hitOff[nseq] = (uint64_t)c.n;
  return c.n;
}

/* -------------------------------------------------------- modmap reference */

struct HxRef {
  HxModset *ms;
  uint32_t cap, max;            /* hit list capacity (2^26 in the tool, modmap.c:363) / length */
  uint32_t *index, *offset, *id;
  uint32_t *depth;              /* occurrences of each modset index in the reference */
  uint32_t *rev, *loc;
};

typedef struct { HxRef *r; uint32_t id; } RefCtx;

static void emit_ref(void *ctx, uint64_t kmer, int pos, int isF)     /* modmap.c:108-118 */
{
  (void)isF;
  RefCtx *c = (RefCtx *)ctx; HxRef *r = c->r;
  uint32_t ix = modset_find(r->ms, kmer, 1);
  if (!ix) return;
  if (r->max + 1 >= r->cap) orc_die("reference size overflow", r->max, r->cap);
  r->index[r->max] = ix;
  ++r->depth[ix];
  r->offset[r->max] = (uint32_t)pos;
  r->id[r->max] = c->id;
  ++r->max;
}

HxRef *orc_ref_build(int bits, int k, int w, int seed, const char *codes,
                     const uint64_t *offs, int64_t nseq, uint32_t counts[4])
{
  HxRef *r = (HxRef *)xcalloc(1, sizeof(HxRef));
  r->ms = orc_modset_new(bits, k, w, seed);
  r->cap = 1u << 26;
  r->index = (uint32_t *)xcalloc(r->cap, sizeof(uint32_t));
  r->offset = (uint32_t *)xcalloc(r->cap, sizeof(uint32_t));
  r->id = (uint32_t *)xcalloc(r->cap, sizeof(uint32_t));
  r->depth = (uint32_t *)xcalloc(r->ms->size, sizeof(uint32_t));
  for (int64_t s = 0; s < nseq; ++s)
    { RefCtx c = { r, (uint32_t)s };
      scan_sequence(&r->ms->hasher, codes + offs[s], (int)(offs[s + 1] - offs[s]), emit_ref, &c);
    }
  /* classify by exact multiplicity in the reference, modmap.c:125-129 */
  uint32_t n1 = 0, n2 = 0, nM = 0;
  HxModset *ms = r->ms;
  for (uint32_t i = 1; i <= ms->max; ++i)
    if (r->depth[i] == 1) { set_copy(ms, i, 1); ++n1; }
    else if (r->depth[i] == 2) { set_copy(ms, i, 2); ++n2; }
    else { set_copy(ms, i, 3); ++nM; }
  if (counts) { counts[0] = r->max; counts[1] = n1; counts[2] = n2; counts[3] = nM; }
  /* referencePack, modmap.c:74-91: loc = exclusive prefix sum of depth over
     index, rev = occurrence ordinals grouped by index, stable */
  r->loc = (uint32_t *)xcalloc((size_t)ms->max + 1, sizeof(uint32_t));
  r->rev = (uint32_t *)xcalloc(r->max ? r->max : 1, sizeof(uint32_t));
  for (uint32_t i = 1; i <= ms->max; ++i) r->loc[i] = r->loc[i - 1] + r->depth[i - 1];
  uint32_t *fill = (uint32_t *)xcalloc((size_t)ms->max + 1, sizeof(uint32_t));
  for (uint32_t n = 0; n < r->max; ++n) { uint32_t ix = r->index[n]; r->rev[r->loc[ix] + fill[ix]++] = n; }
  free(fill);
  return r;
}

void orc_ref_free(HxRef *r)
{
  if (!r) return;
  orc_modset_free(r->ms);
  free(r->index); free(r->offset); free(r->id); free(r->depth); free(r->rev); free(r->loc); free(r);
}

HxModset *orc_ref_modset(HxRef *r) { return r->ms; }
uint32_t orc_ref_max(HxRef *r) { return r->max; }

void orc_ref_export(HxRef *r, uint32_t *index, uint32_t *offset, uint32_t *id,
                    uint32_t *depth, uint32_t *rev, uint32_t *loc)
{
  size_t n = r->max, m = (size_t)r->ms->max + 1;
  if (index) memcpy(index, r->index, n * 4);
  if (offset) memcpy(offset, r->offset, n * 4);
  if (id) memcpy(id, r->id, n * 4);
  if (depth) memcpy(depth, r->depth, m * 4);
  if (rev) memcpy(rev, r->rev, n * 4);
  if (loc) memcpy(loc, r->loc, m * 4);
}

typedef struct {
  HxRef *r; int64_t cap, n;
  uint32_t *seedIndex, *seedPos, *hitId, *hitOffset;
  int32_t *ctr;
} QueryCtx;

static void emit_query(void *ctx, uint64_t kmer, int pos, int isF)   /* modmap.c:201-207, 216-231 */
{
  (void)isF;
  QueryCtx *q = (QueryCtx *)ctx; HxRef *r = q->r; HxModset *ms = r->ms;
  uint32_t ix = modset_find(ms, kmer, 0);
  if (!ix) ++q->ctr[0];                                   /* miss */
  else if (ms->info[ix] & 3) ++q->ctr[ms->info[ix] & 3];  /* copy0 hits are counted nowhere printed */
  if (q->n < q->cap)
    { int64_t n = q->n;
      q->seedIndex[n] = ix; q->seedPos[n] = (uint32_t)pos;
      q->hitId[2 * n] = q->hitId[2 * n + 1] = 0xFFFFFFFFu;
      q->hitOffset[2 * n] = q->hitOffset[2 * n + 1] = 0xFFFFFFFFu;
      if (ix && (ms->info[ix] & 3) != 3)
        { uint32_t l = r->rev[r->loc[ix]];
          q->hitId[2 * n] = r->id[l]; q->hitOffset[2 * n] = r->offset[l];
          if ((ms->info[ix] & 3) != 1)
            { uint32_t l2 = r->rev[r->loc[ix] + 1];
              q->hitId[2 * n + 1] = r->id[l2]; q->hitOffset[2 * n + 1] = r->offset[l2];
            }
        }
    }
  ++q->n;
}

int64_t orc_ref_query(HxRef *r, const char *codes, const uint64_t *offs, int64_t nseq,
                      uint64_t *seedOff, uint32_t *seedIndex, uint32_t *seedPos,
                      uint32_t *hitId, uint32_t *hitOffset, int32_t *counters, int64_t cap)
{
  QueryCtx q = { r, cap, 0, seedIndex, seedPos, hitId, hitOffset, 0 };
  for (int64_t s = 0; s < nseq; ++s)
    { q.ctr = counters + 4 * s;
      q.ctr[0] = q.ctr[1] = q.ctr[2] = q.ctr[3] = 0;     /* {miss, copy1, copy2, multi} */
      seedOff[s] = (uint64_t)q.n;
      scan_sequence(&r->ms->hasher, codes + offs[s], (int)(offs[s + 1] - offs[s]), emit_query, &q);
    }
  seedOff[nseq] = (uint64_t)q.n;
  return q.n;
}
