/* ref_harness_utils.c - harness_api.h implemented by the UNMODIFIED reference.
 * TEST INFRASTRUCTURE ONLY (see harness_api.h).
 *
 * The reference's modutils.c is pulled in as a translation unit, from where it
 * lies under /root/reference (-I, nothing is copied into this repo), with its
 * main() and file-scope globals renamed, so that the harness can call the real
 * static addSequence() (modutils.c:19-31) and depthHistogram() (modutils.c:53-63)
 * and the real seqhash.o / modset.o they link against.
 */
#define _GNU_SOURCE
#define main ref_modutils_main
#define usage ref_modutils_usage
#define outFile ref_modutils_outFile
#define isVerbose ref_modutils_isVerbose
#define reportDepths ref_modutils_reportDepths
#define depthHistogram ref_modutils_depthHistogram
#include "modutils.c"
#undef main
#undef usage

#define HX(name) ref_##name
#include "harness_api.h"

void ref_hasher(int k, int w, int seed, uint64_t out[4])
{
  Seqhash *sh = seqhashCreate(k, w, seed);
  out[0] = sh->mask; out[1] = (uint64_t)sh->shift1; out[2] = sh->factor1; out[3] = sh->factor2;
  seqhashDestroy(sh);
}

int64_t ref_mod_scan(int k, int w, int seed, const char *codes, int len,
                     uint64_t *kmer, int32_t *pos, uint8_t *isF, int64_t cap)
{
  Seqhash *sh = seqhashCreate(k, w, seed);
  SeqhashRCiterator *it = modRCiterator(sh, (char *)codes, len);
  U64 km; int p; bool f; int64_t n = 0;
  while (modRCnext(it, &km, &p, &f))
    { if (n < cap)
        { if (kmer) kmer[n] = km;
          if (pos) pos[n] = p;
          if (isF) isF[n] = f ? 1 : 0;
        }
      ++n;
    }
  seqhashRCiteratorDestroy(it);
  seqhashDestroy(sh);
  return n;
}

HxModset *ref_modset_new(int bits, int k, int w, int seed)
{ return (HxModset *)modsetCreate(seqhashCreate(k, w, seed), bits, 0); }

/* touch every page of the (calloc'ed, still unmapped) index table: a timed bounded sample then measures hashing and
   probing, not the page faults a full-size job pays once (bench.py --impl reference) */
void ref_modset_prefault(HxModset *h)
{ Modset *ms = (Modset *)h; memset(ms->index, 0, ms->tableSize * sizeof(U32)); }

void ref_modset_free(HxModset *h)
{ Modset *ms = (Modset *)h; if (!ms) return; free(ms->depth); modsetDestroy(ms); }

uint64_t ref_modset_add(HxModset *h, const char *codes, const uint64_t *offs, int64_t nseq)
{
  Modset *ms = (Modset *)h;
  uint64_t tot = 0;
  for (int64_t r = 0; r < nseq; ++r)
    tot += (uint64_t)addSequence(ms, (char *)codes + offs[r], (int)(offs[r + 1] - offs[r]));
  return tot;
}

uint32_t ref_modset_max(HxModset *h) { return ((Modset *)h)->max; }

void ref_modset_export(HxModset *h, uint64_t *value, uint16_t *depth, uint8_t *info)
{
  Modset *ms = (Modset *)h;
  for (U32 i = 1; i <= ms->max; ++i)
    { if (value) value[i - 1] = ms->value[i];
      if (depth) depth[i - 1] = ms->depth[i];
      if (info) info[i - 1] = ms->info[i];
    }
}

uint32_t ref_modset_find(HxModset *h, uint64_t kmer) { return modsetIndexFind((Modset *)h, kmer, 0); }

/* the two classification commands are inline in the reference's main()
   (modutils.c:205-219); driven here through the reference's own accessors */
void ref_modset_setcopy(HxModset *h, int c1, int c2, int cM)
{
  Modset *ms = (Modset *)h;
  for (U32 u = 1; u <= ms->max; ++u)
    if (ms->depth[u] < c1) msSetCopy0(ms, u);
    else if (ms->depth[u] < c2) msSetCopy1(ms, u);
    else if (ms->depth[u] < cM) msSetCopy2(ms, u);
    else msSetCopyM(ms, u);
}

void ref_modset_setcopyM(HxModset *h, int cM)
{
  Modset *ms = (Modset *)h;
  for (U32 u = 1; u <= ms->max; ++u) if (ms->depth[u] >= cM) msSetCopyM(ms, u);
}

void ref_modset_hist(HxModset *h, uint32_t *bins)
{
  char *text = 0; size_t len = 0;
  FILE *f = open_memstream(&text, &len);
  ref_modutils_depthHistogram((Modset *)h, f);
  fclose(f);
  memset(bins, 0, 65536 * sizeof(uint32_t));
  char *p = text;
  unsigned d, c; int used;
  while (p && sscanf(p, "DP\t%u\t%u\n%n", &d, &c, &used) == 2) { bins[d] = c; p += used; }
  free(text);
}

int ref_modset_summary(HxModset *h, char *buf, int n)
{
  char *text = 0; size_t len = 0;
  FILE *f = open_memstream(&text, &len);
  modsetSummary((Modset *)h, f);
  fclose(f);
  int o = (int)len < n - 1 ? (int)len : n - 1;
  memcpy(buf, text, (size_t)o); buf[o] = 0;
  free(text);
  return o;
}

void ref_modset_prune(HxModset *h, int min, int max) { modsetDepthPrune((Modset *)h, min, max); }

int ref_modset_merge(HxModset *a, HxModset *b) { return modsetMerge((Modset *)a, (Modset *)b) ? 1 : 0; }

/* the per-read statements of readsetFileRead (modasm.c:151-191) over the reference's own primitives;
   modasm.c itself cannot be pulled in next to modutils.c (both define the same file-scope names) */
int64_t ref_readset(HxModset *h, const char *codes, const uint64_t *offs, int64_t nseq,
                    uint64_t *hitOff, uint32_t *hit, uint16_t *dx, int32_t *nMiss, int64_t cap)
{
  Modset *ms = (Modset *)h;
  int64_t n = 0;
  memset (ms->depth, 0, (ms->max+1)*sizeof(U16)) ;
  for (int64_t r = 0; r < nseq; ++r)
    { hitOff[r] = (uint64_t)n; nMiss[r] = 0;
      SeqhashRCiterator *mi = modRCiterator(ms->hasher, (char *)codes + offs[r], (int)(offs[r + 1] - offs[r]));
      U64 kmer; int lastPos = 0, pos; bool isForward;
      while (modRCnext(mi, &kmer, &pos, &isForward))
        { U32 index = modsetIndexFind(ms, kmer, false);
          if (index)
            { if (n < cap) { hit[n] = isForward ? (index | 0x80000000u) : index; dx[n] = pos - lastPos; }
              lastPos = pos; ++n;
              U16 *di = &ms->depth[index]; ++*di; if (!*di) *di = U16MAX;
            }
          else ++nMiss[r];
        }
      seqhashRCiteratorDestroy(mi);
    }
  hitOff[nseq] = (uint64_t)n;
  return n;
}
