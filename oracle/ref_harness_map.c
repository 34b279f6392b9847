/* ref_harness_map.c - the modmap half of harness_api.h on the UNMODIFIED
 * reference.  TEST INFRASTRUCTURE ONLY (see harness_api.h).
 *
 * modmap.c is pulled in as a translation unit from /root/reference (-I, not
 * copied) with main() and its globals renamed; the harness calls the real
 * referenceCreate / referencePack / referenceFastaRead and the real
 * seqhash.o / modset.o.  The in-memory builder repeats the per-hit statements
 * of referenceFastaRead (modmap.c:106-129) over reference primitives because
 * that function only reads files; ref_ref_build_fasta() runs the real one.
 */
#define _GNU_SOURCE
#define main ref_modmap_main
#define usage ref_modmap_usage
#define outFile ref_modmap_outFile
#define isVerbose ref_modmap_isVerbose
#define numThreads ref_modmap_numThreads
#define params ref_modmap_params
#include "modmap.c"
#undef main
#undef usage

#define HX(name) ref_##name
#include "harness_api.h"

static void classify(Reference *ref, uint32_t counts[4])
{
  U32 n1 = 0, n2 = 0, nM = 0;
  for (U32 i = 1; i <= ref->ms->max; ++i)
    if (ref->depth[i] == 1) { msSetCopy1(ref->ms, i); ++n1; }
    else if (ref->depth[i] == 2) { msSetCopy2(ref->ms, i); ++n2; }
    else { msSetCopyM(ref->ms, i); ++nM; }
  if (counts) { counts[0] = ref->max; counts[1] = n1; counts[2] = n2; counts[3] = nM; }
}

HxRef *ref_ref_build(int bits, int k, int w, int seed, const char *codes,
                     const uint64_t *offs, int64_t nseq, uint32_t counts[4])
{
  Modset *ms = modsetCreate(seqhashCreate(k, w, seed), bits, 0);
  Reference *ref = referenceCreate(ms, 1 << 26);
  for (int64_t s = 0; s < nseq; ++s)
    { SeqhashRCiterator *mi = modRCiterator(ms->hasher, (char *)codes + offs[s], (int)(offs[s + 1] - offs[s]));
      U64 kmer; int pos;
      while (modRCnext(mi, &kmer, &pos, 0))
        { U32 index = modsetIndexFind(ms, kmer, true);
          if (!index) continue;
          if (ref->max + 1 >= ref->size) die("reference size overflow");
          ref->index[ref->max] = index;
          ++ref->depth[index];
          ref->offset[ref->max] = pos;
          ref->id[ref->max] = (U32)s;
          ++ref->max;
        }
      seqhashRCiteratorDestroy(mi);
    }
  classify(ref, counts);
  modsetPack(ms);
  referencePack(ref);
  return (HxRef *)ref;
}

/* the real thing, from a FASTA file through seqio */
HxRef *ref_ref_build_fasta(int bits, int k, int w, int seed, const char *path, uint32_t counts[4])
{
  if (!ref_modmap_outFile) ref_modmap_outFile = fopen("/dev/null", "w");
  Modset *ms = modsetCreate(seqhashCreate(k, w, seed), bits, 0);
  Reference *ref = referenceCreate(ms, 1 << 26);
  referenceFastaRead(ref, (char *)path, true);
  if (counts)
    { counts[0] = ref->max; counts[1] = counts[2] = counts[3] = 0;
      for (U32 i = 1; i <= ms->max; ++i) ++counts[msCopy(ms, i) ? msCopy(ms, i) : 0];
      counts[0] = ref->max;
    }
  return (HxRef *)ref;
}

void ref_ref_free(HxRef *h)
{
  Reference *ref = (Reference *)h;
  if (!ref) return;
  Modset *ms = ref->ms;
  referenceDestroy(ref);
  free(ms->depth); modsetDestroy(ms);
}

HxModset *ref_ref_modset(HxRef *h) { return (HxModset *)((Reference *)h)->ms; }
uint32_t ref_ref_max(HxRef *h) { return ((Reference *)h)->max; }

void ref_ref_export(HxRef *h, uint32_t *index, uint32_t *offset, uint32_t *id,
                    uint32_t *depth, uint32_t *rev, uint32_t *loc)
{
  Reference *r = (Reference *)h;
  size_t n = r->max, m = (size_t)r->ms->max + 1;
  if (index) memcpy(index, r->index, n * 4);
  if (offset) memcpy(offset, r->offset, n * 4);
  if (id) memcpy(id, r->id, n * 4);
  if (depth) memcpy(depth, r->depth, m * 4);
  if (rev) memcpy(rev, r->rev, n * 4);
  if (loc) memcpy(loc, r->loc, m * 4);
}

int64_t ref_ref_query(HxRef *h, const char *codes, const uint64_t *offs, int64_t nseq,
                      uint64_t *seedOff, uint32_t *seedIndex, uint32_t *seedPos,
                      uint32_t *hitId, uint32_t *hitOffset, int32_t *counters, int64_t cap)
{
  Reference *ref = (Reference *)h;
  int64_t n = 0;
  for (int64_t s = 0; s < nseq; ++s)
    { int32_t *ctr = counters + 4 * s;
      ctr[0] = ctr[1] = ctr[2] = ctr[3] = 0;
      seedOff[s] = (uint64_t)n;
      SeqhashRCiterator *mi = modRCiterator(ref->ms->hasher, (char *)codes + offs[s], (int)(offs[s + 1] - offs[s]));
      U64 kmer; int pos;
      while (modRCnext(mi, &kmer, &pos, 0))
        { U32 index = modsetIndexFind(ref->ms, kmer, false);
          if (!index) ++ctr[0];
          else if (msCopy(ref->ms, index)) ++ctr[msCopy(ref->ms, index)];
          if (n < cap)
            { seedIndex[n] = index; seedPos[n] = (uint32_t)pos;
              hitId[2 * n] = hitId[2 * n + 1] = hitOffset[2 * n] = hitOffset[2 * n + 1] = 0xFFFFFFFFu;
              if (index && !msIsCopyM(ref->ms, index))
                { U32 loc = ref->rev[ref->loc[index]];
                  hitId[2 * n] = ref->id[loc]; hitOffset[2 * n] = ref->offset[loc];
                  if (!msIsCopy1(ref->ms, index))
                    { U32 loc2 = ref->rev[ref->loc[index] + 1];
                      hitId[2 * n + 1] = ref->id[loc2]; hitOffset[2 * n + 1] = ref->offset[loc2];
                    }
                }
            }
          ++n;
        }
      seqhashRCiteratorDestroy(mi);
    }
  seedOff[nseq] = (uint64_t)n;
  return n;
}
