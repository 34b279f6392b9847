/* modmap_hot.c - the two functions of the reference's modmap.c that change for the GPU path.
 *
 * Not compiled on its own: csrc/shim/Makefile takes the reference's modmap.c where it lies, renames
 * referenceFastaRead (modmap.c:93-134) and queryProcess (modmap.c:188-281) out of the way with sed, appends this file
 * and links the result (modmap_dropin) against libmodshim.so - INTEGRATION.md section 2 applied at build time,
 * nothing of the reference copied into this repo.  Reference, referenceCreate / Pack / Write / Read, main() and the
 * printers are the reference's own code; this file sees its `Reference` struct because it is appended to it.
 *
 * Both loops become: collect the sequences of the file in pinned batches (seqio stays on the host), run the
 * modimizer iterator of the WHOLE batch on the GPU (modgpuScannerScan == modRCiterator / modRCnext per sequence,
 * in order) and look all of them up / insert them with ONE batched modsetIndexFind (modgpuModsetIndexFindBatch);
 * what stays on the host is the reference's bookkeeping per hit and its serial colinear-block pass per read.
 */
#include "modshim.h"

typedef struct { char *bases ; U64 *off ; size_t cap, used, nSeq, capSeq ; } HotBatch ;

static void hotPut (HotBatch *b, const char *s, U64 len)
{
  if (b->used + len > b->cap)
    { size_t want = (b->used + len) + (b->used + len) / 2 + (1 << 20) ;
      char *nb = (char*) modgpuHostAlloc (want) ;
      if (!nb) die ("%s", (char*) modgpuLastError ()) ;
      if (b->used) memcpy (nb, b->bases, b->used) ;
      if (b->bases) modgpuHostFree (b->bases) ;
      b->bases = nb ; b->cap = want ;
    }
  if (b->nSeq == b->capSeq)
    { b->capSeq = b->capSeq ? 2 * b->capSeq : 1024 ;
      if (!(b->off = (U64*) realloc (b->off, (b->capSeq + 1) * sizeof (U64)))) die ("out of memory") ;
      b->off[0] = 0 ;
    }
  memcpy (b->bases + b->used, s, len) ;
  b->used += len ; b->off[++b->nSeq] = b->used ;
}

/* every modimizer of every sequence of the batch, in (sequence, position) order, and its modset index */
typedef struct { U64 n ; uint64_t *kmer, *seqOff ; uint32_t *pos, *index ; } HotScan ;

static void hotScan (Modset *ms, HotBatch *b, int isAdd, HotScan *h)
{
  ModgpuScanner *sc = modshimScanner (ms->hasher) ;
  U64 w = (U64) ms->hasher->w ;
  U64 cap = b->used / w + b->used / (4 * w) + 4096 ; if (cap > b->used) cap = b->used + 1 ;
  h->seqOff = (uint64_t*) malloc ((b->nSeq + 1) * 8) ;
  for (int attempt = 0 ; ; ++attempt)
    { h->kmer = (uint64_t*) malloc (cap * 8) ; h->pos = (uint32_t*) malloc (cap * 4) ;
      if (!h->seqOff || !h->kmer || !h->pos) die ("out of memory") ;
      h->n = modgpuScannerScan (sc, b->bases, (uint64_t*) b->off, b->nSeq, 0, h->kmer, h->pos, h->seqOff, cap) ;
      if (h->n == UINT64_MAX) die ("%s", (char*) modgpuLastError ()) ;
      if (h->n <= cap || attempt) break ;
      free (h->kmer) ; free (h->pos) ; cap = h->n ;                    /* denser than expected: once more */
    }
  for (U64 i = 0 ; i < h->n ; ++i) h->kmer[i] &= 0x3FFFFFFFFFFFFFFFull ;   /* bit 63 = isForward, not wanted here */
  h->index = (uint32_t*) malloc ((h->n + 1) * 4) ;
  if (!h->index) die ("out of memory") ;
  if (modgpuModsetIndexFindBatch (modshimTwin (ms), h->kmer, h->n, isAdd, h->index)) die ("%s", (char*) modgpuLastError ()) ;
}

static void hotScanFree (HotScan *h) { free (h->kmer) ; free (h->pos) ; free (h->seqOff) ; free (h->index) ; }

/* ---------------------------------------------------------------- -f ---- */
static void hotRefFlush (Reference *ref, HotBatch *b, int *idOf, bool isAdd)
{
  if (!b->nSeq) return ;
  HotScan h ;
  hotScan (ref->ms, b, isAdd, &h) ;
  for (size_t r = 0 ; r < b->nSeq ; ++r)
    for (U64 i = h.seqOff[r] ; i < h.seqOff[r+1] ; ++i)
      { U32 index = h.index[i] ;
	if (!index) continue ;
	if (ref->max+1 >= ref->size) die ("reference size overflow") ;      /* modmap.c:111 */
	ref->index[ref->max] = index ;
	++ref->depth[index] ;
	ref->offset[ref->max] = h.pos[i] ;
	ref->id[ref->max] = idOf[r] ;
	++ref->max ;
      }
  hotScanFree (&h) ;
  b->used = 0 ; b->nSeq = 0 ;
}

void referenceFastaRead (Reference *ref, char *filename, bool isAdd)
{
  U64 totLen = 0 ;
  HotBatch b ; memset (&b, 0, sizeof (b)) ;
  int *idOf = 0 ; size_t idCap = 0 ;

  dna2indexConv['N'] = dna2indexConv['n'] = 0 ;                             /* modmap.c:97 */
  SeqIO *si = seqIOopenRead (filename, dna2indexConv, false) ;
  if (!si) die ("failed to read reference sequence file %s", filename) ;
  while (seqIOread (si))
    { int id ;
      if (!dictAdd (ref->dict, sqioId(si), &id)) die ("duplicate ref sequence name %s", sqioId(si)) ;
      array (ref->len, id, int) = si->seqLen ;
      totLen += si->seqLen ;
      if (b.used + si->seqLen >= ((U64)1 << 31)) hotRefFlush (ref, &b, idOf, isAdd) ;   /* a scan covers < 2^32 bases */
      if (b.nSeq == idCap) { idCap = idCap ? 2 * idCap : 1024 ; idOf = (int*) realloc (idOf, idCap * sizeof (int)) ; }
      idOf[b.nSeq] = id ;
      hotPut (&b, sqioSeq(si), si->seqLen) ;
    }
  seqIOclose (si) ;
  hotRefFlush (ref, &b, idOf, isAdd) ;
  if (b.bases) modgpuHostFree (b.bases) ;
  free (b.off) ; free (idOf) ;
  if (isAdd) modshimSync (ref->ms) ;                                        /* value[] and max of the new entries */

  /* from here on: the reference's own tail (modmap.c:123-133), on the host arrays */
  fprintf (outFile, "  %d hashes from %d reference sequences, total length %lld\n",
	   ref->max, dictMax(ref->dict), totLen) ;
  int i ; U32 *d = &ref->depth[1] ; U32 n1 = 0, n2 = 0, nM = 0 ;
  for (i = 1 ; i <= ref->ms->max ; ++i, ++d)
    if (*d == 1) { msSetCopy1 (ref->ms, i) ; ++n1 ; }
    else if (*d == 2) { msSetCopy2 (ref->ms, i) ; ++n2 ; }
    else { msSetCopyM (ref->ms, i) ; ++nM ; }
  fprintf (outFile, "  %d copy 1, %d copy 2, %d multiple\n", n1, n2, nM) ;
  if (isAdd) modsetPack (ref->ms) ;
  referencePack (ref) ;
}

/* ---------------------------------------------------------------- -q ---- */
/* The colinear-block pass over the seeds of one read (modmap.c:213-276) as a scanner with explicit state: a block is
 * a run of unique / two-copy seeds whose occurrences lie on one reference sequence, move one way through the
 * occurrence list and keep pace with the seed ordinals to within 50.  Occurrence ordinal 0 doubles as "no block open"
 * in the reference (modmap.c:233): kept, like its 32-bit unsigned arithmetic and its two report rules (a finished
 * block with more than two unique seeds; the block open at the end of the read with more than two TWO-COPY seeds). */
typedef struct { U32 first, last, iFirst, iLast ; int nUnique, nPair ; } HotBlock ;

static int hotLeaves (Reference *ref, const HotBlock *b, U32 occ)
{
  if (ref->id[occ] != ref->id[b->first]) return 1 ;
  if (b->first == b->last) return 0 ;
  int drift ;
  if (b->first < b->last)
    { if (occ < b->last) return 1 ;
      drift = (int) (b->last - b->first - b->iLast + b->iFirst) ;
    }
  else
    { if (occ > b->last) return 1 ;
      drift = (int) (b->first - b->last - b->iLast + b->iFirst) ;
    }
  return drift > 50 || drift < -50 ;
}

static void hotBlockPrint (Reference *ref, const HotBlock *b, const char *name, const uint32_t *pos, int nCopy1)
{
  U32 span = b->last > b->first ? b->last - b->first : b->first - b->last ;
  fprintf (outFile, "M\t%s\t%d\t%d\t%d\t%s\t%d\t%d\t%d %d\t%.2f\t%.2f\n", name,
	   (int) pos[b->iFirst], (int) pos[b->iLast], (int) (pos[b->iLast] - pos[b->iFirst]),
	   dictName (ref->dict, ref->id[b->first]), ref->offset[b->first], ref->offset[b->last],
	   b->nUnique, b->nPair, (b->nUnique + b->nPair) / (double) span, b->nUnique / (double) nCopy1) ;
}

static void hotQueryFlush (Reference *ref, HotBatch *b, char **name, U64 *lenOf)
{
  if (!b->nSeq) return ;
  HotScan h ;
  Modset *ms = ref->ms ;
  hotScan (ms, b, 0, &h) ;
  for (size_t r = 0 ; r < b->nSeq ; ++r)
    { const uint32_t *index = h.index + h.seqOff[r], *pos = h.pos + h.seqOff[r] ;
      int nSeed = (int) (h.seqOff[r+1] - h.seqOff[r]) ;
      int missed = 0, copy[4] ; copy[1] = copy[2] = copy[3] = 0 ;
      for (int i = 0 ; i < nSeed ; ++i) if (index[i]) ++copy[msCopy(ms,index[i])] ; else ++missed ;      /* modmap.c:206-207 */
      fprintf (outFile, "Q\t%s\t%llu\t%d miss, %d copy1, %d copy2, %d multi, %.2f hit\n",
	       name[r], lenOf[r], missed, copy[1], copy[2], copy[3], (nSeed-missed)/(double)(nSeed)) ;
      HotBlock k ; memset (&k, 0, sizeof (k)) ;
      for (int i = 0 ; i < nSeed ; ++i)
	{ if (!index[i] || msIsCopyM(ms,index[i])) continue ;                      /* modmap.c:217 */
	  U32 at = ref->loc[index[i]], occ = ref->rev[at] ;
	  bool unique = msIsCopy1(ms,index[i]) ;
	  if (isVerbose && unique)                                                  /* to stdout, modmap.c:221-229 */
	    printf ("  %6d\t%s %d\n", pos[i], dictName(ref->dict,ref->id[occ]), ref->offset[occ]) ;
	  else if (isVerbose)
	    { U32 occ2 = ref->rev[at+1] ;
	      printf ("  %6d\t%s %d\t%s %d\n", pos[i], dictName(ref->dict,ref->id[occ]), ref->offset[occ],
		      dictName(ref->dict,ref->id[occ2]), ref->offset[occ2]) ;
	    }
	  int ends = !k.first || hotLeaves (ref, &k, occ) ;
	  if (ends && k.first && !unique) { occ = ref->rev[at+1] ; ends = hotLeaves (ref, &k, occ) ; }
	  if (ends)
	    { if (k.nUnique > 2) hotBlockPrint (ref, &k, name[r], pos, copy[1]) ;
	      k.nUnique = k.nPair = 0 ; k.first = occ ; k.iFirst = i ;
	    }
	  if (unique) ++k.nUnique ; else ++k.nPair ;
	  k.last = occ ; k.iLast = i ;
	}
      if (k.nPair > 2) hotBlockPrint (ref, &k, name[r], pos, copy[1]) ;
      free (name[r]) ;
    }
  hotScanFree (&h) ;
  b->used = 0 ; b->nSeq = 0 ;
}

void queryProcess (Reference *ref, char *filename)
{
  HotBatch b ; memset (&b, 0, sizeof (b)) ;
  const size_t maxReads = 1 << 16 ;
  char **name = (char**) calloc (maxReads, sizeof (char*)) ;
  U64 *lenOf = (U64*) calloc (maxReads, sizeof (U64)) ;

  dna2indexConv['N'] = dna2indexConv['n'] = 0 ;                             /* modmap.c:193 */
  SeqIO *si = seqIOopenRead (filename, dna2indexConv, false) ;
  if (!si) die ("failed to read query sequence file %s", filename) ;
  while (seqIOread (si))
    { if (b.nSeq == maxReads || b.used + si->seqLen > ((U64)1 << 27)) hotQueryFlush (ref, &b, name, lenOf) ;
      name[b.nSeq] = strdup (sqioId(si)) ; lenOf[b.nSeq] = si->seqLen ;
      hotPut (&b, sqioSeq(si), si->seqLen) ;
    }
  hotQueryFlush (ref, &b, name, lenOf) ;
  seqIOclose (si) ;
  if (b.bases) modgpuHostFree (b.bases) ;
  free (b.off) ; free (name) ; free (lenOf) ;
}
