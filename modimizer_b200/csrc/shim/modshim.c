/* modshim.c - the reference's OWN seqhash symbols, served by libmodgpu (B200).
 *
 * Compiled with the C compiler against the reference's headers where they lie
 * (-I$(REF): the struct layouts are the reference's by construction, nothing is
 * copied into this repo), so an unmodified caller that links this file instead of
 * seqhash.o keeps working:
 *
 *   seqhashCreate / seqhashWrite / seqhashRead / seqhashReport   seqhash.c:20-56
 *   modRCiterator / modRCnext                                    seqhash.c:154-196
 *   seqString                                                    seqhash.c:198-206
 *
 * The iterator runs K1 + K2 (ordered, with strand and position) on the one sequence
 * when it is created and modRCnext pops the result list - same values, same order,
 * same "false after the last one" as the serial iterator.  The result arrays live in
 * the iterator's hashBuf / fBuf members, which the header-inline
 * seqhashRCiteratorDestroy (seqhash.h:54-55) frees.  Errors go through the
 * reference's die() (utils.c:19-30), as everywhere in the reference.
 *
 * The minimizer iterator (seqhash.c:83-152) is not on the GPU path (no tool calls it,
 * SURVEY section 2): calling it dies.  There is no CPU fallback anywhere in here.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "seqhash.h"                 /* the reference's header ($(REF)) */
#include "modgpu.h"

static ModgpuScanner *gScanner ;     /* one scanner, re-created when the hasher changes */
static ModgpuHasher gHasher ;

ModgpuScanner *modshimScanner (Seqhash *sh)
{
  ModgpuHasher h ;
  if (modgpuHasherFromSeqhash (&h, sh)) die ("%s", (char*) modgpuLastError ()) ;
  if (gScanner && (h.k != gHasher.k || h.w != gHasher.w || h.factor1 != gHasher.factor1))
    { modgpuScannerDestroy (gScanner) ; gScanner = 0 ; }
  if (!gScanner)
    { if (!(gScanner = modgpuScannerCreate (&h))) die ("%s", (char*) modgpuLastError ()) ;
      gHasher = h ;
    }
  return gScanner ;
}

Seqhash *seqhashCreate (int k, int w, int seed)
{
  ModgpuHasher h ;
  if (k < 1 || k >= 32) die ("seqhash k %d must be between 1 and 32\n", k) ;     /* seqhash.c:24 */
  if (w < 1) die ("seqhash w %d must be positive\n", w) ;                        /* seqhash.c:25 */
  if (modgpuHasherInit (&h, k, w, seed)) die ("%s", (char*) modgpuLastError ()) ; /* srandom/random like seqhash.c:30-33 */
  Seqhash *sh = (Seqhash*) calloc (1, sizeof (Seqhash)) ;
  if (!sh) die ("seqhashCreate: out of memory") ;
  sh->seed = seed ; sh->k = k ; sh->w = w ;
  sh->mask = h.mask ;
  sh->shift1 = h.shift1 ; sh->shift2 = 2*k ;
  sh->factor1 = h.factor1 ; sh->factor2 = h.factor2 ;
  for (int b = 0 ; b < 4 ; ++b) sh->patternRC[b] = ((U64)(3 - b)) << (2*(k-1)) ;
  return sh ;
}

void seqhashWrite (Seqhash *sh, FILE *f)
{
  static const char tag[8] = "SQHSHv2" ;
  if (fwrite (tag, 8, 1, f) != 1) die ("failed to write seqhash header") ;
  if (fwrite (sh, sizeof (Seqhash), 1, f) != 1) die ("failed to write seqhash") ;
}

Seqhash *seqhashRead (FILE *f)
{
  char tag[8] ;
  Seqhash *sh = (Seqhash*) malloc (sizeof (Seqhash)) ;
  if (!sh) die ("seqhashRead: out of memory") ;
  if (fread (tag, 8, 1, f) != 1) die ("failed to read seqhash header") ;
  if (memcmp (tag, "SQHSHv2", 8)) die ("seqhash read mismatch") ;
  if (fread (sh, sizeof (Seqhash), 1, f) != 1) die ("failed to read seqhash") ;
  return sh ;
}

void seqhashReport (Seqhash *sh, FILE *f)
{ fprintf (f, "SH k %d  w/m %d  s %d\n", sh->k, sh->w, sh->seed) ; }

char *seqString (U64 kmer, int len)
{
  static char buf[33] ;
  if (len < 0 || len > 32) die ("seqString length %d", len) ;
  for (int i = len ; i-- ; kmer >>= 2) buf[i] = "acgt"[kmer & 3] ;
  buf[len] = 0 ;
  return buf ;
}

/* ---- the modimizer iterator ----
 * members reused: hashBuf = the k-mers (bit 63 = isForward), fBuf = the positions (U32, not bool),
 * base = number of results, iStart = next one to hand out */
SeqhashRCiterator *modRCiterator (Seqhash *sh, char *s, int len)
{
  SeqhashRCiterator *it = (SeqhashRCiterator*) calloc (1, sizeof (SeqhashRCiterator)) ;
  if (!it) die ("modRCiterator: out of memory") ;
  it->sh = sh ; it->s = s ; it->sEnd = s + (len > 0 ? len : 0) ;
  if (len < sh->k) { it->isDone = true ; return it ; }                          /* seqhash.c:162 */
  ModgpuScanner *sc = modshimScanner (sh) ;
  uint64_t offs[2] = { 0, (uint64_t) len } ;
  uint64_t cap = (uint64_t) len / (uint64_t) sh->w + (uint64_t) len / (4 * (uint64_t) sh->w) + 1024 ;
  if (cap > (uint64_t) len) cap = (uint64_t) len ;
  for (int attempt = 0 ; attempt < 2 ; ++attempt)
    { free (it->hashBuf) ; free (it->fBuf) ;
      it->hashBuf = (U64*) malloc (cap * sizeof (U64)) ;
      it->fBuf = (bool*) malloc (cap * sizeof (uint32_t)) ;
      if (!it->hashBuf || !it->fBuf) die ("modRCiterator: out of memory") ;
      uint64_t n = modgpuScannerScan (sc, s, offs, 1, 0, (uint64_t*) it->hashBuf, (uint32_t*) it->fBuf, 0, cap) ;
      if (n == UINT64_MAX) die ("%s", (char*) modgpuLastError ()) ;
      if (n <= cap) { it->base = (int) n ; break ; }
      cap = n ;                                                                  /* denser than expected: once more */
    }
  it->iStart = 0 ;
  it->isDone = (it->base == 0) ;
  return it ;
}

bool modRCnext (SeqhashRCiterator *it, U64 *kmer, int *pos, bool *isF)
{
  if (it->isDone) return false ;                                                 /* seqhash.c:181 */
  const U64 v = it->hashBuf[it->iStart] ;
  if (kmer) *kmer = v & 0x3FFFFFFFFFFFFFFFull ;
  if (pos) *pos = (int) ((uint32_t*) it->fBuf)[it->iStart] ;
  if (isF) *isF = (v >> 63) != 0 ;
  it->iMin = (int) ((uint32_t*) it->fBuf)[it->iStart] ;
  if (++it->iStart >= it->base) it->isDone = true ;
  return true ;
}

SeqhashRCiterator *minimizerRCiterator (Seqhash *sh, char *s, int len)
{ die ("minimizerRCiterator: the window-minimizer iterator is not part of the GPU path") ; return 0 ; }

bool minimizerRCnext (SeqhashRCiterator *si, U64 *u, int *pos, bool *isF)
{ die ("minimizerRCnext: the window-minimizer iterator is not part of the GPU path") ; return false ; }
