/* modshim_modset.c - the reference's OWN modset symbols (modset.h:30-42), served by libmodgpu (B200).
 *
 * Compiled against the reference's headers where they lie (-I$(REF)): `Modset` is the reference's struct by
 * construction.  Every Modset created here has a DEVICE TWIN (ModgpuModset, exact first-occurrence numbering) and the
 * two are kept in step so that unmodified caller code keeps working on the public fields:
 *
 *   - the KEY SET lives on the device; value[1..max] and max on the host mirror it after every call that adds keys;
 *   - depth[] and info[] on the host are authoritative whenever control is in caller code (callers do ++ms->depth[i],
 *     msSetCopy*(ms,i) directly): every entry point that computes with them uploads them first
 *     (modgpuModsetSetDepthInfo) and hands the results back (modgpuModsetExport);
 *   - index[] (the reference's own probe table) is only read by modsetWrite: rebuilt on the device on demand
 *     (modgpuModsetReferenceIndex).
 *
 *   modsetCreate / modsetDestroy            modset.c:15-34
 *   modsetIndexFind                         modset.c:45-62   one k-mer per call: a device lookup (isAdd: find-or-insert)
 *   modsetPack                              modset.c:36-43   host array bookkeeping only
 *   modsetDepthPrune                        modset.c:64-77   modgpuModsetPrune
 *   modsetWrite / modsetRead                modset.c:79-104  the file layout; index[] from the device
 *   modsetMerge                             modset.c:106-128 modgpuModsetMerge
 *   modsetSummary                           modset.c:130-153 modgpuModsetSummary
 *
 * modsetIndexFind is exact but pays a kernel launch per k-mer; the throughput path is the batched one:
 *   modshimBatchPut / modshimBatchFlush     == the addSequence loop of modutils.c:19-31 over pinned batches
 * Errors follow the reference: die() (utils.c:19-30).  No CPU fallback: nothing here probes or counts on the host.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "modset.h"                  /* the reference's header ($(REF)) */
#include "modgpu.h"
#include "modshim.h"

#define GDIE() die ("%s", (char*) modgpuLastError ())

typedef struct Twin {
  Modset *ms ;
  ModgpuModset *g ;
  int indexStale ;                   /* ms->index[] does not describe value[1..max] */
  void *feeder ;                     /* the double-buffered pinned batches of modshimBatchPut + their worker thread */
  int inGroup ;                      /* sequences were put since the last flush */
  struct Twin *next ;
} Twin ;

static void feederDestroy (struct Twin *t) ;
static Twin *twins ;

static Twin *twinOf (Modset *ms)
{
  Twin *t ;
  for (t = twins ; t ; t = t->next) if (t->ms == ms) return t ;
  die ("modshim: this Modset was not created by modsetCreate / modsetRead of libmodshim") ;
  return 0 ;
}

ModgpuModset *modshimTwin (void *ms) { return twinOf ((Modset*) ms)->g ; }

/* host depth[] / info[] -> device (the caller may have changed them) */
static void push (Twin *t)
{
  Modset *ms = t->ms ;
  if (ms->max && modgpuModsetSetDepthInfo (t->g, ms->depth + 1, ms->info + 1, ms->max)) GDIE () ;
}

/* device -> host value / depth / info [1..max] and max */
static void pull (Twin *t)
{
  Modset *ms = t->ms ;
  U32 max = modgpuModsetMax (t->g) ;
  if (max == 0xFFFFFFFFu) GDIE () ;
  if (max >= ms->size) die ("hashTableSize %u is too small for %u", ms->size, max) ;      /* modset.c:58 */
  if (max && modgpuModsetExport (t->g, ms->value + 1, ms->depth + 1, ms->info + 1)) GDIE () ;
  if (max != ms->max) t->indexStale = 1 ;
  ms->max = max ;
}

void modshimSync (void *ms) { pull (twinOf ((Modset*) ms)) ; }

static Twin *twinCreate (Modset *ms)
{
  ModgpuHasher h ;
  if (modgpuHasherFromSeqhash (&h, ms->hasher)) GDIE () ;
  Twin *t = (Twin*) calloc (1, sizeof (Twin)) ;
  if (!t) die ("modshim: out of memory") ;
  t->ms = ms ;
  if (!(t->g = modgpuModsetCreateWithHasher (ms->tableBits, &h))) GDIE () ;
  if (modgpuModsetSetExactOrder (t->g, 1)) GDIE () ;            /* index = ++max by first occurrence, modset.c:57 */
  t->next = twins ; twins = t ;
  return t ;
}

Modset *modsetCreate (Seqhash *sh, int bits, U32 size)
{
  if (bits < 20 || bits > 34) die ("table bits %d must be between 20 and 34", bits) ;     /* modset.c:17 */
  Modset *ms = (Modset*) calloc (1, sizeof (Modset)) ;
  if (!ms) die ("modsetCreate: out of memory") ;
  ms->hasher = sh ;
  ms->tableBits = bits ;
  ms->tableSize = (U64)1 << bits ;
  ms->tableMask = ms->tableSize - 1 ;
  if (size >= (ms->tableSize >> 2)) die ("Modset size %u is too big for %d bits", size, bits) ;   /* modset.c:24 */
  ms->size = size ? size : (U32) ((ms->tableSize >> 2) - 1) ;
  ms->index = (U32*) calloc (ms->tableSize, sizeof (U32)) ;
  ms->value = (U64*) malloc ((size_t) ms->size * sizeof (U64)) ;
  ms->depth = (U16*) calloc (ms->size, sizeof (U16)) ;
  ms->info = (U8*) calloc (ms->size, sizeof (U8)) ;
  if (!ms->index || !ms->value || !ms->depth || !ms->info) die ("modsetCreate: out of memory") ;
  twinCreate (ms) ;
  return ms ;
}

void modsetDestroy (Modset *ms)
{
  Twin **p, *t = 0 ;
  for (p = &twins ; *p ; p = &(*p)->next) if ((*p)->ms == ms) { t = *p ; *p = t->next ; break ; }
  if (t)
    { feederDestroy (t) ;
      modgpuModsetDestroy (t->g) ;
      free (t) ;
    }
  free (ms->index) ; free (ms->value) ; free (ms->depth) ; free (ms->info) ; free (ms) ;
}

bool modsetPack (Modset *ms)                      /* modset.c:36-43: shrink the per-item arrays to max+1 */
{
  if (ms->size == ms->max + 1) return false ;
  size_t n = (size_t) ms->max + 1 ;
  ms->value = (U64*) realloc (ms->value, n * sizeof (U64)) ;
  ms->depth = (U16*) realloc (ms->depth, n * sizeof (U16)) ;
  ms->info = (U8*) realloc (ms->info, n * sizeof (U8)) ;
  if (!ms->value || !ms->depth || !ms->info) die ("modsetPack: out of memory") ;
  ms->size = (U32) n ;
  return true ;
}

U32 modsetIndexFind (Modset *ms, U64 kmer, int isAdd)
{
  Twin *t = twinOf (ms) ;
  uint64_t k = kmer ; uint32_t index = 0 ;
  if (modgpuModsetIndexFindBatch (t->g, &k, 1, 0, &index)) GDIE () ;
  if (index || !isAdd) return index ;
  if (ms->max + 1 >= ms->size) die ("hashTableSize %u is too small for %u", ms->size, ms->max + 1) ;   /* modset.c:58 */
  if (modgpuModsetIndexFindBatch (t->g, &k, 1, 1, &index)) GDIE () ;
  if (index != ms->max + 1) die ("modshim: device numbering %u out of step with the host (%u)", index, ms->max + 1) ;
  ms->max = index ;
  ms->value[index] = kmer ;
  t->indexStale = 1 ;
  return index ;
}

void modsetDepthPrune (Modset *ms, int min, int max)
{
  Twin *t = twinOf (ms) ;
  U32 N = ms->max ;
  push (t) ;
  if (modgpuModsetPrune (t->g, min, max)) GDIE () ;
  pull (t) ;
  t->indexStale = 1 ;
  fprintf (stderr, "  pruned Modset from %d to %d with min %d <= depth < max %d\n", N, ms->max, min, max) ;   /* modset.c:75-76 */
}

void modsetWrite (Modset *ms, FILE *f)
{
  Twin *t = twinOf (ms) ;
  if (t->indexStale)
    { if (modgpuModsetReferenceIndex (t->g, ms->index)) GDIE () ;
      t->indexStale = 0 ;
    }
  if (fwrite ("MSHSTv2",8,1,f) != 1) die ("failed to write modset header") ;
  if (fwrite (&ms->tableBits,sizeof(int),1,f) != 1) die ("failed to write bits") ;
  U32 size = ms->max+1 ; if (fwrite (&size,sizeof(U32),1,f) != 1) die ("failed to write size") ;
  seqhashWrite (ms->hasher, f) ;
  if (fwrite (ms->index,sizeof(U32),ms->tableSize,f) != ms->tableSize) die ("fail write index") ;
  if (fwrite (ms->value,sizeof(U64),size,f) != size) die ("failed to write value") ;
  if (fwrite (ms->depth,sizeof(U16),size,f) != size) die ("failed to write depth") ;
  if (fwrite (ms->info,sizeof(U8),size,f) != size) die ("failed to write info") ;
}

Modset *modsetRead (FILE *f)
{
  char name[8] ;
  if (fread (name,8,1,f) != 1) die ("failed to read modset header") ;
  if (memcmp (name, "MSHSTv2", 8)) die ("bad modset header %.8s != MSHSTv2", name) ;
  int bits ; if (fread (&bits,sizeof(int),1,f) != 1) die ("failed to read bits") ;
  U32 size ; if (fread (&size,sizeof(U32),1,f) != 1) die ("failed to read size") ;
  if (!size) die ("bad modset size 0") ;
  Seqhash *sh = seqhashRead (f) ;
  Modset *ms = modsetCreate (sh, bits, size) ;
  if (fread (ms->index,sizeof(U32),ms->tableSize,f) != ms->tableSize) die ("failed read index") ;
  if (fread (ms->value,sizeof(U64),size,f) != size) die ("failed to read value") ;
  if (fread (ms->depth,sizeof(U16),size,f) != size) die ("failed to read depth") ;
  if (fread (ms->info,sizeof(U8),size,f) != size) die ("failed to read info") ;
  ms->max = size - 1 ;
  Twin *t = twinOf (ms) ;
  if (ms->max && modgpuModsetImport (t->g, ms->value + 1, ms->depth + 1, ms->info + 1, ms->max)) GDIE () ;
  return ms ;
}

bool modsetMerge (Modset *ms1, Modset *ms2)
{
  Twin *t1 = twinOf (ms1), *t2 = twinOf (ms2) ;
  Seqhash *sh1 = ms1->hasher, *sh2 = ms2->hasher ;
  if (sh1->w != sh2->w || sh1->k != sh2->k || sh1->factor1 != sh2->factor1) return false ;   /* modset.c:111 */
  U64 newSize = (U64) ms1->max + ms2->max + 1 ;                                               /* modset.c:113-118 */
  if (newSize >= (ms1->tableSize >> 2)) newSize = (ms1->tableSize >> 2) - 1 ;
  { U32 old = ms1->size ;                           /* resize (utils.h:54) to newSize, growing or shrinking */
    ms1->value = (U64*) realloc (ms1->value, newSize * sizeof (U64)) ;
    ms1->depth = (U16*) realloc (ms1->depth, newSize * sizeof (U16)) ;
    ms1->info = (U8*) realloc (ms1->info, newSize * sizeof (U8)) ;
    if (!ms1->value || !ms1->depth || !ms1->info) die ("modsetMerge: out of memory") ;
    if (newSize > old)
      { memset (ms1->depth + old, 0, (newSize - old) * sizeof (U16)) ;
        memset (ms1->info + old, 0, (newSize - old) * sizeof (U8)) ;
      }
    ms1->size = (U32) newSize ;
  }
  push (t1) ; push (t2) ;
  int rc = modgpuModsetMerge (t1->g, t2->g) ;
  if (rc < 0) GDIE () ;
  pull (t1) ;
  t1->indexStale = 1 ;
  return rc == 1 ;
}

void modsetSummary (Modset *ms, FILE *f)
{
  Twin *t = twinOf (ms) ;
  char buf[1024] ;
  push (t) ;
  if (modgpuModsetSummary (t->g, buf, sizeof (buf)) < 0) GDIE () ;
  fputs (buf, f) ;
}

/* ---- the batched add: == the addSequence loop of modutils.c:19-31 --------------------------------------------
 * seqio reuses its buffer on the next seqIOread (seqio.h:46-48), so every sequence is copied into a pinned batch.
 * Two batches: while the caller's thread parses into one (seqIOread is the slow side: ~0.17 Gbases/s per core), a
 * worker thread feeds the other through modgpuModsetAdd (H2D + kernels), so the file reader never waits for the GPU
 * and the GPU never waits for more than one batch.  Only the worker touches the twin between two flushes.
 * modshimBatchFlush adds what is left, joins the worker, syncs the host arrays and returns the number of hashes added
 * since the previous flush (the reference's totHash). */
#include <pthread.h>

typedef struct { char *bases ; U64 *off ; size_t cap, used, nSeq ; } Half ;

typedef struct Feeder {
  Twin *t ;
  Half h[2] ; int cur ;              /* the half being filled by the caller */
  size_t capSeq ;
  pthread_t th ; int started ;
  pthread_mutex_t mu ; pthread_cond_t cv ;
  int job ;                          /* half handed to the worker, -1 none, -2 quit */
  int busy ;
  U64 hashes ;
  char err[512] ;
} Feeder ;

static void *feederMain (void *arg)
{
  Feeder *f = (Feeder*) arg ;
  modgpuSetDevice (modgpuModsetDevice (f->t->g)) ;       /* the current device is per thread */
  pthread_mutex_lock (&f->mu) ;
  for (;;)
    { while (f->job == -1) pthread_cond_wait (&f->cv, &f->mu) ;
      if (f->job == -2) break ;
      Half *h = &f->h[f->job] ;
      pthread_mutex_unlock (&f->mu) ;
      U64 n = modgpuModsetAdd (f->t->g, h->bases, (uint64_t*) h->off, h->nSeq, 0) ;     /* 0: the bytes are codes 0..3 */
      pthread_mutex_lock (&f->mu) ;
      if (n == UINT64_MAX) { if (!f->err[0]) snprintf (f->err, sizeof (f->err), "%s", modgpuLastError ()) ; }
      else f->hashes += n ;
      h->used = 0 ; h->nSeq = 0 ;
      f->job = -1 ; f->busy = 0 ;
      pthread_cond_broadcast (&f->cv) ;
    }
  pthread_mutex_unlock (&f->mu) ;
  return 0 ;
}

static void feederWait (Feeder *f)                /* until the worker is idle; its error becomes ours */
{
  pthread_mutex_lock (&f->mu) ;
  while (f->busy) pthread_cond_wait (&f->cv, &f->mu) ;
  pthread_mutex_unlock (&f->mu) ;
  if (f->err[0]) die ("%s", f->err) ;
}

static void feederSend (Feeder *f)                /* hand the current half to the worker, continue in the other */
{
  if (!f->h[f->cur].nSeq) return ;
  feederWait (f) ;                                /* the other half is free again */
  pthread_mutex_lock (&f->mu) ;
  f->job = f->cur ; f->busy = 1 ;
  pthread_cond_broadcast (&f->cv) ;
  pthread_mutex_unlock (&f->mu) ;
  f->cur ^= 1 ;
}

static Feeder *feederOf (Twin *t)
{
  if (t->feeder) return (Feeder*) t->feeder ;
  Feeder *f = (Feeder*) calloc (1, sizeof (Feeder)) ;
  if (!f) die ("modshim: out of memory") ;
  f->t = t ; f->job = -1 ; f->capSeq = 1 << 22 ;
  pthread_mutex_init (&f->mu, 0) ; pthread_cond_init (&f->cv, 0) ;
  for (int i = 0 ; i < 2 ; ++i)
    { f->h[i].cap = (size_t) 1 << 28 ;
      if (!(f->h[i].bases = (char*) modgpuHostAlloc (f->h[i].cap))) GDIE () ;
      if (!(f->h[i].off = (U64*) calloc (f->capSeq + 1, sizeof (U64)))) die ("modshim: out of memory") ;
    }
  if (pthread_create (&f->th, 0, feederMain, f)) die ("modshim: cannot start the feeder thread") ;
  f->started = 1 ;
  t->feeder = f ;
  return f ;
}

static void feederDestroy (Twin *t)
{
  Feeder *f = (Feeder*) t->feeder ;
  if (!f) return ;
  pthread_mutex_lock (&f->mu) ;
  while (f->busy) pthread_cond_wait (&f->cv, &f->mu) ;
  f->job = -2 ;
  pthread_cond_broadcast (&f->cv) ;
  pthread_mutex_unlock (&f->mu) ;
  pthread_join (f->th, 0) ;
  for (int i = 0 ; i < 2 ; ++i) { modgpuHostFree (f->h[i].bases) ; free (f->h[i].off) ; }
  free (f) ; t->feeder = 0 ;
}

void modshimBatchPut (void *vms, const char *s, long long len)
{
  Twin *t = twinOf ((Modset*) vms) ;
  Feeder *f = feederOf (t) ;
  if (len < 0) len = 0 ;                           /* int len < k has no k-mer (seqhash.c:162) */
  if (!t->inGroup) { push (t) ; t->inGroup = 1 ; } /* first sequence of a group: the caller's depths are current */
  Half *h = &f->h[f->cur] ;
  if (h->used + (size_t) len > h->cap || h->nSeq == f->capSeq)
    { feederSend (f) ; h = &f->h[f->cur] ; }
  if ((size_t) len > h->cap)                       /* one sequence longer than a batch: grow this half (it is empty) */
    { feederWait (f) ;
      modgpuHostFree (h->bases) ;
      h->cap = (size_t) len + (size_t) len / 4 ;
      if (!(h->bases = (char*) modgpuHostAlloc (h->cap))) GDIE () ;
    }
  memcpy (h->bases + h->used, s, (size_t) len) ;
  h->used += (size_t) len ; h->off[++h->nSeq] = h->used ;
}

unsigned long long modshimBatchFlush (void *vms)
{
  Twin *t = twinOf ((Modset*) vms) ;
  U64 n = 0 ;
  if (t->feeder)
    { Feeder *f = (Feeder*) t->feeder ;
      feederSend (f) ;
      feederWait (f) ;
      n = f->hashes ; f->hashes = 0 ;
    }
  t->inGroup = 0 ;                                 /* the next group pushes the caller's depths again */
  pull (t) ;
  return n ;
}
