/* modmap_gpu.c - a modmap whose index build and seed lookup run on the B200, as a C host over libmodgpu.
 *
 * Our own command interpreter (not the reference's main).  The reference's seqio parses the files, its
 * dict / array objects (compiled in place from $(REF)) hold the sequence names and lengths so that the
 * .ref file has the reference's own layout, and include/modgpu.h does the work:
 *
 *   -K k  -W w  -S seed  -B tableBits  -v  -o FILE
 *   -f genome.fa     modgpuReferenceBuild  == referenceFastaRead + referencePack   modmap.c:93-134, 74-91
 *   -w root          root.mod + root.ref in the reference's formats                modmap.c:136-156
 *   -q reads.fa      modgpuReferenceQuery  == the seed loop of queryProcess        modmap.c:196-231
 *                    prints the Q line of every read and, with -v, its seed lines; the colinear-block
 *                    ("M" line) pass of modmap.c:232-276 is serial per read and runs here on the host over
 *                    the seed lists the GPU returned (blockScan below)
 *
 * tests/test_gpu_cli.py: same lines and same .ref bytes as the stock modmap, and the stock modmap answers
 * queries from the files written here exactly as from its own.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "seqio.h"                   /* the reference's headers ($(REF)) */
#include "dict.h"
#include "array.h"
#include "modgpu.h"

#define CHECK(x) do { if (x) die ("%s", (char*) modgpuLastError ()) ; } while (0)

static FILE *out ;
static int verbose ;
static ModgpuReference *gref ;
static DICT *names ;                 /* reference sequence names; ids are dict indices from 0 (dict.c:166-170) */
static Array lens ;                  /* of U32, indexed by the dict index */

/* a growing pinned batch of sequences */
typedef struct { char *bases ; U64 *off ; size_t cap, used, nSeq, capSeq ; } Batch ;

static void batchInit (Batch *b, size_t cap, size_t capSeq)
{
  memset (b, 0, sizeof (*b)) ;
  b->cap = cap ; b->capSeq = capSeq ;
  if (!(b->bases = (char*) modgpuHostAlloc (cap))) die ("%s", (char*) modgpuLastError ()) ;
  b->off = (U64*) calloc (capSeq + 1, sizeof (U64)) ;
}

static void batchGrow (Batch *b, size_t need)
{
  if (b->used + need > b->cap)
    { size_t cap = 2 * b->cap ; while (b->used + need > cap) cap *= 2 ;
      char *p = (char*) modgpuHostAlloc (cap) ;
      if (!p) die ("%s", (char*) modgpuLastError ()) ;
      memcpy (p, b->bases, b->used) ; modgpuHostFree (b->bases) ; b->bases = p ; b->cap = cap ;
    }
  if (b->nSeq == b->capSeq)
    { b->capSeq *= 2 ; b->off = (U64*) realloc (b->off, (b->capSeq + 1) * sizeof (U64)) ; }
}

static void batchPut (Batch *b, const char *s, U64 len)
{ batchGrow (b, len) ; memcpy (b->bases + b->used, s, len) ; b->used += len ; b->off[++b->nSeq] = b->used ; }

static void buildReference (char *filename, int k, int w, int seed, int bits)
{
  dna2indexConv['N'] = dna2indexConv['n'] = 0 ;                                 /* modmap.c:97 */
  SeqIO *si = seqIOopenRead (filename, dna2indexConv, false) ;
  if (!si) die ("failed to read reference sequence file %s", filename) ;
  names = dictCreate (1024) ; lens = arrayCreate (1024, U32) ;                  /* as referenceCreate, modmap.c:59-60 */
  Batch b ; batchInit (&b, (size_t) 1 << 28, 1024) ;
  U64 totLen = 0 ;
  while (seqIOread (si))
    { int id ;
      if (!dictAdd (names, sqioId(si), &id)) die ("duplicate ref sequence name %s", sqioId(si)) ;
      array (lens, id, int) = si->seqLen ;
      totLen += si->seqLen ;
      batchPut (&b, sqioSeq(si), si->seqLen) ;
    }
  seqIOclose (si) ;
  uint32_t counts[4] ;
  gref = modgpuReferenceBuild (bits, k, w, seed, b.bases, (uint64_t*) b.off, b.nSeq, 0, counts) ;
  if (!gref) die ("%s", (char*) modgpuLastError ()) ;                            /* incl. "reference size overflow" */
  fprintf (out, "  %d hashes from %d reference sequences, total length %lld\n", (int) counts[0], dictMax (names), (long long) totLen) ;
  fprintf (out, "  %d copy 1, %d copy 2, %d multiple\n", (int) counts[1], (int) counts[2], (int) counts[3]) ;
  modgpuHostFree (b.bases) ; free (b.off) ;
}

static void put (const void *p, size_t size, size_t n, FILE *f, const char *what)
{ if (fwrite (p, size, n, f) != n) die ("failed to write %s", (char*) what) ; }

static void writeReference (char *root)
{
  char path[4096] ;
  snprintf (path, sizeof (path), "%s.mod", root) ;
  CHECK (modgpuModsetWriteMod (modgpuReferenceModset (gref), path, 1)) ;        /* fopenTag -> fzopen: gzip'd, utils.c:108-139 */
  FILE *f = fopenTag (root, "ref", "w") ;                                        /* the reference's own gzip stream */
  if (!f) die ("failed to open %s.ref to write", root) ;
  U32 n = modgpuReferenceMax (gref), m = modgpuModsetMax (modgpuReferenceModset (gref)) + 1 ;
  U32 *index = (U32*) calloc ((size_t) n + 1, 4), *offset = (U32*) calloc ((size_t) n + 1, 4), *id = (U32*) calloc ((size_t) n + 1, 4),
      *rev = (U32*) calloc ((size_t) n + 1, 4), *depth = (U32*) calloc (m, 4), *loc = (U32*) calloc (m, 4) ;
  CHECK (modgpuReferenceExport (gref, index, offset, id, depth, rev, loc)) ;
  put ("RFMSHv1", 8, 1, f, "reference header") ;
  put (&n, sizeof (U32), 1, f, "size") ; put (&n, sizeof (U32), 1, f, "max") ;
  put (index, sizeof (U32), n, f, "ref index") ; put (offset, sizeof (U32), n, f, "ref offset") ; put (id, sizeof (U32), n, f, "ref id") ;
  put (depth, sizeof (U32), m, f, "depth") ; put (rev, sizeof (U32), n, f, "rev") ; put (loc, sizeof (U32), m, f, "loc") ;
  if (!arrayWrite (lens, f)) die ("failed write ref len") ;
  if (!dictWrite (names, f)) die ("failed write ref dict") ;
  fclose (f) ;
  free (index) ; free (offset) ; free (id) ; free (rev) ; free (depth) ; free (loc) ;
}

/* host copies of the occurrence arrays, fetched once for the block pass */
static U32 *hostLoc, *hostRev, *hostId, *hostOffset ;

static void fetchOccurrences (void)
{
  if (hostLoc) return ;
  U32 n = modgpuReferenceMax (gref), m = modgpuModsetMax (modgpuReferenceModset (gref)) + 1 ;
  U32 *index = (U32*) calloc ((size_t) n + 1, 4), *depth = (U32*) calloc (m, 4) ;
  hostOffset = (U32*) calloc ((size_t) n + 1, 4) ; hostId = (U32*) calloc ((size_t) n + 1, 4) ;
  hostRev = (U32*) calloc ((size_t) n + 1, 4) ; hostLoc = (U32*) calloc (m, 4) ;
  CHECK (modgpuReferenceExport (gref, index, hostOffset, hostId, depth, hostRev, hostLoc)) ;
  free (index) ; free (depth) ;
}

/* The colinear-block pass over the seeds of one read (modmap.c:213-276), restated as a scanner with explicit
 * state.  A block is a run of unique / two-copy seeds whose occurrences lie on one reference sequence, move in
 * one direction through the occurrence list and keep pace with the seed ordinals to within 50.  Occurrence
 * ordinal 0 doubles as "no block open" in the reference (modmap.c:233), kept.  All arithmetic is 32-bit
 * unsigned as there.  The reference reports a finished block when it holds more than two unique seeds, and
 * the block still open at the end of the read when it holds more than two TWO-COPY seeds (modmap.c:266) -
 * reproduced as is. */
typedef struct { U32 first, last, iFirst, iLast ; int nUnique, nPair ; } Block ;

static int leavesBlock (const Block *b, U32 occ)
{
  if (hostId[occ] != hostId[b->first]) return 1 ;
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

static void blockPrint (const Block *b, const char *readName, const uint32_t *sPos, int nCopy1)
{
  U32 span = b->last > b->first ? b->last - b->first : b->first - b->last ;
  fprintf (out, "M\t%s\t%d\t%d\t%d\t%s\t%d\t%d\t%d %d\t%.2f\t%.2f\n", readName,
           (int) sPos[b->iFirst], (int) sPos[b->iLast], (int) (sPos[b->iLast] - sPos[b->iFirst]),
           dictName (names, hostId[b->first]), (int) hostOffset[b->first], (int) hostOffset[b->last],
           b->nUnique, b->nPair, (b->nUnique + b->nPair) / (double) span, b->nUnique / (double) nCopy1) ;
}

static void blockScan (const char *readName, const uint32_t *sIndex, const uint32_t *sPos, const uint32_t *hId,
                       const uint32_t *hOff, U32 nSeed, int nCopy1)
{
  Block b ; memset (&b, 0, sizeof (b)) ;
  for (U32 i = 0 ; i < nSeed ; ++i)
    { if (hId[2*i] == 0xFFFFFFFFu) continue ;                                    /* miss or multi-copy */
      int unique = hId[2*i+1] == 0xFFFFFFFFu ;
      if (verbose && unique)                                                    /* seed lines go to stdout, modmap.c:221-229 */
        printf ("  %6d\t%s %d\n", (int) sPos[i], dictName (names, hId[2*i]), (int) hOff[2*i]) ;
      else if (verbose)
        printf ("  %6d\t%s %d\t%s %d\n", (int) sPos[i], dictName (names, hId[2*i]), (int) hOff[2*i],
                dictName (names, hId[2*i+1]), (int) hOff[2*i+1]) ;
      U32 at = hostLoc[sIndex[i]], occ = hostRev[at] ;
      int ends = !b.first || leavesBlock (&b, occ) ;
      if (ends && b.first && !unique)                                           /* the other copy may continue it */
        { occ = hostRev[at + 1] ; ends = leavesBlock (&b, occ) ; }
      if (ends)
        { if (b.nUnique > 2) blockPrint (&b, readName, sPos, nCopy1) ;
          b.nUnique = b.nPair = 0 ; b.first = occ ; b.iFirst = i ;
        }
      if (unique) ++b.nUnique ; else ++b.nPair ;
      b.last = occ ; b.iLast = i ;
    }
  if (b.nPair > 2) blockPrint (&b, readName, sPos, nCopy1) ;
}

/* one batch of reads through the seed loop, printed read by read */
static void queryFlush (Batch *b, char **ids, U64 *lenOf)
{
  if (!b->nSeq) return ;
  fetchOccurrences () ;
  uint64_t cap = b->used + 16 ;
  uint64_t *seedOff = (uint64_t*) calloc (b->nSeq + 1, 8) ;
  uint32_t *sIndex = (uint32_t*) malloc (cap * 4), *sPos = (uint32_t*) malloc (cap * 4),
           *hId = (uint32_t*) malloc (cap * 8), *hOff = (uint32_t*) malloc (cap * 8) ;
  int32_t *ctr = (int32_t*) calloc (b->nSeq * 4, 4) ;
  uint64_t n = modgpuReferenceQuery (gref, b->bases, (uint64_t*) b->off, b->nSeq, 0, seedOff, sIndex, sPos, hId, hOff, ctr, cap) ;
  if (n == UINT64_MAX) die ("%s", (char*) modgpuLastError ()) ;
  for (size_t r = 0 ; r < b->nSeq ; ++r)
    { int32_t *c = ctr + 4*r ;                                                   /* { miss, copy1, copy2, multi } */
      uint64_t a = seedOff[r], e = seedOff[r+1], ns = e - a ;
      fprintf (out, "Q\t%s\t%llu\t%d miss, %d copy1, %d copy2, %d multi, %.2f hit\n",
               ids[r], (unsigned long long) lenOf[r], c[0], c[1], c[2], c[3], (int)(ns - c[0]) / (double)(int) ns) ;
      blockScan (ids[r], sIndex + a, sPos + a, hId + 2*a, hOff + 2*a, (U32) ns, c[1]) ;
      free (ids[r]) ;
    }
  free (seedOff) ; free (sIndex) ; free (sPos) ; free (hId) ; free (hOff) ; free (ctr) ;
  b->used = 0 ; b->nSeq = 0 ;
}

static void queryFile (char *filename)
{
  dna2indexConv['N'] = dna2indexConv['n'] = 0 ;                                 /* modmap.c:193 */
  SeqIO *si = seqIOopenRead (filename, dna2indexConv, false) ;
  if (!si) die ("failed to read query sequence file %s", filename) ;
  const size_t maxReads = 1 << 16 ;
  Batch b ; batchInit (&b, (size_t) 1 << 27, maxReads) ;
  char **ids = (char**) calloc (maxReads, sizeof (char*)) ;
  U64 *lenOf = (U64*) calloc (maxReads, sizeof (U64)) ;
  while (seqIOread (si))
    { if (b.nSeq == maxReads || b.used + si->seqLen > ((size_t) 1 << 27)) queryFlush (&b, ids, lenOf) ;
      ids[b.nSeq] = strdup (sqioId(si)) ; lenOf[b.nSeq] = si->seqLen ;
      batchPut (&b, sqioSeq(si), si->seqLen) ;
    }
  queryFlush (&b, ids, lenOf) ;
  seqIOclose (si) ;
  modgpuHostFree (b.bases) ; free (b.off) ; free (ids) ; free (lenOf) ;
}

int main (int argc, char **argv)
{
  int k = 19, w = 31, seed = 17, bits = 28 ;                                     /* modmap.c:314-317 */
  out = stdout ;
  timeUpdate (stdout) ;
  for (int a = 1 ; a < argc ; )
    { char *cmd = argv[a] ;
      if (cmd[0] != '-') die ("option/command %s does not start with '-'", cmd) ;
      int nArg = 0 ;
      while (a + 1 + nArg < argc && argv[a + 1 + nArg][0] != '-') ++nArg ;
      fprintf (stderr, "COMMAND %s", cmd) ;
      for (int i = 1 ; i <= nArg ; ++i) fprintf (stderr, " %s", argv[a + i]) ;
      fputc ('\n', stderr) ;
      char **arg = argv + a + 1 ;
#define IS(x,y) (!strcmp (cmd, x) || !strcmp (cmd, y))
#define NEED(n) do { if (nArg < (n)) die ("command %s needs %d arguments", cmd, (n)) ; } while (0)
      if (IS ("-K", "--kmer")) { NEED (1) ; k = atoi (arg[0]) ; }
      else if (IS ("-W", "--window")) { NEED (1) ; w = atoi (arg[0]) ; }
      else if (IS ("-S", "--seed")) { NEED (1) ; seed = atoi (arg[0]) ; }
      else if (IS ("-B", "--tableBits")) { NEED (1) ; bits = atoi (arg[0]) ; }
      else if (IS ("-v", "--verbose")) verbose = !verbose ;
      else if (IS ("-o", "--output"))
        { NEED (1) ;
          if (!strcmp (arg[0], "-")) out = stdout ;
          else if (!(out = fopen (arg[0], "w")))
            { fprintf (stderr, "can't open output file %s - resetting to stdout\n", arg[0]) ; out = stdout ; }
        }
      else if (IS ("-f", "--referenceFasta"))
        { NEED (1) ;
          if (k <= 0 || w <= 0) die ("k %d, w %d must be > 0", k, w) ;
          fprintf (out, "  modmap initialised with k = %d, w = %d, random seed = %d\n", k, w, seed) ;
          buildReference (arg[0], k, w, seed, bits) ;
        }
      else if (IS ("-w", "--referenceWrite"))
        { NEED (1) ; if (!gref) die ("need to read a reference before writing it") ; writeReference (arg[0]) ; }
      else if (IS ("-q", "--query"))
        { NEED (1) ; if (!gref) die ("need to read a reference before processing query sequences") ; queryFile (arg[0]) ; }
      else if (IS ("-r", "--referenceRead"))
        die ("-r is not part of this driver: build with -f, or let the stock modmap -r read the files written by -w") ;
      else die ("unknown command %s", cmd) ;
      a += 1 + nArg ;
      timeUpdate (out) ;
    }
  fprintf (out, "total resources used: ") ; timeTotal (out) ;
  if (out != stdout) { printf ("total resources used: ") ; timeTotal (stdout) ; }
  if (gref) modgpuReferenceDestroy (gref) ;
  return 0 ;
}
