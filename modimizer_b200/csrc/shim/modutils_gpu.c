/* modutils_gpu.c - a modutils whose hot path runs on the B200, as a C host over libmodgpu.
 *
 * Our own command interpreter (not the reference's main), for the commands that sit on the
 * hot path: the reference's seqio (compiled in place from $(REF)) parses the files on the host
 * and fills pinned batches, the C ABI of include/modgpu.h does everything else, and every line
 * of output has the format of the reference tool - tests/test_gpu_cli.py
 * compares the files it writes with the ones the stock modutils writes, byte for byte.
 *
 *   -o FILE                   output file ('-' = stdout)
 *   -c [B [k [w [seed]]]]     create (modutils.c:139-157)          -r FILE   read a .mod
 *   -a FILE | -x FILE         add reads / 10x reads (modutils.c:33-51)
 *   -H FILE                   depth histogram (modutils.c:53-63)
 *   -s c1 c2 cM | -sM m       copy classes (modutils.c:205-219)
 *   -p min max                prune (modset.c:64-77)                -m FILE   merge a .mod (modset.c:106-128)
 *   -w FILE | -wt FILE        write .mod (gzip, like fzopen) / text dump (modutils.c:191-200)
 *   -P FILE                   refpaint (modutils.c:260-273)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "seqio.h"                   /* the reference's headers ($(REF)) */
#include "seqhash.h"
#include "modgpu.h"

ModgpuScanner *modshimScanner (Seqhash *sh) ;

static FILE *out ;
static ModgpuModset *gms ;
static Seqhash *hasher ;
static int tableBits ;

#define CHECK(x) do { if (x) die ("%s", (char*) modgpuLastError ()) ; } while (0)

static void summary (ModgpuModset *ms)
{ char buf[1024] ;
  if (modgpuModsetSummary (ms, buf, sizeof (buf)) < 0) die ("%s", (char*) modgpuLastError ()) ;   /* returns bytes written */
  fputs (buf, out) ;
}

/* one pinned batch of reads: bases + offsets, flushed through modgpuModsetAdd */
typedef struct { char *bases ; U64 *off ; size_t cap, used, nSeq, capSeq ; U64 hashes ; } Batch ;

static void batchFlush (Batch *b)
{
  if (!b->nSeq) return ;
  U64 n = modgpuModsetAdd (gms, b->bases, (uint64_t*) b->off, b->nSeq, 0) ;     /* 0: the bytes are codes 0..3 */
  if (n == UINT64_MAX) die ("%s", (char*) modgpuLastError ()) ;
  b->hashes += n ; b->used = 0 ; b->nSeq = 0 ;
}

static void batchPut (Batch *b, const char *s, U64 len)
{
  if (len > b->cap) { batchFlush (b) ; modgpuHostFree (b->bases) ; b->cap = len + len / 4 ;
                      if (!(b->bases = (char*) modgpuHostAlloc (b->cap))) die ("%s", (char*) modgpuLastError ()) ; }
  if (b->used + len > b->cap || b->nSeq == b->capSeq) batchFlush (b) ;
  memcpy (b->bases + b->used, s, len) ;      /* seqio reuses its buffer on the next seqIOread (seqio.h:46-48) */
  b->used += len ; b->off[++b->nSeq] = b->used ;
}

static void addFile (char *filename, int is10x)
{
  Batch b ; memset (&b, 0, sizeof (b)) ;
  b.cap = (size_t) 1 << 28 ; b.capSeq = 1 << 22 ;
  if (!(b.bases = (char*) modgpuHostAlloc (b.cap))) die ("%s", (char*) modgpuLastError ()) ;
  b.off = (U64*) calloc (b.capSeq + 1, sizeof (U64)) ;
  dna2indexConv['N'] = dna2indexConv['n'] = 0 ;                                 /* modutils.c:39 */
  SeqIO *si = seqIOopenRead (filename, dna2indexConv, false) ;
  if (!si) die ("failed to open sequence file %s", filename) ;
  U64 nSeq = 0, totLen = 0 ;
  while (seqIOread (si))
    { ++nSeq ; totLen += si->seqLen ;
      if (is10x && (nSeq & 1))                                                    /* modutils.c:44: int len = seqLen-23; */
        { if (si->seqLen > 23) batchPut (&b, sqioSeq(si) + 23, si->seqLen - 23) ;   /* a negative length has no k-mer (seqhash.c:162) */
          else batchPut (&b, sqioSeq(si), 0) ;
        }
      else batchPut (&b, sqioSeq(si), si->seqLen) ;
    }
  batchFlush (&b) ;
  seqIOclose (si) ;
  fprintf (out, "added %llu sequences total length %llu total hashes %llu, new max %u\n",
           (unsigned long long) nSeq, (unsigned long long) totLen, (unsigned long long) b.hashes, modgpuModsetMax (gms)) ;
  modgpuHostFree (b.bases) ; free (b.off) ;
}

/* sync-to-host: value/depth/info[1..max] in the reference's index order */
static U32 exportSet (U64 **value, U16 **depth, U8 **info)
{
  U32 max = modgpuModsetMax (gms) ;
  *value = (U64*) malloc (((size_t) max + 1) * sizeof (U64)) ;
  *depth = (U16*) malloc (((size_t) max + 1) * sizeof (U16)) ;
  *info = (U8*) malloc ((size_t) max + 1) ;
  if (max) CHECK (modgpuModsetExport (gms, *value + 1, *depth + 1, *info + 1)) ;
  return max ;
}

static void refPaint (char *filename)
{
  SeqIO *si = seqIOopenRead (filename, dna2indexConv, false) ;
  if (!si) die ("failed to open ref seq file %s", filename) ;
  U64 *value ; U16 *depth ; U8 *info ;
  exportSet (&value, &depth, &info) ;
  ModgpuScanner *sc = modshimScanner (hasher) ;
  while (seqIOread (si))
    { printf ("painting %s length %d\n", sqioId(si), (int) si->seqLen) ;
      if ((int) si->seqLen < hasher->k) continue ;
      uint64_t offs[2] = { 0, si->seqLen }, cap = si->seqLen ;
      uint64_t *km = (uint64_t*) malloc (cap * 8) ; uint32_t *pos = (uint32_t*) malloc (cap * 4) ;
      uint64_t n = modgpuScannerScan (sc, sqioSeq(si), offs, 1, 0, km, pos, 0, cap) ;
      if (n == UINT64_MAX) die ("%s", (char*) modgpuLastError ()) ;
      uint32_t *index = (uint32_t*) malloc ((n + 1) * 4) ;
      for (uint64_t i = 0 ; i < n ; ++i) km[i] &= 0x3FFFFFFFFFFFFFFFull ;
      if (n) CHECK (modgpuModsetFind (gms, km, n, index, 0)) ;
      for (uint64_t i = 0 ; i < n ; ++i)
        if (index[i]) printf ("  %d\t%d\n", (int) pos[i], depth[index[i]]) ;
      free (km) ; free (pos) ; free (index) ;
    }
  seqIOclose (si) ;
  free (value) ; free (depth) ; free (info) ;
}

int main (int argc, char **argv)
{
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
#define NEEDSET() do { if (!gms) die ("command %s needs a modset: -c or -r first", cmd) ; } while (0)
      if (IS ("-o", "--output"))
        { NEED (1) ;
          if (!strcmp (arg[0], "-")) out = stdout ;
          else if (!(out = fopen (arg[0], "w")))
            { fprintf (stderr, "can't open output file %s - resetting to stdout\n", arg[0]) ; out = stdout ; }
        }
      else if (IS ("-c", "--create") && !gms)
        { int B = 28, k = 19, w = 31, s = 17 ;                                   /* modutils.c:140 */
          if (nArg > 0 && (!(B = atoi (arg[0])) || B < 20 || B > 34)) die ("bad modbuild B %s", arg[0]) ;
          if (nArg > 1 && (!(k = atoi (arg[1])) || k < 1)) die ("bad modbuild k %s", arg[1]) ;
          if (nArg > 2 && !(w = atoi (arg[2]))) die ("bad modbuild w %s", arg[2]) ;
          if (nArg > 3 && !(s = atoi (arg[3]))) die ("bad modbuild w %s", arg[3]) ;
          if (nArg > 4) nArg = 4 ;
          hasher = seqhashCreate (k, w, s) ;
          seqhashReport (hasher, out) ;
          tableBits = B ;
          if (!(gms = modgpuModsetCreate (B, k, w, s))) die ("%s", (char*) modgpuLastError ()) ;
          CHECK (modgpuModsetSetExactOrder (gms, 1)) ;                           /* the reference's index numbering */
        }
      else if (IS ("-r", "--read") && !gms)
        { NEED (1) ;
          if (!(gms = modgpuModsetReadMod (arg[0]))) die ("%s", (char*) modgpuLastError ()) ;
          CHECK (modgpuModsetSetExactOrder (gms, 1)) ;
          const ModgpuHasher *h = modgpuModsetHasher (gms) ;
          hasher = seqhashCreate (h->k, h->w, h->seed) ;
          hasher->factor1 = h->factor1 ; hasher->factor2 = h->factor2 ;          /* the file's own constants */
          tableBits = modgpuModsetBits (gms) ;
          summary (gms) ;
        }
      else if (IS ("-a", "--add")) { NEEDSET () ; NEED (1) ; addFile (arg[0], 0) ; summary (gms) ; }
      else if (IS ("-x", "--add10x")) { NEEDSET () ; NEED (1) ; addFile (arg[0], 1) ; summary (gms) ; }
      else if (IS ("-H", "--hist"))
        { NEEDSET () ; NEED (1) ;
          FILE *f = fopen (arg[0], "w") ;
          if (!f) die ("failed to open histogram file %s", arg[0]) ;
          uint32_t *bins = (uint32_t*) calloc (65536, sizeof (uint32_t)) ;
          CHECK (modgpuModsetHistogram (gms, bins)) ;
          for (int d = 0 ; d < 65536 ; ++d) if (bins[d]) fprintf (f, "DP\t%u\t%u\n", d, bins[d]) ;
          free (bins) ; fclose (f) ;
        }
      else if (IS ("-s", "--setcopy"))
        { NEEDSET () ; NEED (3) ; uint32_t t[4] ;
          CHECK (modgpuModsetSetCopy (gms, atoi (arg[0]), atoi (arg[1]), atoi (arg[2]), t)) ; summary (gms) ;
        }
      else if (IS ("-sM", "--setcopyM"))
        { NEEDSET () ; NEED (1) ; uint32_t t[4] ; CHECK (modgpuModsetSetCopyM (gms, atoi (arg[0]), t)) ; summary (gms) ; }
      else if (IS ("-p", "--prune"))
        { NEEDSET () ; NEED (2) ; CHECK (modgpuModsetPrune (gms, atoi (arg[0]), atoi (arg[1]))) ; summary (gms) ; }
      else if (IS ("-m", "--merge"))
        { NEEDSET () ; NEED (1) ;
          ModgpuModset *other = modgpuModsetReadMod (arg[0]) ;
          if (!other) die ("%s", (char*) modgpuLastError ()) ;
          summary (other) ;
          int merged = modgpuModsetMerge (gms, other) ;                          /* 1 merged, 0 hashers differ, < 0 error */
          if (merged < 0) die ("%s", (char*) modgpuLastError ()) ;
          if (!merged) fprintf (stderr, "modset %s incompatible with current - unable to merge\n", arg[0]) ;
          modgpuModsetDestroy (other) ;
          summary (gms) ;
        }
      else if (IS ("-w", "--write")) { NEEDSET () ; NEED (1) ; CHECK (modgpuModsetWriteMod (gms, arg[0], 1)) ; }
      else if (IS ("-wt", "--writetext"))
        { NEEDSET () ; NEED (1) ;
          FILE *f = fopen (arg[0], "w") ;
          if (!f) die ("failed to open text file %s", arg[0]) ;
          U64 *value ; U16 *depth ; U8 *info ;
          U32 max = exportSet (&value, &depth, &info) ;
          fprintf (f, "modset bits %d size %d k %d w %d seed %d\n", tableBits, max + 1, hasher->k, hasher->w, hasher->seed) ;
          for (U32 i = 1 ; i <= max ; ++i)
            fprintf (f, "%d\t%s\t%d\t%d\n", i, seqString (value[i], hasher->k), depth[i], info[i]) ;
          free (value) ; free (depth) ; free (info) ; fclose (f) ;
        }
      else if (IS ("-P", "--refpaint")) { NEEDSET () ; NEED (1) ; refPaint (arg[0]) ; }
      else die ("unknown command %s", cmd) ;
      a += 1 + nArg ;
      timeUpdate (out) ;
    }
  fprintf (out, "total resources used: ") ; timeTotal (out) ;
  if (out != stdout) { printf ("total resources used: ") ; timeTotal (stdout) ; }
  if (gms) modgpuModsetDestroy (gms) ;
  return 0 ;
}
