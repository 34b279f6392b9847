/* modutils_hot.c - the ONE function of the reference's modutils.c that changes for the GPU path.
 *
 * Not compiled on its own: csrc/shim/Makefile takes the reference's modutils.c where it lies, renames its
 * addSequenceFile (modutils.c:33-51) out of the way with sed, appends this file and links the result
 * (modutils_dropin) against libmodshim.so - the patch of INTEGRATION.md section 1 applied at build time, nothing of
 * the reference copied into this repo.  Everything else in the tool - the command interpreter, -w -r -wt -rt -p -s
 * -sM -m -H -d -P, the printers - is the reference's own code running on libmodshim's modset.h / seqhash.h symbols.
 *
 * The per-read loop addSequence (modutils.c:19-31: modRCiterator, modsetIndexFind, ++depth) becomes: copy each read
 * into the pinned batch of the Modset (seqio reuses its buffer, seqio.h:46-48) and let the GPU add whole batches.
 */
#include "modshim.h"

static bool addSequenceFile (Modset *ms, char *filename, bool is10x)
{
  U64 nSeq = 0, totLen = 0, totHash = 0 ;

  dna2indexConv['N'] = dna2indexConv['n'] = 0 ;                       /* modutils.c:39 */
  SeqIO *si = seqIOopenRead (filename, dna2indexConv, false) ;
  if (!si) return false ;
  while (seqIOread (si))
    { ++nSeq ; totLen += si->seqLen ;
      if (is10x && (nSeq & 0x1)) modshimBatchPut (ms, sqioSeq(si)+23, (long long) si->seqLen - 23) ;   /* modutils.c:44 */
      else modshimBatchPut (ms, sqioSeq(si), (long long) si->seqLen) ;
    }
  seqIOclose (si) ;
  totHash = modshimBatchFlush (ms) ;                                  /* the host arrays and ms->max are current again */
  fprintf (outFile, "added %llu sequences total length %llu total hashes %llu, new max %u\n",
	   nSeq, totLen, totHash, ms->max) ;
  return true ;
}
