/* modshim.h - libmodshim.so: the reference's OWN seqhash symbols, served by libmodgpu (B200).
 *
 * The shim is compiled against the reference's headers where they lie (modimizer_b200/csrc/shim/Makefile,
 * -I/root/reference), so the structs are the reference's by construction.  An unmodified caller that links
 * libmodshim instead of seqhash.o gets, under the unchanged names and signatures of seqhash.h:
 *
 *   Seqhash *seqhashCreate (int k, int w, int seed)                         seqhash.c:20-37
 *   void seqhashWrite (Seqhash*, FILE*) ; Seqhash *seqhashRead (FILE*)      seqhash.c:41-53
 *   void seqhashReport (Seqhash*, FILE*)                                    seqhash.c:55-56
 *   SeqhashRCiterator *modRCiterator (Seqhash*, char *s, int len)           seqhash.c:154-177   K1 + K2 on the GPU
 *   bool modRCnext (SeqhashRCiterator*, U64 *kmer, int *pos, bool *isF)     seqhash.c:179-196   pops the result list
 *   char *seqString (U64 kmer, int len)                                     seqhash.c:198-206
 *   minimizerRCiterator / minimizerRCnext                                   seqhash.c:83-152    die(): not on the GPU path
 *
 * seqhashRCiteratorDestroy, seqhashDestroy, seqhash() and seqhashString() are inline in the reference's header and
 * keep working on the shim's objects (the iterator's result arrays live in its hashBuf / fBuf members).
 * Errors follow the reference: die() (utils.c:19-30).
 *
 * The modset half of the reference API (modsetIndexFind + ++ms->depth[index], one k-mer per call on host arrays)
 * cannot be accelerated call by call; its callers switch to the batched entry points of modgpu.h instead
 * (INTEGRATION.md).  modimizer_b200/csrc/shim/modutils_gpu.c is a complete C host doing exactly that with the
 * reference's seqio, and tests/test_gpu_cli.py holds it byte-identical to the stock modutils.
 *
 * This header only adds the one extra symbol the shim exports; include the reference's seqhash.h for the rest.
 */
#ifndef MODSHIM_H
#define MODSHIM_H
#include "modgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* the scanner (modgpu.h) the iterator uses for this hasher: callers that want whole batches instead of one
 * sequence per modRCiterator call share it.  `seqhash` is a reference Seqhash*. */
ModgpuScanner *modshimScanner(void *seqhash);

#ifdef __cplusplus
}
#endif
#endif
