/* modshim.h - libmodshim.so: the reference's OWN seqhash AND modset symbols, served by libmodgpu (B200).
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
 * The modset half (modset.h:30-42) is exported under its own names too (modshim_modset.c): a host Modset with the
 * reference's layout whose arrays are valid whenever control is in caller code, backed by a device twin.  The
 * UNMODIFIED modutils.c / modmap.c link against libmodshim.so and produce the stock tools' bytes
 * (tests/test_gpu_cli.py: modutils_shim, modmap_shim); modsetIndexFind is then one device lookup per call - exact,
 * not fast.  The throughput path is the batched one: the two caller loops patched as INTEGRATION.md shows
 * (csrc/shim/modutils_hot.c, modmap_hot.c, applied to the reference's sources at build time: modutils_dropin,
 * modmap_dropin).
 *
 * This header declares what the shim adds; include the reference's seqhash.h / modset.h for the rest.
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

/* ---- modset.h under its own names (modshim_modset.c): modsetCreate, modsetDestroy, modsetWrite, modsetRead,
 * modsetIndexFind, modsetSummary, modsetPack, modsetDepthPrune, modsetMerge (modset.h:30-42) operate on the
 * reference's host Modset, every one of which has a device twin; include the reference's modset.h for them.
 * The additions below are the batched path a caller switches its per-read loop to (`ms` is a reference Modset*): */
ModgpuModset *modshimTwin(void *ms);                                  /* the device twin, for modgpu.h calls */
void modshimSync(void *ms);                                           /* device -> ms->value/depth/info[1..max], ms->max */
/* == the addSequence loop (modutils.c:19-31): copy one sequence (codes 0..3) into the pinned batch of `ms`; full
 * batches are added on the GPU as they fill */
void modshimBatchPut(void *ms, const char *s, long long len);
/* add what is waiting, sync the host arrays, return the hashes added since the last flush (modutils.c: totHash) */
unsigned long long modshimBatchFlush(void *ms);

#ifdef __cplusplus
}
#endif
#endif
