/* harness_api.h - the one C interface shared by the two CPU checkers:
 *
 *   liboracle.so       (HX = orc_)  oracle/modoracle.c, a from-scratch restatement
 *                                   of the reference algorithm; travels to the GPU box
 *   _ref/libmodref.so  (HX = ref_)  oracle/ref_harness.c linked against the
 *                                   UNMODIFIED reference objects (seqhash.o modset.o ...)
 *                                   compiled in place from /root/reference
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or
 * executed by the product (modimizer_b200/, include/): only tests/,
 * __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of
 * bench.py use it, and only as the checker or the CPU baseline.
 *
 * Sequences are passed as the reference passes them to its hot loop: byte
 * codes 0..3 (reference seqio.c:643-652 after the N->0 patch, modutils.c:39),
 * concatenated, with nseq+1 offsets.
 */
#ifndef HARNESS_API_H
#define HARNESS_API_H

#include <stdint.h>

#ifndef HX
#error "define HX(name) before including harness_api.h"
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct HxModset HxModset;
typedef struct HxRef HxRef;

/* hasher constants as seqhashCreate derives them (seqhash.c:20-37):
   out = { mask, shift1, factor1, factor2 } */
void HX(hasher)(int k, int w, int seed, uint64_t out[4]);

/* modRCiterator/modRCnext over one sequence (seqhash.c:154-196).  Stores the
   first `cap` results, returns the total number produced. */
int64_t HX(mod_scan)(int k, int w, int seed, const char *codes, int len,
                     uint64_t *kmer, int32_t *pos, uint8_t *isF, int64_t cap);

/* modset (modset.c) */
HxModset *HX(modset_new)(int bits, int k, int w, int seed);
void HX(modset_free)(HxModset *ms);
/* touch every page of the index table (no semantic change; for timing bounded samples) */
void HX(modset_prefault)(HxModset *ms);
/* the addSequence loop of modutils.c:19-31 over a batch; returns total hashes */
uint64_t HX(modset_add)(HxModset *ms, const char *codes, const uint64_t *offs, int64_t nseq);
uint32_t HX(modset_max)(HxModset *ms);
/* entries 1..max, in index order, into arrays of length max */
void HX(modset_export)(HxModset *ms, uint64_t *value, uint16_t *depth, uint8_t *info);
uint32_t HX(modset_find)(HxModset *ms, uint64_t kmer);              /* modset.c:45-62, isAdd=0 */
void HX(modset_setcopy)(HxModset *ms, int c1, int c2, int cM);      /* modutils.c:205-214 */
void HX(modset_setcopyM)(HxModset *ms, int cM);                     /* modutils.c:215-219 */
void HX(modset_hist)(HxModset *ms, uint32_t *bins65536);            /* modutils.c:53-63 */
int HX(modset_summary)(HxModset *ms, char *buf, int n);             /* modset.c:130-153 text */
void HX(modset_prune)(HxModset *ms, int min, int max);              /* modset.c:64-77 */
int HX(modset_merge)(HxModset *a, HxModset *b);                     /* modset.c:106-128 */

/* hot loop of modasm's readsetFileRead (modasm.c:151-191): depth is zeroed and
   re-counted from these reads; per read the hits (index | 0x80000000 when the
   k-mer was seen forward) with dx = pos - previous hit's pos (U16), and the
   number of modimizers that miss the set.  Returns the hit total. */
int64_t HX(readset)(HxModset *ms, const char *codes, const uint64_t *offs, int64_t nseq,
                    uint64_t *hitOff, uint32_t *hit, uint16_t *dx, int32_t *nMiss, int64_t cap);

/* modmap reference index (modmap.c:93-134 + referencePack :74-91).
   counts = { nHashes, nCopy1, nCopy2, nMulti } */
HxRef *HX(ref_build)(int bits, int k, int w, int seed, const char *codes,
                     const uint64_t *offs, int64_t nseq, uint32_t counts[4]);
void HX(ref_free)(HxRef *r);
HxModset *HX(ref_modset)(HxRef *r);
uint32_t HX(ref_max)(HxRef *r);
/* index/offset/id/rev have ref_max entries; depth/loc have modset_max+1 */
void HX(ref_export)(HxRef *r, uint32_t *index, uint32_t *offset, uint32_t *id,
                    uint32_t *depth, uint32_t *rev, uint32_t *loc);
/* the seed loop of queryProcess (modmap.c:196-231): per read the seeds
   (index,pos) in order, the Q-line counters {miss,copy1,copy2,multi}, and for
   each seed the reference hit (id,offset) pairs the -v lines print
   (0xFFFFFFFF where the reference prints nothing).  Returns the seed total;
   stores at most cap seeds. */
int64_t HX(ref_query)(HxRef *r, const char *codes, const uint64_t *offs, int64_t nseq,
                      uint64_t *seedOff, uint32_t *seedIndex, uint32_t *seedPos,
                      uint32_t *hitId, uint32_t *hitOffset, /* 2 per seed */
                      int32_t *counters /* 4 per read */, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif
