/* modgpu_stub.c - TEST INFRASTRUCTURE: the handful of include/modgpu.h entry points that csrc/shim/modmap_gpu.c calls,
 * answered by the CPU oracle (oracle/liboracle.so), so that the HOST logic of that driver - batching, the .ref
 * layout, the Q / seed lines and the colinear-block pass - is checked against the stock modmap in the CPU test
 * suite (tests/test_oracle_vs_ref.py).  Never linked into the product; the product driver links libmodgpu.so. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#define HX(n) orc_##n
#include "../../oracle/harness_api.h"

typedef struct { HxRef *r ; } Ref ;

const char *modgpuLastError (void) { return "stub" ; }
void *modgpuHostAlloc (size_t n) { return malloc (n) ; }
void modgpuHostFree (void *p) { free (p) ; }
void *modgpuReferenceBuild (int bits, int k, int w, int seed, const char *bases, const uint64_t *offs, uint64_t nSeq,
                            int isAscii, uint32_t counts[4])
{ Ref *x = (Ref*) malloc (sizeof (Ref)) ; x->r = orc_ref_build (bits, k, w, seed, bases, offs, (int64_t) nSeq, counts) ; return x ; }
void modgpuReferenceDestroy (void *x) { orc_ref_free (((Ref*) x)->r) ; free (x) ; }
void *modgpuReferenceModset (void *x) { return orc_ref_modset (((Ref*) x)->r) ; }
uint32_t modgpuReferenceMax (void *x) { return orc_ref_max (((Ref*) x)->r) ; }
uint32_t modgpuModsetMax (void *ms) { return orc_modset_max ((HxModset*) ms) ; }
int modgpuModsetWriteMod (void *ms, const char *path, int gzip) { return 0 ; }   /* device-side writer: GPU test only */
int modgpuReferenceExport (void *x, uint32_t *index, uint32_t *offset, uint32_t *id, uint32_t *depth, uint32_t *rev, uint32_t *loc)
{ orc_ref_export (((Ref*) x)->r, index, offset, id, depth, rev, loc) ; return 0 ; }
uint64_t modgpuReferenceQuery (void *x, const char *bases, const uint64_t *offs, uint64_t nSeq, int isAscii, uint64_t *seedOff,
                               uint32_t *seedIndex, uint32_t *seedPos, uint32_t *hitId, uint32_t *hitOffset, int32_t *counters, uint64_t cap)
{ return (uint64_t) orc_ref_query (((Ref*) x)->r, bases, offs, (int64_t) nSeq, seedOff, seedIndex, seedPos, hitId, hitOffset, counters, (int64_t) cap) ; }
