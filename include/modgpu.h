/* modgpu.h - C ABI of libmodgpu.so: modimizer's data-parallel hot path on B200.
 *
 * Everything hot is hand-written CUDA for sm_100a behind these entry points;
 * there is no CPU fallback: every call fails loudly (non-zero return +
 * modgpuLastError()) when no CUDA device is usable.  Plain pointers and sizes
 * only - no torch types, no C++ in the signatures.
 *
 * The reference (richarddurbin/modimizer) has no FFI: its "operator API" is
 * the C headers seqhash.h / modset.h plus three caller loops.  Each entry
 * point cites the reference interface it replaces (file:line in the reference
 * tree).  INTEGRATION.md shows the patch a maintainer applies to modutils.c /
 * modmap.c to bind them.
 *
 * Conventions
 *   - sequences: concatenated bytes + nSeq+1 offsets (offs[0] = 0).  Bytes are
 *     either reference codes 0..3 (what seqIOread hands to addSequence after
 *     dna2indexConv, seqio.c:643-652, with N->0 as patched at modutils.c:39) or
 *     raw ASCII ACGTN/acgtn (isAscii != 0, same mapping, N/n -> a).  Code bytes
 *     must be 0..3 (the reference never produces anything else: seqIOread dies on
 *     characters dna2indexConv does not map); other values give undefined k-mers.
 *   - return value: 0 on success, a negative MODGPU_E* code otherwise; functions
 *     returning counts use UINT64_MAX for failure.
 *   - "d_" parameters are device pointers; `stream` is a cudaStream_t passed as
 *     void* (0 = the legacy default stream).
 */
#ifndef MODGPU_H
#define MODGPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MODGPU_OK            0
#define MODGPU_ENODEVICE    -1   /* no usable CUDA device / driver */
#define MODGPU_ECUDA        -2   /* a CUDA runtime call or kernel failed */
#define MODGPU_EINVAL       -3   /* bad argument (k, w, bits, sizes ...) */
#define MODGPU_EFULL        -4   /* table or output list overflow (reference: die(), modset.c:58) */
#define MODGPU_ENOMEM       -5
#define MODGPU_ESKEW        -6   /* sharded add: a group of batches was too skewed for its buckets and was skipped on every rank */

const char *modgpuLastError(void);
int modgpuDeviceCount(void);
int modgpuSetDevice(int dev);
const char *modgpuVersion(void);

/* ------------------------------------------------------------------ hasher
 * replaces: Seqhash + seqhashCreate (seqhash.h:15-23, seqhash.c:20-37).
 * factor1 is taken from libc srandom()/random() exactly as the reference does,
 * on the host, never recomputed on the device. */
typedef struct ModgpuHasher {
  int32_t k;            /* 1..31                                    seqhash.c:24 */
  int32_t w;            /* the modimizer modulus d, >= 1            seqhash.c:25 */
  int32_t seed;
  int32_t shift1;       /* 64 - 2k                                  seqhash.c:32 */
  uint64_t mask;        /* 2k low bits                              seqhash.c:27 */
  uint64_t factor1;     /* odd multiplier                           seqhash.c:31 */
  uint64_t factor2;     /* set, unused on the path                  seqhash.c:33 */
} ModgpuHasher;

int modgpuHasherInit(ModgpuHasher *h, int k, int w, int seed);
/* adopt the constants of an existing reference Seqhash (80-byte POD, seqhash.h:15-23) */
int modgpuHasherFromSeqhash(ModgpuHasher *h, const void *seqhash);
/* seqhash() (seqhash.h:58) on the host, for callers that need one value */
uint64_t modgpuHash(const ModgpuHasher *h, uint64_t kmer);

/* ------------------------------------------------- device-level kernels --
 * K1  pack2bit + read-end flags    replaces the per-byte conversion loop of
 *                                  seqIOread (seqio.c:322,328-331)
 * K2  hash_select                  replaces modRCiterator/modRCnext
 *                                  (seqhash.c:60-79,154-196)
 * Layouts: packed = 32 bases per uint64, first base in the top two bits;
 *          ends   = 1 bit per base (bit g&31 of word g>>5), set on the last
 *                   base of every non-empty sequence.
 * Buffers must hold modgpuPackedWords()/modgpuEndsWords() elements. */
uint64_t modgpuPackedWords(uint64_t nBases);
uint64_t modgpuEndsWords(uint64_t nBases);

int modgpuPack2bit(const uint8_t *d_bases, uint64_t nBases, int isAscii,
                   uint64_t *d_packed, void *stream);
int modgpuMarkEnds(const uint64_t *d_offs, uint64_t nSeq, uint64_t nBases,
                   uint32_t *d_ends, void *stream);

/* flags for modgpuHashSelect */
#define MODGPU_SEL_ORDERED   1   /* output list in (sequence, position) order */
#define MODGPU_SEL_STRAND    2   /* bit 63 of each k-mer = isForward (seqhash.c:184) */
#define MODGPU_SEL_NOTMA     4   /* plain global loads instead of TMA bulk staging */
#define MODGPU_SEL_NOPREFILTER 8 /* force the generic per-position evaluation */
#define MODGPU_SEL_NOFUSE 16     /* modset add: keep K2 and the bucket scatter of K3 as separate kernels */
#define MODGPU_SEL_NOFUSEPACK 32 /* modset add: keep K1 as a separate kernel (packed stream through HBM) */
#define MODGPU_SEL_NOLUT 64      /* count mode: arithmetic prefilter instead of the shared-memory candidate table */
#define MODGPU_SEL_GEN1 0x10000  /* count mode: the first-generation kernel (hash_count_kernel) for A/B runs */
#define MODGPU_SEL_APPEND 128    /* modgpuModsetSelectBuckets*: keep the fill counts, append to the buckets of the previous
                                    batches (multi-GPU deferred build: several batches share one peer build) */

/* workspace bytes modgpuHashSelect needs for nBases (look-back descriptors) */
uint64_t modgpuHashSelectWorkspace(uint64_t nBases);

/* Selected modimizers of a packed batch.  d_kmers/d_gpos receive at most cap
 * entries (d_gpos may be NULL; gpos = global base offset of the k-mer start
 * within the batch, < 2^32); *d_count (device, uint64) receives the total
 * number selected, which may exceed cap (-> caller retries with a larger cap). */
int modgpuHashSelect(const ModgpuHasher *h, const uint64_t *d_packed, const uint32_t *d_ends,
                     uint64_t nBases, uint64_t *d_kmers, uint32_t *d_gpos, uint64_t cap,
                     uint64_t *d_count, void *d_workspace, int flags, void *stream);

/* global offset -> (sequence id, position in sequence) for a selected list
 * (the `id`/`offset` columns of modmap.c:113-116) */
int modgpuLocate(const uint32_t *d_gpos, uint64_t n, const uint64_t *d_offs, uint64_t nSeq,
                 uint32_t *d_id, uint32_t *d_pos, void *stream);

/* ---------------------------------------------------------- device table --
 * K3..K6: open-addressing modset in HBM.  replaces Modset + modsetIndexFind
 * (modset.h:17-28, modset.c:45-62) and the count increment of modutils.c:26.
 * One 16-byte slot {kmer, count, index<<2|copy} per entry, linear probing on
 * a Fibonacci hash of the k-mer (NOT the reference's `hash & mask`, whose low
 * bits are constant for selected k-mers), 2^(bits-1) slots for the reference's
 * `bits` so the reference's own capacity rule max < 2^(bits-2) (modset.c:24-26)
 * keeps the load factor <= 0.5. */
typedef struct ModgpuTable ModgpuTable;

ModgpuTable *modgpuTableCreate(int bits, void *stream);
void modgpuTableDestroy(ModgpuTable *t);
int modgpuTableClear(ModgpuTable *t, void *stream);
uint64_t modgpuTableSlots(const ModgpuTable *t);
/* device pointer of the slot array (for fused / peer kernels) */
void *modgpuTableDevicePtr(const ModgpuTable *t);

/* find-or-insert + count (modsetIndexFind(..,true) + ++depth, modutils.c:25-26).
 * d_slot (nullable) receives the slot of every k-mer.  With exactOrder != 0 the
 * list must be in input order and new entries are numbered by first occurrence
 * (index = ++max, modset.c:57) when modgpuTableNumber() runs. */
int modgpuTableInsert(ModgpuTable *t, const uint64_t *d_kmers, uint64_t n,
                      uint32_t *d_slot, int exactOrder, void *stream);
/* distinct entries so far (reference ms->max); synchronises the stream */
uint64_t modgpuTableEntries(ModgpuTable *t, void *stream);
/* assign dense indices max+1.. to entries inserted since the last call.
 * With (d_slot,n) of an exactOrder insert: first-occurrence order (reference
 * numbering); otherwise slot order.  d_index (nullable) receives the index of
 * every list element (ref->index[n] = index, modmap.c:112). */
int modgpuTableNumber(ModgpuTable *t, const uint32_t *d_slot, uint64_t n,
                      uint32_t *d_index, void *stream);
/* lookup only (modsetIndexFind(..,false), modmap.c:202): per k-mer
 * index<<2 | copy, or 0 when absent */
int modgpuTableLookup(const ModgpuTable *t, const uint64_t *d_kmers, uint64_t n,
                      uint32_t *d_out, void *stream);
/* depth histogram over entries, 65536 bins of depth clamped to 65535
 * (depthHistogram, modutils.c:53-63; saturation modutils.c:26) */
int modgpuTableHistogram(const ModgpuTable *t, uint32_t *d_bins65536, void *stream);
/* copy classes from depth thresholds (-s: modutils.c:205-214; -sM: 215-219 when
 * c1 = c2 = -1) or, with exact != 0, from exact multiplicity 1/2/else
 * (modmap.c:125-129).  d_classCounts (nullable, 4 x uint32) receives tallies. */
int modgpuTableClassify(ModgpuTable *t, int c1, int c2, int cM, int exact,
                        uint32_t *d_classCounts, void *stream);
/* dense export in index order 1..max: value[i-1], depth (clamped), info&3
 * (modsetPack / the -wt dump, modset.c:36-43, modutils.c:191-200) */
int modgpuTableExport(ModgpuTable *t, uint64_t *d_value, uint16_t *d_depth, uint8_t *d_info,
                      uint32_t *d_count32 /* nullable: unclamped counts */, void *stream);
/* load a host/reference modset (value/depth/info, 1..max) into the table */
int modgpuTableImport(ModgpuTable *t, const uint64_t *d_value, const uint16_t *d_depth,
                      const uint8_t *d_info, uint64_t n, void *stream);

/* ------------------------------------------- host-level batched modset --
 * The three caller loops of the reference, batched, with HOST buffers:
 *   modgpuModsetAdd      == addSequence over a batch        modutils.c:19-31
 *   modgpuReferenceBuild == referenceFastaRead's loop+pack  modmap.c:93-134,74-91
 *   modgpuReferenceQuery == queryProcess's seed loop        modmap.c:196-231
 * Host<->device copies happen inside these calls (pinned staging, chunked,
 * overlapped with the kernels). */
typedef struct ModgpuModset ModgpuModset;

ModgpuModset *modgpuModsetCreate(int bits, int k, int w, int seed);   /* modsetCreate, modset.c:15-31 */
void modgpuModsetDestroy(ModgpuModset *ms);
const ModgpuHasher *modgpuModsetHasher(const ModgpuModset *ms);
int modgpuModsetBits(const ModgpuModset *ms);                          /* ms->tableBits */
/* the CUDA device the set lives on: a thread other than its creator calls modgpuSetDevice(this) before using it */
int modgpuModsetDevice(const ModgpuModset *ms);
ModgpuTable *modgpuModsetTable(ModgpuModset *ms);
/* use the caller's stream (e.g. torch's current stream) for all work */
int modgpuModsetSetStream(ModgpuModset *ms, void *stream);
/* MODGPU_SEL_* flags forwarded to the hash/select kernel (tuning, A/B tests);
 * bits 8..15 = insert-locality override + 1: 0 auto, 1 off, v = 2^(v-1) table regions; bits 16.. = more MODGPU_SEL_* flags */
int modgpuModsetSetFlags(ModgpuModset *ms, int flags);
/* exactOrder != 0: keep the reference's first-occurrence index numbering */
int modgpuModsetSetExactOrder(ModgpuModset *ms, int exactOrder);

/* returns the number of hashes added (reference "total hashes"), UINT64_MAX on error */
uint64_t modgpuModsetAdd(ModgpuModset *ms, const char *bases, const uint64_t *offs,
                         uint64_t nSeq, int isAscii);
/* same with the batch already resident in device memory */
uint64_t modgpuModsetAddDevice(ModgpuModset *ms, const uint8_t *d_bases, const uint64_t *d_offs,
                               uint64_t nSeq, uint64_t nBases, int isAscii);
/* same with the sequences in the reference's own 2-bit packing (sqioSeqPack, seqio.c:557-570, the payload of its
 * "binary" seqio files): four bases per byte, first base in the top two bits, a last byte with fewer than four bases
 * holding them in its low bits.  Sequence r has offs[r+1] - offs[r] bases and its (len+3)/4 bytes start at
 * packed + byteOffs[r] (byteOffs has nSeq+1 entries, non-decreasing).  0.25 bytes per base cross PCIe. */
uint64_t modgpuModsetAddPacked(ModgpuModset *ms, const uint8_t *packed, const uint64_t *byteOffs,
                               const uint64_t *offs, uint64_t nSeq);
uint32_t modgpuModsetMax(ModgpuModset *ms);                            /* ms->max */
/* host arrays of length max, index order (sync-to-host of value/depth/info) */
int modgpuModsetExport(ModgpuModset *ms, uint64_t *value, uint16_t *depth, uint8_t *info);
int modgpuModsetHistogram(ModgpuModset *ms, uint32_t *bins65536);      /* modutils.c:53-63 */
int modgpuModsetSetCopy(ModgpuModset *ms, int c1, int c2, int cM, uint32_t classCounts[4]);
int modgpuModsetSetCopyM(ModgpuModset *ms, int cM, uint32_t classCounts[4]);
/* batched modsetIndexFind(..,false): out[i] = index or 0 */
int modgpuModsetFind(ModgpuModset *ms, const uint64_t *kmers, uint64_t n, uint32_t *index, uint8_t *copy);
/* modsetSummary text (modset.c:130-153); returns bytes written */
int modgpuModsetSummary(ModgpuModset *ms, char *buf, int n);
/* upload a host modset (e.g. from modsetRead): entries 1..max */
int modgpuModsetImport(ModgpuModset *ms, const uint64_t *value, const uint16_t *depth,
                       const uint8_t *info, uint64_t n);

/* K1+K2 only: select the modimizers of a device-resident batch into the
 * object's own device list (no table access).  *d_kmers stays valid until the
 * next call on this object.  For callers that route k-mers themselves, e.g. the
 * hash-sharded multi-GPU build (modimizer_b200/dist.py). */
int modgpuModsetSelectDevice(ModgpuModset *ms, const uint8_t *d_bases, const uint64_t *d_offs,
                             uint64_t nSeq, uint64_t nBases, int isAscii,
                             const uint64_t **d_kmers, uint64_t *nSelected);
/* same with the batch in host memory (one chunk, < 2^32 bases) */
int modgpuModsetSelectHost(ModgpuModset *ms, const char *bases, const uint64_t *offs,
                           uint64_t nSeq, int isAscii, const uint64_t **d_kmers, uint64_t *nSelected);
/* Multi-GPU count mode without host round trips: K1 + K2 with the selected k-mers
 * written into nOwners segments (segCap entries each) of the caller's send buffer,
 * segment o holding the k-mers owned by rank o (modgpuOwnerOf); d_counts[o] (uint32)
 * receives their number and exceeds segCap when the segment overflowed (the
 * caller then repeats the batch through modgpuModsetSelectDevice). */
int modgpuModsetSelectOwnersDevice(ModgpuModset *ms, const uint8_t *d_bases, const uint64_t *d_offs,
                                   uint64_t nSeq, uint64_t nBases, int isAscii, uint32_t nOwners,
                                   uint64_t *d_segments, uint64_t segCap, uint32_t *d_counts);
int modgpuModsetSelectOwnersHost(ModgpuModset *ms, const char *bases, const uint64_t *offs, uint64_t nSeq,
                                 int isAscii, uint32_t nOwners, uint64_t *d_segments, uint64_t segCap,
                                 uint32_t *d_counts);
/* Fully fused multi-GPU build: K2 scatters every selected k-mer into bucket
 * (owner * R + region) of the caller's send buffer (R = modgpuModsetRegions(),
 * bucketCap entries per bucket, fill counts in d_cursors[nOwners * R]); after an
 * equal-split all-to-all of buckets and cursors the owner builds each of its
 * table regions in shared memory straight from the nSrc received buckets.
 * K-mers beyond a bucket's capacity travel in per-owner overflow segments
 * (overflowCap entries, counts in d_ovfCounts[nOwners], which exceed overflowCap
 * when even those overflowed).  *d_count (device) = number selected. */
uint32_t modgpuModsetRegions(ModgpuModset *ms);
int modgpuModsetSelectBucketsDevice(ModgpuModset *ms, const uint8_t *d_bases, const uint64_t *d_offs,
                                    uint64_t nSeq, uint64_t nBases, int isAscii, uint32_t nOwners,
                                    uint64_t *d_buckets, uint32_t bucketCap, uint32_t *d_cursors,
                                    uint64_t *d_overflow, uint64_t overflowCap, uint32_t *d_ovfCounts,
                                    uint64_t *d_count);
int modgpuModsetSelectBucketsHost(ModgpuModset *ms, const char *bases, const uint64_t *offs, uint64_t nSeq,
                                  int isAscii, uint32_t nOwners, uint64_t *d_buckets, uint32_t bucketCap,
                                  uint32_t *d_cursors, uint64_t *d_overflow, uint64_t overflowCap,
                                  uint32_t *d_ovfCounts, uint64_t *d_count);
int modgpuModsetBuildFromBuckets(ModgpuModset *ms, const uint64_t *d_buckets, const uint32_t *d_cursors,
                                 uint32_t bucketCap, uint32_t nSrc, const uint64_t *d_overflow,
                                 uint64_t overflowCap, const uint32_t *d_ovfCounts);
/* insert + count nSegs received segments whose fill counts are in device memory
 * (expectedN sizes the table's region buckets) */
int modgpuModsetInsertSegments(ModgpuModset *ms, const uint64_t *d_segments, uint32_t nSegs, uint64_t segCap,
                               const uint32_t *d_counts, uint64_t expectedN);
/* insert + count a device list of k-mers into the object's table */
int modgpuModsetInsertDevice(ModgpuModset *ms, const uint64_t *d_kmers, uint64_t n);
int modgpuModsetClear(ModgpuModset *ms);
/* Deferred build for streaming many batches into one set (new; the reference's loop has no counterpart: it pays one
 * random probe per k-mer, addSequence modutils.c:19-31).  With nChunks > 1 the selected k-mers of up to nChunks device
 * chunks wait in the table's per-region buckets and the regions are built once for all of them - a populated table is
 * read and written once per nChunks chunks instead of once per chunk.  Results are identical.  What changes is when a
 * full table (modset.c:58, the reference dies) is reported: by modgpuModsetFlush or by the first call that reads the
 * set (every reader flushes first), not by the modgpuModsetAdd* call that overfilled it.  nChunks <= 1: default. */
int modgpuModsetSetAccumulate(ModgpuModset *ms, int nChunks);
/* apply what is waiting; MODGPU_EFULL when the table is over its capacity */
int modgpuModsetFlush(ModgpuModset *ms);

/* ------------------------------------------------ whole-set operations --
 * modsetDepthPrune (modset.c:64-77): keep entries with min <= depth < max
 * (max == 0: no upper bound), renumbered in their old index order. */
int modgpuModsetPrune(ModgpuModset *ms, int min, int max);
/* modsetMerge (modset.c:106-128): union of b into a; depths add (clamp 65535), copy
 * numbers add (clamp 3), new entries numbered in b's index order.  Returns 1 when
 * merged, 0 when the hashers differ (reference returns false), < 0 on error. */
int modgpuModsetMerge(ModgpuModset *a, ModgpuModset *b);
/* modsetCreate for an existing hasher (e.g. read from a file) */
ModgpuModset *modgpuModsetCreateWithHasher(int bits, const ModgpuHasher *h);
/* modsetWrite / modsetRead (modset.c:79-104): the reference's "MSHSTv2" file, readable by an
 * unmodified `modutils -r`; gzip != 0 compresses like fzopen(path,"w") (utils.c:108-127) */
int modgpuModsetWriteMod(ModgpuModset *ms, const char *path, int gzip);
ModgpuModset *modgpuModsetReadMod(const char *path);
/* hot loop of modasm's readsetFileRead (modasm.c:151-191): per read the hits
 * (index | 0x80000000 when forward) and dx (U16 distance to the previous hit) of
 * the modimizers found in the set, the misses per read, and depth re-counted from
 * these reads (resetDepth != 0 zeroes it first).  Returns the hit total. */
uint64_t modgpuModsetReadset(ModgpuModset *ms, const char *bases, const uint64_t *offs, uint64_t nSeq, int isAscii,
                             int resetDepth, uint64_t *hitOff, uint32_t *hit, uint16_t *dx, int32_t *nMiss, uint64_t cap);

/* ------------------------------------------- host Modset <-> device twin --
 * The reference's callers use the public fields of Modset directly (ms->depth[index] modutils.c:26, ms->value[i] /
 * ms->max modutils.c:57-69, ms->info through modset.h:53-69), so a drop-in (libmodshim.so, include/modshim.h) has to
 * keep host arrays and device table in step.  These three calls are what it needs beside Export / Import:
 *
 * batched modsetIndexFind (modset.c:45-62) on host arrays: index[i] = the index of kmers[i], 0 when absent; with
 * isAdd != 0 absent k-mers are inserted and numbered ++max in input order (the reference's numbering); depths are
 * not touched (the reference's caller increments them itself) */
int modgpuModsetIndexFindBatch(ModgpuModset *ms, const uint64_t *kmers, uint64_t n, int isAdd, uint32_t *index);
/* the caller changed depth[] / info[] of entries 1..n on the host (++depth, msSetCopy*, msSetMinor ...): upload them.
 * Either pointer may be NULL.  info is the whole byte (modset.h:44-52). */
int modgpuModsetSetDepthInfo(ModgpuModset *ms, const uint16_t *depth, const uint8_t *info, uint64_t n);
/* the reference's own index[] table (2^bits U32; home slot hash & mask, odd double-hashing stride, entries
 * inserted in index order: modset.c:48-57), built on the device and copied to `index` (host) */
int modgpuModsetReferenceIndex(ModgpuModset *ms, uint32_t *index);

/* per-kernel accumulated device time (ms) since the last reset, measured with
 * CUDA events on the object's stream when profiling is enabled */
#define MODGPU_T_PACK 0
#define MODGPU_T_SELECT 1
#define MODGPU_T_INSERT 2
#define MODGPU_T_OTHER 3
#define MODGPU_T_N 4
int modgpuModsetProfile(ModgpuModset *ms, int enable);
int modgpuModsetTimes(ModgpuModset *ms, double ms_out[MODGPU_T_N], uint64_t launches_out[MODGPU_T_N]);

/* ------------------------------------------------------ modmap reference */
typedef struct ModgpuReference ModgpuReference;

/* counts = { nHashes, nCopy1, nCopy2, nMulti } (the two lines of modmap.c:123-130) */
ModgpuReference *modgpuReferenceBuild(int bits, int k, int w, int seed, const char *bases,
                                      const uint64_t *offs, uint64_t nSeq, int isAscii,
                                      uint32_t counts[4]);
void modgpuReferenceDestroy(ModgpuReference *r);
ModgpuModset *modgpuReferenceModset(ModgpuReference *r);
uint32_t modgpuReferenceMax(ModgpuReference *r);
/* host copies of the Reference arrays (modmap.c:35-47): index/offset/id/rev
 * have max entries, depth/loc have modset max+1 */
int modgpuReferenceExport(ModgpuReference *r, uint32_t *index, uint32_t *offset, uint32_t *id,
                          uint32_t *depth, uint32_t *rev, uint32_t *loc);
/* seed loop of queryProcess: same outputs as the oracle's ref_query
 * (oracle/harness_api.h).  Returns the seed total (UINT64_MAX on error). */
uint64_t modgpuReferenceQuery(ModgpuReference *r, const char *bases, const uint64_t *offs,
                              uint64_t nSeq, int isAscii, uint64_t *seedOff,
                              uint32_t *seedIndex, uint32_t *seedPos,
                              uint32_t *hitId, uint32_t *hitOffset,
                              int32_t *counters, uint64_t cap);

/* -------------------------------------------------------------- scanner --
 * modRCiterator / modRCnext (seqhash.c:154-196) for a batch of sequences with HOST buffers on both sides:
 * every modimizer in (sequence, position) order - k-mer with bit 63 = isForward (seqhash.c:184), its start
 * position inside its sequence, and seqOff[r] .. seqOff[r+1] = the results of sequence r (nSeq+1 entries).
 * Returns the total (results beyond cap are not stored), UINT64_MAX on error.  include/modshim.h wraps it
 * under the reference's own names. */
typedef struct ModgpuScanner ModgpuScanner;
ModgpuScanner *modgpuScannerCreate(const ModgpuHasher *h);
void modgpuScannerDestroy(ModgpuScanner *sc);
uint64_t modgpuScannerScan(ModgpuScanner *sc, const char *bases, const uint64_t *offs, uint64_t nSeq, int isAscii,
                           uint64_t *kmers, uint32_t *pos, uint64_t *seqOff, uint64_t cap);

/* ------------------------------------------------ multi-GPU partitioning --
 * owner of a k-mer in a table sharded over nOwners GPUs (independent of the
 * slot hash); counts per owner; scatter into per-owner contiguous segments
 * (d_cursors = exclusive prefix of the counts on entry). */
uint32_t modgpuOwnerOf(uint64_t kmer, uint32_t nOwners);
int modgpuOwnerCount(const uint64_t *d_kmers, uint64_t n, uint32_t nOwners, uint64_t *d_counts, void *stream);
int modgpuOwnerScatter(const uint64_t *d_kmers, uint64_t n, uint32_t nOwners, uint64_t *d_cursors,
                       uint64_t *d_out, void *stream);

/* ------------------------------------------------- peer memory (NVLink) --
 * The multi-GPU exchange without a copy: every rank scatters its selected k-mers
 * into per-(owner, region) buckets in ITS OWN memory (modgpuModsetSelectBuckets*),
 * and the owner's region build reads the buckets of all ranks directly through
 * peer-mapped pointers (modgpuModsetBuildFromPeers) - the transfer happens inside
 * the build kernel, bucket by bucket, and only filled entries cross NVLink.
 * One process per GPU: buffers are shared through CUDA IPC handles (64 bytes). */
#define MODGPU_PEER_HANDLE_BYTES 64
#define MODGPU_MAX_PEERS 16
void *modgpuPeerAlloc(size_t bytes);
void modgpuPeerFree(void *d_ptr);
int modgpuPeerExport(void *d_ptr, void *handle64);
void *modgpuPeerOpen(const void *handle64);
int modgpuPeerClose(void *d_ptr);
/* d_buckets[s] / d_overflow[s]: source rank s's bucket array / overflow segment FOR THIS OWNER (peer-mapped,
 * nRegions x bucketCap / overflowCap k-mers); d_cursors, d_ovfCounts: LOCAL copies of the fill counts,
 * [nSrc][nRegions] and [nSrc] (the count exchange doubles as the cross-GPU barrier). */
int modgpuModsetBuildFromPeers(ModgpuModset *ms, const uint64_t *const *d_buckets, const uint32_t *d_cursors,
                               uint32_t bucketCap, uint32_t nSrc, const uint64_t *const *d_overflow,
                               uint64_t overflowCap, const uint32_t *d_ovfCounts);

/* -------------------------------------------- the sharded modset, in C --
 * One process (or thread) per GPU; the table is sharded by modgpuOwnerOf, every rank feeds ITS OWN chunk of the input.
 * The reference has no distributed mode (its recipe is modsetMerge, modset.c:106-128, modutils.c:101-103); counts are
 * commutative sums, so the union of the shards equals the single-GPU modset for any number of GPUs.
 * Per group of batches: select into per-(owner, region) buckets in the selecting rank's memory, ONE equal-split
 * all-to-all of the fill-count rows (it doubles as the cross-GPU barrier), region build on the owner reading the
 * peers' buckets over NVLink.  A group is applied on every rank or on none (MODGPU_ESKEW from Synchronize).
 * The communicator is three callbacks; modgpuCommFromNccl binds them to an ncclComm_t (libnccl loaded at run time). */
typedef struct ModgpuComm {
  void *ctx;
  int rank, world;
  /* equal-split all-to-all of DEVICE memory, ordered on `stream`: bytesPerPeer bytes to and from every rank (self included) */
  int (*alltoall)(void *ctx, const void *d_send, void *d_recv, size_t bytesPerPeer, void *stream);
  /* HOST all-gather of `bytes` per rank, blocking (set-up only: sizes and the 64-byte IPC handles) */
  int (*allgather)(void *ctx, const void *in, void *out, size_t bytes);
  /* HOST barrier, blocking (before peer buffers are unmapped) */
  int (*barrier)(void *ctx);
} ModgpuComm;
int modgpuCommFromNccl(ModgpuComm *out, void *ncclComm /* ncclComm_t */, int rank, int world);
void modgpuCommNcclRelease(ModgpuComm *c);

typedef struct ModgpuSharded ModgpuSharded;
/* bits = the PER-GPU table bits (capacity 2^(bits-2) entries per GPU).  Calls marked COLLECTIVE must be made by every rank. */
ModgpuSharded *modgpuShardedCreate(int bits, int k, int w, int seed, const ModgpuComm *comm);
void modgpuShardedDestroy(ModgpuSharded *s);                                   /* COLLECTIVE */
ModgpuModset *modgpuShardedLocal(ModgpuSharded *s);                            /* this rank's shard: Export, Histogram, Max ... */
int modgpuShardedSetStream(ModgpuSharded *s, void *stream);
/* COLLECTIVE: size and map the peer buckets for batches of up to maxBasesPerBatch bases per rank (implied by the first add) */
int modgpuShardedReserve(ModgpuSharded *s, uint64_t maxBasesPerBatch);
/* COLLECTIVE: up to nBatches batches share one count exchange and one peer build (streaming many batches into one set) */
int modgpuShardedSetAccumulate(ModgpuSharded *s, int nBatches);
/* COLLECTIVE: size the overflow segments for the worst case (every k-mer of a group to one owner): no group is skipped */
int modgpuShardedSetRobust(ModgpuSharded *s, int on);
/* this rank's batch (host / device memory, < 2^32 bases); COLLECTIVE once per group (every rank adds the same number of batches) */
int modgpuShardedAdd(ModgpuSharded *s, const char *bases, const uint64_t *offs, uint64_t nSeq, int isAscii);
int modgpuShardedAddDevice(ModgpuSharded *s, const uint8_t *d_bases, const uint64_t *d_offs, uint64_t nSeq,
                           uint64_t nBases, int isAscii);
int modgpuShardedFlush(ModgpuSharded *s);                                      /* COLLECTIVE: apply the batches that are waiting */
/* COLLECTIVE: flush and wait; *nSelected = k-mers this rank selected since the last call */
int modgpuShardedSynchronize(ModgpuSharded *s, uint64_t *nSelected);
int modgpuShardedClear(ModgpuSharded *s);

/* ---------------------------------------------------- pinned host memory --
 * "seqio parsing stays on the host, feeding pinned buffers" */
void *modgpuHostAlloc(size_t bytes);
void modgpuHostFree(void *p);

/* --------------------------------------------- synthetic inputs on device --
 * include/modgpu_synth.h generators, for benchmarks and parity tests */
int modgpuSynthGenome(uint64_t seed, uint64_t start, uint64_t n, int dupMode,
                      uint8_t *d_codes, void *stream);
int modgpuSynthReads(const void *spec /* MgReadSpec */, uint64_t firstRead, uint64_t nReads,
                     int ont, uint8_t *d_codes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MODGPU_H */
