// sharded.cu - the hash-sharded modset over the GPUs of one node, as a C host (one process or thread per GPU).
//
// The reference has no distributed mode: its recipe is one modset per input and modsetMerge (modset.c:106-128,
// modutils.c:101-103).  Here the reads are sharded by input chunk and the table by an independent hash of the k-mer
// (mg_owner), and the exchange is fused into the kernels on either side of it, over peer memory:
//
//   select   hash_count_kernel<OUT=3> scatters every selected k-mer into the bucket (owner, table region) in the
//            SELECTING rank's own HBM (buffers from modgpuPeerAlloc, mapped into the other ranks by CUDA IPC);
//   counts   ONE equal-split all-to-all per group of batches carries, to every owner, its row of
//            { R bucket fill counts, overflow count, "this sender lost k-mers" flag }; its completion on a rank
//            also means every rank's select has finished: it is the cross-GPU barrier;
//   build    region_build_pipe_kernel<PEER> on the owner builds each 2048-slot region in shared memory reading
//            the G source buckets through the peer-mapped pointers over NVLink.
//
// The group is TRANSACTIONAL: every owner receives every sender's flag, so all ranks take the same decision from
// the same data without another collective - when any overflow segment overflowed (heavily skewed input) no rank
// applies anything of the group, modgpuShardedSynchronize reports it, and the caller repeats the group after
// modgpuShardedSetRobust (overflow segments sized for the worst case).
//
// The only things this file needs from a communicator are in ModgpuComm (include/modgpu.h): an equal-split
// all-to-all of device memory ordered on a stream, and a host all-gather / barrier for set-up and tear-down.
// modgpuCommFromNccl binds them to an ncclComm_t (libnccl is loaded at run time, no link dependency); the Python
// mirror (modimizer_b200/dist.py) binds them to torch.distributed.
#include <dlfcn.h>
#include <math.h>
#include <string.h>
#include <vector>
#include "mg_api.h"

int mg_table_build_from_peers_ex(ModgpuTable *t, const uint64_t *const *d_buckets, const uint32_t *d_cursors, uint64_t cursorStride,
                                 uint32_t cap, uint32_t nSrc, const uint64_t *const *d_overflow, uint64_t overflowCap,
                                 const uint32_t *d_ovfCounts, uint64_t ovfStride, const uint32_t *d_guard, cudaStream_t st);

struct ModgpuSharded {
  ModgpuComm comm;
  ModgpuModset *ms = nullptr;
  int G = 1, rank = 0;
  uint32_t R = 0;                       // table regions per rank
  uint32_t accumulate = 1;              // batches per group (one count exchange + one peer build per group)
  bool robust = false;                  // overflow segments sized for the worst case
  // peer buffers: two sets (a group's buckets are read by the peers while the next group is being selected)
  bool reserved = false;
  uint64_t reservedBases = 0;
  uint32_t cap = 0;                     // k-mers per (owner, region) bucket
  uint64_t ovfCap = 0;                  // k-mers per owner overflow segment
  void *mine[4] = { nullptr, nullptr, nullptr, nullptr };          // sb[0], sb[1], so[0], so[1]
  std::vector<void *> opened;
  const uint64_t *bptr[2][MODGPU_MAX_PEERS], *optr[2][MODGPU_MAX_PEERS];
  DevBuf cursors, ovf, rowsSend, rowsRecv, misc;                     // misc: cnt u64 | selAcc u64 | skipAcc u32 | guard u32
  PinBuf hMisc;
  uint32_t batch = 0, pending = 0;
};

static uint64_t *sh_cnt(ModgpuSharded *s) { return (uint64_t *)s->misc.p; }
static uint64_t *sh_selacc(ModgpuSharded *s) { return (uint64_t *)s->misc.p + 1; }
static uint32_t *sh_skipacc(ModgpuSharded *s) { return (uint32_t *)((uint64_t *)s->misc.p + 2); }
static uint32_t *sh_guard(ModgpuSharded *s) { return (uint32_t *)((uint64_t *)s->misc.p + 2) + 1; }

// rows[o] = { cursors[o][0..R), ovf[o], any overflow segment of THIS sender beyond its capacity }
__global__ void __launch_bounds__(256) rows_pack_kernel(const uint32_t *__restrict__ cursors, const uint32_t *__restrict__ ovf,
                                                        uint32_t G, uint32_t R, uint32_t ovfCap, uint32_t *__restrict__ rows)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, n = (uint64_t)G * (R + 2);
  uint32_t lost = 0;
  for (uint32_t o = 0; o < G; ++o) lost |= (ovf[o] > ovfCap) ? 1u : 0u;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    { const uint32_t o = (uint32_t)(i / (R + 2)), r = (uint32_t)(i % (R + 2));
      rows[i] = r < R ? cursors[(uint64_t)o * R + r] : (r == R ? ovf[o] : lost);
    }
}

// after the exchange: guard = some sender lost k-mers; the group's counters
__global__ void rows_finish_kernel(const uint32_t *__restrict__ rowsRecv, uint32_t G, uint32_t R, uint32_t *guard,
                                   const uint64_t *cnt, uint64_t *selAcc, uint32_t *skipAcc)
{
  uint32_t g = 0;
  for (uint32_t s = 0; s < G; ++s) g |= rowsRecv[(uint64_t)s * (R + 2) + R + 1];
  *guard = g;
  if (g) *skipAcc = 1u; else *selAcc += *cnt;
}

extern "C" ModgpuSharded *modgpuShardedCreate(int bits, int k, int w, int seed, const ModgpuComm *comm)
{
  if (!comm || comm->world < 1 || comm->world > MODGPU_MAX_PEERS || comm->rank < 0 || comm->rank >= comm->world ||
      (comm->world > 1 && (!comm->alltoall || !comm->allgather || !comm->barrier)))
    { mg_set_error("modgpuShardedCreate: bad communicator (world 1..%d, all three callbacks)", MODGPU_MAX_PEERS); return nullptr; }
  ModgpuSharded *s = new ModgpuSharded();
  s->comm = *comm; s->G = comm->world; s->rank = comm->rank;
  s->ms = modgpuModsetCreate(bits, k, w, seed);
  if (!s->ms) { delete s; return nullptr; }
  s->R = modgpuModsetRegions(s->ms);
  if (s->misc.ensure(256) || s->hMisc.ensure(256) ||
      mg_check_cuda(cudaMemsetAsync(s->misc.p, 0, 256, s->ms->stream), "memset", __FILE__, __LINE__))
    { modgpuShardedDestroy(s); return nullptr; }
  return s;
}

static void release_peers(ModgpuSharded *s)
{
  if (!s->reserved) return;
  cudaStreamSynchronize(s->ms->stream);
  if (s->G > 1) s->comm.barrier(s->comm.ctx);                       // nobody still reads my buffers
  for (void *p : s->opened) modgpuPeerClose(p);
  s->opened.clear();
  for (int i = 0; i < 4; ++i) { if (s->mine[i]) modgpuPeerFree(s->mine[i]); s->mine[i] = nullptr; }
  s->reserved = false;
}

extern "C" void modgpuShardedDestroy(ModgpuSharded *s)
{
  if (!s) return;
  release_peers(s);
  s->cursors.release(); s->ovf.release(); s->rowsSend.release(); s->rowsRecv.release(); s->misc.release(); s->hMisc.release();
  if (s->ms) modgpuModsetDestroy(s->ms);
  delete s;
}

extern "C" ModgpuModset *modgpuShardedLocal(ModgpuSharded *s) { return s->ms; }

extern "C" int modgpuShardedSetStream(ModgpuSharded *s, void *stream) { return modgpuModsetSetStream(s->ms, stream); }

// COLLECTIVE: size and map the peer buckets for batches of up to maxBasesPerBatch bases per rank
extern "C" int modgpuShardedReserve(ModgpuSharded *s, uint64_t maxBasesPerBatch)
{
  const int G = s->G;
  release_peers(s);
  std::vector<uint64_t> all(G, maxBasesPerBatch);
  if (G > 1 && s->comm.allgather(s->comm.ctx, &maxBasesPerBatch, all.data(), 8)) { mg_set_error("modgpuShardedReserve: allgather failed"); return MODGPU_ECUDA; }
  uint64_t nb = 0;
  for (uint64_t x : all) nb = x > nb ? x : nb;
  const uint64_t w = (uint64_t)s->ms->hasher.w;
  const uint64_t expected = (nb / (w ? w : 1) + 1) * s->accumulate;        // a group of batches shares the buckets
  const double mean = (double)expected / ((double)G * s->R);
  // bucket capacity: Poisson mean + 10 % + 4 sigma; what does not fit travels in the owner's overflow segment
  uint64_t cap = ((uint64_t)(1.1 * mean + 4.0 * sqrt(mean) + 8.0) + 1) & ~1ull;
  if (cap > 0x7FFFFFFFull) { mg_set_error("modgpuShardedReserve: bucket capacity overflow"); return MODGPU_EINVAL; }
  s->cap = (uint32_t)cap;
  // robust: skewed input also selects more than 1/w of its windows (poly-A: every window is the same modimizer), so the
  // worst case is every window of a group, all to one owner
  s->ovfCap = s->robust ? nb * s->accumulate + 65536 : (expected / 4 > 65536 ? expected / 4 : 65536);
  if (s->ovfCap > 0xFFFFFFF0ull) s->ovfCap = 0xFFFFFFF0ull;
  if (s->robust)
    { size_t freeB = 0, totalB = 0;
      cudaMemGetInfo(&freeB, &totalB);
      if ((double)G * (double)s->ovfCap * 16.0 > 0.5 * (double)freeB)
        { mg_set_error("modgpuShardedReserve: robust overflow segments for batches of %llu bases need %.1f GB per GPU: feed smaller batches",
                       (unsigned long long)nb, (double)G * (double)s->ovfCap * 16.0 / 1e9);
          return MODGPU_ENOMEM;
        }
    }
  const size_t bBytes = (size_t)G * s->R * cap * 8, oBytes = (size_t)G * s->ovfCap * 8;
  int ok = 1;
  for (int i = 0; i < 4; ++i) { s->mine[i] = modgpuPeerAlloc(i < 2 ? bBytes : oBytes); if (!s->mine[i]) ok = 0; }
  struct Blob { int ok; unsigned char h[4][MODGPU_PEER_HANDLE_BYTES]; };
  Blob me; memset(&me, 0, sizeof(me));
  for (int i = 0; ok && i < 4; ++i) if (modgpuPeerExport(s->mine[i], me.h[i])) ok = 0;
  me.ok = ok;
  std::vector<Blob> blobs(G);
  blobs[s->rank] = me;
  if (G > 1 && s->comm.allgather(s->comm.ctx, &me, blobs.data(), sizeof(Blob))) { mg_set_error("modgpuShardedReserve: allgather failed"); return MODGPU_ECUDA; }
  for (int r = 0; r < G; ++r) ok = ok && blobs[r].ok;
  for (int src = 0; ok && src < G; ++src)
    { void *p[4];
      for (int i = 0; i < 4; ++i)
        { if (src == s->rank) p[i] = s->mine[i];
          else
            { p[i] = modgpuPeerOpen(blobs[src].h[i]);
              if (!p[i]) { ok = 0; break; }
              s->opened.push_back(p[i]);
            }
        }
      if (!ok) break;
      for (int b = 0; b < 2; ++b)                                      // already offset to THIS owner's part of the source's arrays
        { s->bptr[b][src] = (const uint64_t *)p[b] + (size_t)s->rank * s->R * cap;
          s->optr[b][src] = (const uint64_t *)p[2 + b] + (size_t)s->rank * s->ovfCap;
        }
    }
  // every rank must have mapped every peer
  std::vector<int> oks(G, ok);
  if (G > 1 && s->comm.allgather(s->comm.ctx, &ok, oks.data(), sizeof(int))) { mg_set_error("modgpuShardedReserve: allgather failed"); return MODGPU_ECUDA; }
  for (int r = 0; r < G; ++r) ok = ok && oks[r];
  s->reserved = true;                                                  // (so that release_peers frees what was allocated)
  if (!ok)
    { release_peers(s);
      if (!modgpuLastError()[0]) mg_set_error("modgpuShardedReserve: a rank could not allocate or map the peer buckets");
      return MODGPU_ECUDA;
    }
  int rc;
  const size_t rowWords = (size_t)G * (s->R + 2);
  if ((rc = s->cursors.ensure((size_t)G * s->R * 4 + 64)) || (rc = s->ovf.ensure((size_t)G * 4 + 64)) ||
      (rc = s->rowsSend.ensure(rowWords * 4)) || (rc = s->rowsRecv.ensure(rowWords * 4)))
    return rc;
  s->reservedBases = nb; s->batch = 0; s->pending = 0;
  return MODGPU_OK;
}

extern "C" int modgpuShardedSetAccumulate(ModgpuSharded *s, int nBatches)
{
  int rc = modgpuShardedFlush(s);
  if (rc) return rc;
  release_peers(s);                                                    // the buckets are sized for a group: mapped again by the next add
  s->accumulate = nBatches > 1 ? (uint32_t)nBatches : 1u;
  return MODGPU_OK;
}

extern "C" int modgpuShardedSetRobust(ModgpuSharded *s, int on)
{
  int rc = modgpuShardedFlush(s);
  if (rc) return rc;
  release_peers(s);
  s->robust = on != 0;
  return MODGPU_OK;
}

// COLLECTIVE: the fill counts of the waiting batches go to the owners, the owners build from the peers' buckets
extern "C" int modgpuShardedFlush(ModgpuSharded *s)
{
  if (s->G == 1) return modgpuModsetFlush(s->ms);
  if (!s->reserved || !s->pending) return MODGPU_OK;
  ModgpuModset *ms = s->ms;
  cudaStream_t st = ms->stream;
  const int b = (int)(s->batch & 1);
  ++s->batch; s->pending = 0;
  const uint32_t G = (uint32_t)s->G, R = s->R;
  { ProfScope p(ms, MODGPU_T_OTHER, 1);
    rows_pack_kernel<<<(unsigned)mg_num_sms() * 2, 256, 0, st>>>((const uint32_t *)s->cursors.p, (const uint32_t *)s->ovf.p, G, R,
                                                                s->ovfCap > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)s->ovfCap, (uint32_t *)s->rowsSend.p);
    MG_LAUNCH_CHECK("rows_pack");
  }
  if (s->comm.alltoall(s->comm.ctx, s->rowsSend.p, s->rowsRecv.p, (size_t)(R + 2) * 4, (void *)st))
    { mg_set_error("modgpuShardedFlush: the all-to-all of the fill counts failed"); return MODGPU_ECUDA; }
  { ProfScope p(ms, MODGPU_T_OTHER, 1);
    rows_finish_kernel<<<1, 1, 0, st>>>((const uint32_t *)s->rowsRecv.p, G, R, sh_guard(s), sh_cnt(s), sh_selacc(s), sh_skipacc(s));
    MG_LAUNCH_CHECK("rows_finish");
  }
  { ProfScope p(ms, MODGPU_T_INSERT, 2);
    const uint32_t *rows = (const uint32_t *)s->rowsRecv.p;
    int rc = mg_table_build_from_peers_ex(ms->table, s->bptr[b], rows, R + 2, s->cap, G, s->optr[b], s->ovfCap, rows + R, R + 2, sh_guard(s), st);
    if (rc) return rc;
  }
  ms->dirty = true;
  return MODGPU_OK;
}

template <class SEL>
static int sharded_add(ModgpuSharded *s, uint64_t nBases, SEL select)
{
  if (!s->reserved || nBases > s->reservedBases + s->reservedBases / 8)
    { // (a larger batch than reserved would still work - its surplus travels in the overflow segments - but remapping
      // is cheap next to the skipped group a too-small segment costs; COLLECTIVE, so every rank must see the same sizes:
      // callers with uneven batches reserve explicitly)
      int rc = modgpuShardedFlush(s);
      if (rc) return rc;
      if ((rc = modgpuShardedReserve(s, nBases))) return rc;
    }
  ModgpuModset *ms = s->ms;
  const int b = (int)(s->batch & 1);
  const int keep = ms->selFlags;
  if (s->pending) ms->selFlags |= MODGPU_SEL_APPEND;                 // joins the batches already waiting: the fill counts are kept
  int rc = select((uint64_t *)s->mine[b], s->cap, (uint32_t *)s->cursors.p, (uint64_t *)s->mine[2 + b], s->ovfCap, (uint32_t *)s->ovf.p, sh_cnt(s));
  ms->selFlags = keep;
  if (rc) return rc;
  if (++s->pending >= s->accumulate) return modgpuShardedFlush(s);
  return MODGPU_OK;
}

extern "C" int modgpuShardedAddDevice(ModgpuSharded *s, const uint8_t *d_bases, const uint64_t *d_offs, uint64_t nSeq,
                                      uint64_t nBases, int isAscii)
{
  if (s->G == 1)
    { const uint64_t n = modgpuModsetAddDevice(s->ms, d_bases, d_offs, nSeq, nBases, isAscii);
      if (n == 0xFFFFFFFFFFFFFFFFull) return MODGPU_ECUDA;
      ((uint64_t *)s->hMisc.p)[8] += n;
      return MODGPU_OK;
    }
  return sharded_add(s, nBases, [&](uint64_t *sb, uint32_t cap, uint32_t *cur, uint64_t *so, uint64_t oc, uint32_t *ovf, uint64_t *cnt) {
    return modgpuModsetSelectBucketsDevice(s->ms, d_bases, d_offs, nSeq, nBases, isAscii, (uint32_t)s->G, sb, cap, cur, so, oc, ovf, cnt);
  });
}

extern "C" int modgpuShardedAdd(ModgpuSharded *s, const char *bases, const uint64_t *offs, uint64_t nSeq, int isAscii)
{
  if (!offs) { mg_set_error("modgpuShardedAdd: null offsets"); return MODGPU_EINVAL; }
  if (s->G == 1)
    { const uint64_t n = modgpuModsetAdd(s->ms, bases, offs, nSeq, isAscii);
      if (n == 0xFFFFFFFFFFFFFFFFull) return MODGPU_ECUDA;
      ((uint64_t *)s->hMisc.p)[8] += n;
      return MODGPU_OK;
    }
  return sharded_add(s, nSeq ? offs[nSeq] : 0, [&](uint64_t *sb, uint32_t cap, uint32_t *cur, uint64_t *so, uint64_t oc, uint32_t *ovf, uint64_t *cnt) {
    return modgpuModsetSelectBucketsHost(s->ms, bases, offs, nSeq, isAscii, (uint32_t)s->G, sb, cap, cur, so, oc, ovf, cnt);
  });
}

// COLLECTIVE: finish the waiting groups; *nSelected = k-mers THIS rank selected (and that reached their owners) since
// the last call.  MODGPU_ESKEW: at least one group was skipped on every rank (nothing of it was applied).
extern "C" int modgpuShardedSynchronize(ModgpuSharded *s, uint64_t *nSelected)
{
  uint64_t *h = (uint64_t *)s->hMisc.p;
  if (s->G == 1)
    { int rc = modgpuModsetFlush(s->ms);
      if (nSelected) *nSelected = h[8];
      h[8] = 0;
      return rc;
    }
  int rc = modgpuShardedFlush(s);
  if (rc) return rc;
  cudaStream_t st = s->ms->stream;
  MG_CUDA(cudaMemcpyAsync(h, sh_selacc(s), 16, cudaMemcpyDeviceToHost, st));
  MG_CUDA(cudaMemsetAsync(sh_selacc(s), 0, 16, st));
  MG_CUDA(cudaStreamSynchronize(st));
  if (nSelected) *nSelected = h[0];
  if ((uint32_t)h[1])
    { mg_set_error("sharded add: an overflow segment overflowed (heavily skewed batches); the affected group(s) were skipped on "
                   "every rank - nothing of them was applied.  modgpuShardedSetRobust(s, 1) and add those batches again");
      return MODGPU_ESKEW;
    }
  if (modgpuTableEntries(s->ms->table, st) == 0xFFFFFFFFFFFFFFFFull) return MODGPU_EFULL;
  return MODGPU_OK;
}

extern "C" int modgpuShardedClear(ModgpuSharded *s)
{
  s->pending = 0;                       // batches waiting in the peer buckets are dropped with the rest (no peer has been told to read them)
  ((uint64_t *)s->hMisc.p)[8] = 0;
  return modgpuModsetClear(s->ms);
}

// ------------------------------------------------------------------ NCCL --
// libnccl is loaded at run time (the torch-bundled or the system one, whichever the process already has or finds):
// libmodgpu.so has no link dependency on it.  Only the handful of entry points below are used.
typedef int (*nccl_fn0)(void);
typedef int (*nccl_sendrecv)(void *, size_t, int, int, void *, void *);
typedef int (*nccl_allgather)(const void *, void *, size_t, int, void *, void *);
struct NcclBind { void *comm; int rank, world; nccl_fn0 groupStart, groupEnd; nccl_sendrecv send, recv; nccl_allgather allGather;
                  void *d_stage; size_t stageBytes; cudaStream_t st; };

static int nccl_alltoall(void *ctx, const void *d_send, void *d_recv, size_t bytesPerPeer, void *stream)
{
  NcclBind *n = (NcclBind *)ctx;
  if (n->groupStart()) return -1;
  for (int p = 0; p < n->world; ++p)
    { if (n->send((void *)((const char *)d_send + (size_t)p * bytesPerPeer), bytesPerPeer, 0 /* ncclInt8 */, p, n->comm, stream)) return -1;
      if (n->recv((char *)d_recv + (size_t)p * bytesPerPeer, bytesPerPeer, 0, p, n->comm, stream)) return -1;
    }
  return n->groupEnd() ? -1 : 0;
}

static int nccl_allgather_host(void *ctx, const void *in, void *out, size_t bytes)
{
  NcclBind *n = (NcclBind *)ctx;
  const size_t need = bytes * (size_t)(n->world + 1);
  if (need > n->stageBytes)
    { if (n->d_stage) cudaFree(n->d_stage);
      n->d_stage = nullptr; n->stageBytes = 0;
      if (cudaMalloc(&n->d_stage, need + 4096) != cudaSuccess) return -1;
      n->stageBytes = need + 4096;
    }
  char *dIn = (char *)n->d_stage, *dOut = dIn + bytes;
  if (cudaMemcpyAsync(dIn, in, bytes, cudaMemcpyHostToDevice, n->st) != cudaSuccess) return -1;
  if (n->allGather(dIn, dOut, bytes, 0, n->comm, n->st)) return -1;
  if (cudaMemcpyAsync(out, dOut, bytes * (size_t)n->world, cudaMemcpyDeviceToHost, n->st) != cudaSuccess) return -1;
  return cudaStreamSynchronize(n->st) == cudaSuccess ? 0 : -1;
}

static int nccl_barrier(void *ctx)
{
  NcclBind *n = (NcclBind *)ctx;
  int x = 0;
  std::vector<int> all((size_t)n->world);
  return nccl_allgather_host(ctx, &x, all.data(), sizeof(int));
}

extern "C" int modgpuCommFromNccl(ModgpuComm *out, void *ncclComm, int rank, int world)
{
  if (!out || !ncclComm) { mg_set_error("modgpuCommFromNccl: null argument"); return MODGPU_EINVAL; }
  void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { mg_set_error("modgpuCommFromNccl: libnccl.so.2 not found (%s)", dlerror()); return MODGPU_EINVAL; }
  NcclBind *n = new NcclBind();
  memset(n, 0, sizeof(*n));
  n->comm = ncclComm; n->rank = rank; n->world = world;
  n->groupStart = (nccl_fn0)dlsym(lib, "ncclGroupStart"); n->groupEnd = (nccl_fn0)dlsym(lib, "ncclGroupEnd");
  n->send = (nccl_sendrecv)dlsym(lib, "ncclSend"); n->recv = (nccl_sendrecv)dlsym(lib, "ncclRecv");
  n->allGather = (nccl_allgather)dlsym(lib, "ncclAllGather");
  if (!n->groupStart || !n->groupEnd || !n->send || !n->recv || !n->allGather)
    { delete n; mg_set_error("modgpuCommFromNccl: libnccl lacks a required entry point"); return MODGPU_EINVAL; }
  if (mg_check_cuda(cudaStreamCreateWithFlags(&n->st, cudaStreamNonBlocking), "cudaStreamCreate", __FILE__, __LINE__)) { delete n; return MODGPU_ECUDA; }
  out->ctx = n; out->rank = rank; out->world = world;
  out->alltoall = nccl_alltoall; out->allgather = nccl_allgather_host; out->barrier = nccl_barrier;
  return MODGPU_OK;
}

extern "C" void modgpuCommNcclRelease(ModgpuComm *c)
{
  if (!c || !c->ctx) return;
  NcclBind *n = (NcclBind *)c->ctx;
  if (n->d_stage) cudaFree(n->d_stage);
  if (n->st) cudaStreamDestroy(n->st);
  delete n;
  c->ctx = nullptr;
}
