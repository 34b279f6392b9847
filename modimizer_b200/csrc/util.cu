// util.cu - error plumbing, device discovery, hasher construction, pinned host
// memory.  Every public entry point of libmodgpu reports failure through a
// negative return code plus modgpuLastError(); nothing falls back to the CPU.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "mg_device.cuh"

static thread_local char g_err[512] = "";

void mg_set_error(const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int mg_check_cuda(cudaError_t e, const char *what, const char *file, int line)
{
  if (e == cudaSuccess) return MODGPU_OK;
  mg_set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  cudaGetLastError();                                   // clear the sticky launch error, keep ours
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return MODGPU_ENODEVICE;
  if (e == cudaErrorMemoryAllocation) return MODGPU_ENOMEM;
  return MODGPU_ECUDA;
}

int mg_num_sms()
{
  static int sms = 0;
  if (!sms)
    { int dev = 0;
      if (cudaGetDevice(&dev) != cudaSuccess ||
          cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
        sms = 148;                                      // B200
    }
  return sms;
}

extern "C" const char *modgpuLastError(void) { return g_err; }
extern "C" const char *modgpuVersion(void) { return "modimizer_b200 0.1 (sm_100a)"; }

extern "C" int modgpuDeviceCount(void)
{
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { mg_check_cuda(e, "cudaGetDeviceCount", __FILE__, __LINE__); return 0; }
  return n;
}

extern "C" int modgpuSetDevice(int dev)
{
  MG_CUDA(cudaSetDevice(dev));
  return MODGPU_OK;
}

// seqhashCreate (reference seqhash.c:20-37): the multiplier comes from libc
// srandom()/random() exactly as in the reference, so that a modset built here
// hashes like one built by the reference on the same box.
extern "C" int modgpuHasherInit(ModgpuHasher *h, int k, int w, int seed)
{
  if (k < 1 || k >= 32) { mg_set_error("seqhash k %d must be between 1 and 32", k); return MODGPU_EINVAL; }
  if (w < 1) { mg_set_error("seqhash w %d must be positive", w); return MODGPU_EINVAL; }
  memset(h, 0, sizeof(*h));
  h->k = k; h->w = w; h->seed = seed;
  h->mask = (((uint64_t)1) << (2 * k)) - 1;
  h->shift1 = 64 - 2 * k;
  srandom((unsigned)seed);
  uint64_t a = (uint64_t)random();
  uint64_t b = (uint64_t)random();
  h->factor1 = (a << 32) | b | 1;
  a = (uint64_t)random();
  b = (uint64_t)random();
  h->factor2 = (a << 32) | b | 1;
  return MODGPU_OK;
}

// layout of the reference's Seqhash (seqhash.h:15-23): int seed,k,w; U64 mask;
// int shift1,shift2; U64 factor1,factor2; U64 patternRC[4]  (80 bytes, LP64)
extern "C" int modgpuHasherFromSeqhash(ModgpuHasher *h, const void *seqhash)
{
  const unsigned char *p = (const unsigned char *)seqhash;
  int32_t seed, k, w, shift1;
  uint64_t mask, f1, f2;
  memcpy(&seed, p + 0, 4); memcpy(&k, p + 4, 4); memcpy(&w, p + 8, 4);
  memcpy(&mask, p + 16, 8); memcpy(&shift1, p + 24, 4);
  memcpy(&f1, p + 32, 8); memcpy(&f2, p + 40, 8);
  if (k < 1 || k >= 32 || w < 1 || shift1 != 64 - 2 * k || !(f1 & 1))
    { mg_set_error("not a reference Seqhash (k %d w %d shift1 %d)", k, w, shift1); return MODGPU_EINVAL; }
  memset(h, 0, sizeof(*h));
  h->k = k; h->w = w; h->seed = seed; h->shift1 = shift1; h->mask = mask; h->factor1 = f1; h->factor2 = f2;
  return MODGPU_OK;
}

extern "C" uint64_t modgpuHash(const ModgpuHasher *h, uint64_t kmer) { return (kmer * h->factor1) >> h->shift1; }

extern "C" void *modgpuHostAlloc(size_t bytes)
{
  void *p = nullptr;
  if (mg_check_cuda(cudaMallocHost(&p, bytes ? bytes : 1), "cudaMallocHost", __FILE__, __LINE__)) return nullptr;
  return p;
}

extern "C" void modgpuHostFree(void *p) { if (p) cudaFreeHost(p); }

// ---- peer memory (multi-GPU exchange over NVLink): buffers another rank's kernels read directly.
// One process per GPU, so the mapping goes through CUDA IPC; cudaIpcOpenMemHandle enables peer
// access between the two devices on first use.
extern "C" void *modgpuPeerAlloc(size_t bytes)
{
  void *p = nullptr;
  if (mg_check_cuda(cudaMalloc(&p, bytes ? bytes : 256), "cudaMalloc (peer buffer)", __FILE__, __LINE__)) return nullptr;
  return p;
}

extern "C" void modgpuPeerFree(void *p) { if (p) cudaFree(p); }

extern "C" int modgpuPeerExport(void *d_ptr, void *handle64)
{
  static_assert(sizeof(cudaIpcMemHandle_t) == MODGPU_PEER_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  MG_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
  memcpy(handle64, &h, sizeof(h));
  return MODGPU_OK;
}

extern "C" void *modgpuPeerOpen(const void *handle64)
{
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void *p = nullptr;
  if (mg_check_cuda(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle", __FILE__, __LINE__)) return nullptr;
  return p;
}

extern "C" int modgpuPeerClose(void *d_ptr)
{
  if (d_ptr) MG_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return MODGPU_OK;
}
