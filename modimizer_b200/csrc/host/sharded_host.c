/* sharded_host.c - the multi-GPU modset build as a C host: one process per GPU, NCCL for the plumbing, no Python.
 *
 * The reference's only parallel recipe is "one modset per input, then modsetMerge" (modset.c:106-128,
 * modutils.c:101-103).  This driver is what a C caller of libmodgpu does instead: every rank feeds ITS shard of the
 * input to modgpuShardedAdd*, the table is sharded by k-mer hash, the owners build from the peers' buckets over NVLink.
 *
 *   sharded_host --gpus N [--gbases 3.1] [--k 31] [--d 64] [--bits 28] [--steps 5] [--warmup 3] [--accumulate 1]
 *                [--check MBASES] [--skew]
 *
 * Input: the synthetic genome of include/modgpu_synth.h, rank r taking bases [r * n, (r+1) * n), generated on the device
 * (not timed).  A step = clear + add + synchronize, timed with CUDA events on the modset's stream, max over ranks.
 * --check M: every rank also feeds the first M Mbases of its shard into a second sharded set; rank 0 builds ONE
 * single-GPU modset from the concatenation of all ranks' samples and the summed shard histograms / entry counts must
 * equal it.  --skew: a poly-A shard (every k-mer to one bucket of one owner) must be reported as MODGPU_ESKEW by every
 * rank with nothing applied, and succeed after modgpuShardedSetRobust.
 * Prints one JSON line on rank 0.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <unistd.h>
#include <sys/wait.h>
#include <cuda_runtime.h>
#include <nccl.h>
#include "modgpu.h"

#define DIE(...) do { fprintf (stderr, "sharded_host[%d]: ", rank) ; fprintf (stderr, __VA_ARGS__) ; fputc ('\n', stderr) ; exit (2) ; } while (0)
#define MG(x) do { if ((x) != 0) DIE ("%s: %s", #x, modgpuLastError ()) ; } while (0)
#define CU(x) do { cudaError_t e_ = (x) ; if (e_ != cudaSuccess) DIE ("%s: %s", #x, cudaGetErrorString (e_)) ; } while (0)
#define NC(x) do { ncclResult_t r_ = (x) ; if (r_ != ncclSuccess) DIE ("%s: %s", #x, ncclGetErrorString (r_)) ; } while (0)

static int rank ;

static double argd (int argc, char **argv, const char *name, double def)
{ for (int i = 1 ; i + 1 < argc ; ++i) if (!strcmp (argv[i], name)) return atof (argv[i+1]) ; return def ; }
static int argf (int argc, char **argv, const char *name)
{ for (int i = 1 ; i < argc ; ++i) if (!strcmp (argv[i], name)) return 1 ; return 0 ; }

static int worker (int world, ncclUniqueId id, int argc, char **argv)
{
  const double gbases = argd (argc, argv, "--gbases", 3.1) ;
  const int k = (int) argd (argc, argv, "--k", 31), d = (int) argd (argc, argv, "--d", 64), bits = (int) argd (argc, argv, "--bits", 28) ;
  const int steps = (int) argd (argc, argv, "--steps", 5), warmup = (int) argd (argc, argv, "--warmup", 3) ;
  const int accumulate = (int) argd (argc, argv, "--accumulate", 1) ;
  const double checkM = argd (argc, argv, "--check", 0) ;
  CU (cudaSetDevice (rank)) ;
  MG (modgpuSetDevice (rank)) ;
  ncclComm_t nc ;
  NC (ncclCommInitRank (&nc, world, id, rank)) ;
  ModgpuComm comm ;
  MG (modgpuCommFromNccl (&comm, nc, rank, world)) ;

  uint64_t nb = (uint64_t) (gbases * 1e9) ; nb -= nb % 32 ;
  if (nb >= ((uint64_t) 1 << 32) - (1 << 20)) DIE ("--gbases: at most 4.29 per rank and batch") ;
  uint8_t *dBases ; uint64_t *dOffs ;
  CU (cudaMalloc ((void**) &dBases, nb + 64)) ;
  const int nRec = 24 ;
  uint64_t hOffs[25] ;
  for (int i = 0 ; i <= nRec ; ++i) hOffs[i] = (uint64_t) ((double) nb / nRec * i) ;
  hOffs[nRec] = nb ;
  CU (cudaMalloc ((void**) &dOffs, sizeof (hOffs))) ;
  CU (cudaMemcpy (dOffs, hOffs, sizeof (hOffs), cudaMemcpyHostToDevice)) ;
  MG (modgpuSynthGenome (12345, (uint64_t) rank * nb, nb, 1, dBases, 0)) ;
  CU (cudaDeviceSynchronize ()) ;

  ModgpuSharded *sh = modgpuShardedCreate (bits, k, d, 17, &comm) ;
  if (!sh) DIE ("modgpuShardedCreate: %s", modgpuLastError ()) ;
  if (accumulate > 1) MG (modgpuShardedSetAccumulate (sh, accumulate)) ;
  MG (modgpuShardedReserve (sh, nb)) ;

  cudaEvent_t e0, e1 ; CU (cudaEventCreate (&e0)) ; CU (cudaEventCreate (&e1)) ;
  uint64_t nSel = 0 ;
  float msTotal = 0 ;
  for (int s = 0 ; s < warmup + steps ; ++s)
    { if (s == warmup)
	{ CU (cudaDeviceSynchronize ()) ; comm.barrier (comm.ctx) ;
	  CU (cudaEventRecord (e0, 0)) ;               /* legacy stream: ordered against the modset's (blocking) work by the syncs */
	}
      MG (modgpuShardedClear (sh)) ;
      for (int a = 0 ; a < accumulate ; ++a) MG (modgpuShardedAddDevice (sh, dBases, dOffs, nRec, nb, 0)) ;
      MG (modgpuShardedSynchronize (sh, &nSel)) ;
    }
  CU (cudaDeviceSynchronize ()) ;
  CU (cudaEventRecord (e1, 0)) ; CU (cudaEventSynchronize (e1)) ;
  CU (cudaEventElapsedTime (&msTotal, e0, e1)) ;
  double msMine = msTotal / steps, *msAll = (double*) calloc (world, sizeof (double)) ;
  comm.allgather (comm.ctx, &msMine, msAll, sizeof (double)) ;
  double msMax = 0 ; for (int r = 0 ; r < world ; ++r) if (msAll[r] > msMax) msMax = msAll[r] ;
  uint64_t mine[2] = { nSel, modgpuModsetMax (modgpuShardedLocal (sh)) }, *all = (uint64_t*) calloc (2 * world, 8) ;
  comm.allgather (comm.ctx, mine, all, 16) ;
  uint64_t totSel = 0, totEntries = 0 ;
  for (int r = 0 ; r < world ; ++r) { totSel += all[2*r] ; totEntries += all[2*r+1] ; }

  /* ---- parity: sharded over `world` GPUs == one modset on one GPU, on a bounded sample of every shard */
  int parityOk = -1 ; uint64_t sample = 0 ;
  if (checkM > 0)
    { sample = (uint64_t) (checkM * 1e6) ; if (sample > nb) sample = nb ; sample -= sample % 32 ;
      uint64_t one[2] = { 0, sample } ;
      uint64_t *dOne ; CU (cudaMalloc ((void**) &dOne, 16)) ; CU (cudaMemcpy (dOne, one, 16, cudaMemcpyHostToDevice)) ;
      int pbits = 20 ; while (((uint64_t) 1 << (pbits - 2)) < 2 * (sample / (uint64_t) d + 1024) && pbits < 30) ++pbits ;
      ModgpuSharded *ps = modgpuShardedCreate (pbits, k, d, 17, &comm) ;
      if (!ps) DIE ("modgpuShardedCreate: %s", modgpuLastError ()) ;
      uint64_t psel = 0 ;
      MG (modgpuShardedAddDevice (ps, dBases, dOne, 1, sample, 0)) ;
      MG (modgpuShardedSynchronize (ps, &psel)) ;
      uint32_t *hist = (uint32_t*) calloc (65536, 4), *histAll = (uint32_t*) calloc ((size_t) 65536 * world, 4) ;
      MG (modgpuModsetHistogram (modgpuShardedLocal (ps), hist)) ;
      comm.allgather (comm.ctx, hist, histAll, 65536 * 4) ;
      uint64_t pm[2] = { psel, modgpuModsetMax (modgpuShardedLocal (ps)) }, *pmAll = (uint64_t*) calloc (2 * world, 8) ;
      comm.allgather (comm.ctx, pm, pmAll, 16) ;
      if (rank == 0)
	{ /* the same samples into ONE table on this GPU (capacity for all of them) */
	  int obits = pbits ; while (((uint64_t) 1 << (obits - 2)) < 2 * (uint64_t) world * (sample / (uint64_t) d + 1024) && obits < 32) ++obits ;
	  ModgpuModset *single = modgpuModsetCreate (obits, k, d, 17) ;
	  if (!single) DIE ("modgpuModsetCreate: %s", modgpuLastError ()) ;
	  uint8_t *dTmp ; CU (cudaMalloc ((void**) &dTmp, sample + 64)) ;
	  uint64_t tot = 0 ;
	  for (int r = 0 ; r < world ; ++r)
	    { MG (modgpuSynthGenome (12345, (uint64_t) r * nb, sample, 1, dTmp, 0)) ;
	      CU (cudaDeviceSynchronize ()) ;
	      uint64_t n = modgpuModsetAddDevice (single, dTmp, dOne, 1, sample, 0) ;
	      if (n == UINT64_MAX) DIE ("modgpuModsetAddDevice: %s", modgpuLastError ()) ;
	      tot += n ;
	    }
	  uint32_t *h1 = (uint32_t*) calloc (65536, 4) ;
	  MG (modgpuModsetHistogram (single, h1)) ;
	  uint64_t sSel = 0, sEnt = 0 ;
	  for (int r = 0 ; r < world ; ++r) { sSel += pmAll[2*r] ; sEnt += pmAll[2*r+1] ; }
	  parityOk = (sSel == tot && sEnt == modgpuModsetMax (single)) ;
	  for (int bin = 0 ; bin < 65536 && parityOk ; ++bin)
	    { uint64_t sum = 0 ; for (int r = 0 ; r < world ; ++r) sum += histAll[(size_t) r * 65536 + bin] ;
	      if (sum != h1[bin]) parityOk = 0 ;
	    }
	  modgpuModsetDestroy (single) ; cudaFree (dTmp) ; free (h1) ;
	}
      modgpuShardedDestroy (ps) ;
      cudaFree (dOne) ; free (hist) ; free (histAll) ; free (pmAll) ;
    }

  /* ---- skew: a group that cannot fit its overflow segments is skipped EVERYWHERE, then fits after SetRobust */
  int skewOk = -1 ;
  if (argf (argc, argv, "--skew"))
    { const uint64_t n = 1 << 24 ;
      uint8_t *dA ; CU (cudaMalloc ((void**) &dA, n + 64)) ; CU (cudaMemset (dA, 0, n + 64)) ;   /* poly-A */
      uint64_t one[2] = { 0, n }, *dOne ; CU (cudaMalloc ((void**) &dOne, 16)) ; CU (cudaMemcpy (dOne, one, 16, cudaMemcpyHostToDevice)) ;
      ModgpuSharded *ps = modgpuShardedCreate (24, 15, 4, 17, &comm) ;
      if (!ps) DIE ("modgpuShardedCreate: %s", modgpuLastError ()) ;
      uint64_t sel = 0 ;
      MG (modgpuShardedAddDevice (ps, dA, dOne, 1, n, 0)) ;
      int rc = modgpuShardedSynchronize (ps, &sel) ;
      uint64_t entries = 0 ;
      { uint64_t m = modgpuModsetMax (modgpuShardedLocal (ps)), *mAll = (uint64_t*) calloc (world, 8) ;
	comm.allgather (comm.ctx, &m, mAll, 8) ; for (int r = 0 ; r < world ; ++r) entries += mAll[r] ; free (mAll) ; }
      skewOk = (rc == MODGPU_ESKEW && entries == 0) ;                       /* reported, and nothing was applied anywhere */
      MG (modgpuShardedSetRobust (ps, 1)) ;
      MG (modgpuShardedAddDevice (ps, dA, dOne, 1, n, 0)) ;
      rc = modgpuShardedSynchronize (ps, &sel) ;
      uint32_t *hist = (uint32_t*) calloc (65536, 4), *histAll = (uint32_t*) calloc ((size_t) 65536 * world, 4) ;
      MG (modgpuModsetHistogram (modgpuShardedLocal (ps), hist)) ;
      comm.allgather (comm.ctx, hist, histAll, 65536 * 4) ;
      uint64_t top = 0, nEnt = 0 ;
      for (int r = 0 ; r < world ; ++r) { top += histAll[(size_t) r * 65536 + 65535] ; for (int b = 0 ; b < 65536 ; ++b) nEnt += histAll[(size_t) r * 65536 + b] ; }
      /* poly-A at k=15 d=4: one k-mer (hash 0), seen n-14 times by every rank, depth saturates */
      skewOk = skewOk && rc == 0 && sel == n - 14 && nEnt == 1 && top == 1 ;
      modgpuShardedDestroy (ps) ; cudaFree (dA) ; cudaFree (dOne) ; free (hist) ; free (histAll) ;
    }

  if (rank == 0)
    { printf ("{\"tool\": \"sharded_host (C, NCCL)\", \"n_gpus\": %d, \"k\": %d, \"d\": %d, \"tableBits_per_gpu\": %d, \"bases_per_gpu\": %llu, "
	      "\"accumulate\": %d, \"steps\": %d, \"warmup\": %d, \"ms_per_step\": %.4f, \"gbases_per_s\": %.1f, \"selected\": %llu, \"entries\": %llu",
	      world, k, d, bits, (unsigned long long) nb, accumulate, steps, warmup, msMax,
	      (double) world * nb * accumulate / (msMax * 1e-3) / 1e9, (unsigned long long) totSel, (unsigned long long) totEntries) ;
      if (parityOk >= 0) printf (", \"parity\": {\"bases_per_gpu\": %llu, \"against\": \"one modset on one GPU from all ranks' samples\", \"ok\": %s}",
				 (unsigned long long) sample, parityOk ? "true" : "false") ;
      if (skewOk >= 0) printf (", \"skew_transactional\": %s", skewOk ? "true" : "false") ;
      printf ("}\n") ;
    }
  modgpuShardedDestroy (sh) ;
  modgpuCommNcclRelease (&comm) ;
  ncclCommDestroy (nc) ;
  return (parityOk == 0 || skewOk == 0) ? 1 : 0 ;
}

int main (int argc, char **argv)
{
  int world = (int) argd (argc, argv, "--gpus", 1) ;
  if (world < 1 || world > MODGPU_MAX_PEERS) { fprintf (stderr, "--gpus 1..%d\n", MODGPU_MAX_PEERS) ; return 2 ; }
  /* fork first (no CUDA / NCCL state in the parent yet), then rank 0 hands the NCCL id to the others through pipes */
  ncclUniqueId id ;
  pid_t pids[MODGPU_MAX_PEERS] ;
  int pipes[MODGPU_MAX_PEERS][2] ;
  for (int r = 1 ; r < world ; ++r)
    { if (pipe (pipes[r])) { perror ("pipe") ; return 2 ; }
      pids[r] = fork () ;
      if (pids[r] < 0) { perror ("fork") ; return 2 ; }
      if (!pids[r])
	{ rank = r ; close (pipes[r][1]) ;
	  if (read (pipes[r][0], &id, sizeof (id)) != (ssize_t) sizeof (id)) { fprintf (stderr, "rank %d: no NCCL id\n", r) ; return 2 ; }
	  close (pipes[r][0]) ;
	  return worker (world, id, argc, argv) ;
	}
      close (pipes[r][0]) ;
    }
  rank = 0 ;
  if (ncclGetUniqueId (&id) != ncclSuccess) { fprintf (stderr, "ncclGetUniqueId failed\n") ; return 2 ; }
  for (int r = 1 ; r < world ; ++r) { if (write (pipes[r][1], &id, sizeof (id)) != (ssize_t) sizeof (id)) return 2 ; close (pipes[r][1]) ; }
  int rc = worker (world, id, argc, argv) ;
  for (int r = 1 ; r < world ; ++r) { int st = 0 ; waitpid (pids[r], &st, 0) ; if (!WIFEXITED (st) || WEXITSTATUS (st)) rc = rc ? rc : 3 ; }
  return rc ;
}
