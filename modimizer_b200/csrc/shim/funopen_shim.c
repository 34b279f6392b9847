/* funopen_shim.c - BSD funopen() on top of glibc fopencookie(), so that the
 * unmodified reference utils.c (fzopen, utils.c:108-127) links on Linux.
 * Test infrastructure for oracle/_ref only. */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <sys/types.h>
#include "funopen_shim.h"

typedef struct {
  void *cookie;
  int (*rd)(void *, char *, int);
  int (*wr)(void *, const char *, int);
  fpos_t (*sk)(void *, fpos_t, int);
  int (*cl)(void *);
} Shim;

static ssize_t shim_read(void *c, char *buf, size_t n)
{ Shim *s = (Shim *)c; return s->rd ? (ssize_t)s->rd(s->cookie, buf, (int)n) : -1; }

static ssize_t shim_write(void *c, const char *buf, size_t n)
{ Shim *s = (Shim *)c; return s->wr ? (ssize_t)s->wr(s->cookie, buf, (int)n) : -1; }

static int shim_close(void *c)
{ Shim *s = (Shim *)c; int r = s->cl ? s->cl(s->cookie) : 0; free(s); return r; }

FILE *funopen(const void *cookie,
              int (*readfn)(void *, char *, int),
              int (*writefn)(void *, const char *, int),
              fpos_t (*seekfn)(void *, fpos_t, int),
              int (*closefn)(void *))
{
  Shim *s = (Shim *)calloc(1, sizeof(Shim));
  if (!s) return 0;
  s->cookie = (void *)cookie; s->rd = readfn; s->wr = writefn; s->sk = seekfn; s->cl = closefn;
  /* seeking a gz stream through the reference's cast of gzseek is not needed by
     any caller on the path (modsetWrite/Read stream sequentially) */
  cookie_io_functions_t io = { shim_read, shim_write, 0, shim_close };
  const char *mode = (readfn && writefn) ? "r+" : writefn ? "w" : "r";
  return fopencookie(s, mode, io);
}
