/* funopen_shim.h - prototype of BSD funopen() for glibc builds of the reference.
 * reference utils.c:119 calls funopen() without a declaration; on LP64 glibc an
 * implicit int return truncates the FILE* (SURVEY.md section 8(c)).  Force-included
 * with -include when compiling the reference's utils.c for oracle/_ref. */
#ifndef FUNOPEN_SHIM_H
#define FUNOPEN_SHIM_H
#include <stdio.h>
#include <sys/types.h>
FILE *funopen(const void *cookie,
              int (*readfn)(void *, char *, int),
              int (*writefn)(void *, const char *, int),
              fpos_t (*seekfn)(void *, fpos_t, int),
              int (*closefn)(void *));
#endif
