// hostcopy_bench.c -- host copy bandwidth of the box: glibc memcpy vs explicit non-temporal stores, 1..16 threads.
// gcc -O2 -mavx2 -pthread tools/hostcopy_bench.c -o /tmp/hostcopy_bench && /tmp/hostcopy_bench
#include <immintrin.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static void copy_nt(void *dst, const void *src, size_t n) {
  const __m256i *s = (const __m256i *)src; __m256i *d = (__m256i *)dst;
  for (size_t i = 0; i < n / 32; i += 4) {
    __m256i a = _mm256_loadu_si256(s + i), b = _mm256_loadu_si256(s + i + 1), c = _mm256_loadu_si256(s + i + 2), e = _mm256_loadu_si256(s + i + 3);
    _mm256_stream_si256(d + i, a); _mm256_stream_si256(d + i + 1, b); _mm256_stream_si256(d + i + 2, c); _mm256_stream_si256(d + i + 3, e);
  }
  _mm_sfence();
}
typedef struct { char *dst; const char *src; size_t n; int mode; } job_t;
static void *run(void *p) {
  job_t *j = (job_t *)p;
  for (size_t off = 0; off < j->n; off += 1 << 20) {
    if (j->mode) copy_nt(j->dst + off, j->src + off, 1 << 20); else memcpy(j->dst + off, j->src + off, 1 << 20);
  }
  return 0;
}
int main(void) {
  const size_t total = (size_t)2 << 30;
  char *src = aligned_alloc(4096, total), *dst = aligned_alloc(4096, total);
  memset(src, 1, total); memset(dst, 2, total);
  for (int mode = 0; mode < 2; ++mode)
    for (int nt = 1; nt <= 16; nt *= 2) {
      double best = 1e9;
      for (int rep = 0; rep < 3; ++rep) {
        pthread_t th[16]; job_t jb[16];
        double t0 = now();
        for (int t = 0; t < nt; ++t) { jb[t] = (job_t){dst + total / nt * t, src + total / nt * t, total / nt, mode}; pthread_create(&th[t], 0, run, &jb[t]); }
        for (int t = 0; t < nt; ++t) pthread_join(th[t], 0);
        double dt = now() - t0; if (dt < best) best = dt;
      }
      printf("%s threads=%2d  %.1f GB/s\n", mode ? "nt-store" : "memcpy  ", nt, total / best / 1e9);
    }
  return 0;
}
