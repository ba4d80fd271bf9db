// Host staging of pageable columns: what feeds the DMA engine best?
//   A  the library's scheme up to round 2: every column has a full-size pinned buffer; a few threads copy 2 MB chunks into it
//      with non-temporal stores, the caller's thread queues the DMA of every chunk as soon as it has landed.
//      Host DRAM sees the data three times (read source, write staging, DMA reads staging).
//   B  a small RING of pinned chunks that stays in the last-level cache: the threads copy with ordinary stores, the DMA
//      engine reads the chunk out of the cache, the slot is reused once its DMA is complete.  Host DRAM sees the data once.
// One "call" = 63 MB (a 1.21 M-observation file: 52 B/obs), synchronous (the call ends when its last DMA is complete), the
// sources rotate over 8 different buffers.   nvcc -O3 -o tools/stage_ring_bench tools/stage_ring_bench.cu -lpthread
#include <cuda_runtime.h>
#include <immintrin.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <atomic>
#include <thread>
#include <vector>

static double now_ms() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }

__attribute__((target("avx2"))) static void copy_nt(void *dst, const void *src, size_t n) {
  char *d = (char *)dst; const char *s = (const char *)src;
  for (size_t i = 0; i + 128 <= n; i += 128) {
    __m256i a = _mm256_loadu_si256((const __m256i *)(s + i)), b = _mm256_loadu_si256((const __m256i *)(s + i + 32)),
            c = _mm256_loadu_si256((const __m256i *)(s + i + 64)), e = _mm256_loadu_si256((const __m256i *)(s + i + 96));
    _mm256_stream_si256((__m256i *)(d + i), a); _mm256_stream_si256((__m256i *)(d + i + 32), b);
    _mm256_stream_si256((__m256i *)(d + i + 64), c); _mm256_stream_si256((__m256i *)(d + i + 96), e);
  }
  _mm_sfence();
}

struct Job {   // one call: chunks [0, nchunk); chunk i may be copied once i < freed
  const char *src; char *dst_base; size_t chunk, total; int nchunk, ring;   // ring = 0: dst = dst_base + off, else slot i % ring
  bool nt;
  std::atomic<int> next{0}, freed{0}, exited{0};
  std::vector<std::atomic<int>> done;
  Job(int n) : done(n) {}
};
static std::atomic<Job *> g_job{nullptr};
static std::atomic<int> g_gen{0}, g_quit{0};

static void work(Job *j) {
  for (;;) {
    const int i = j->next.fetch_add(1);
    if (i >= j->nchunk) return;
    while (i >= j->freed.load(std::memory_order_acquire)) _mm_pause();
    const size_t off = (size_t)i * j->chunk, n = std::min(j->chunk, j->total - off);
    char *dst = j->ring ? j->dst_base + (size_t)(i % j->ring) * j->chunk : j->dst_base + off;
    if (j->nt) copy_nt(dst, j->src + off, n); else memcpy(dst, j->src + off, n);
    j->done[i].store(1, std::memory_order_release);
  }
}
static void worker() {
  int seen = 0;
  while (!g_quit.load()) {
    if (g_gen.load(std::memory_order_acquire) == seen) { _mm_pause(); continue; }
    seen = g_gen.load();
    Job *j = g_job.load();
    if (j) { work(j); j->exited.fetch_add(1, std::memory_order_release); }
  }
}

int main(int argc, char **argv) {
  const int nthreads = argc > 1 ? atoi(argv[1]) : 7;
  const size_t total = (size_t)63 << 20;
  const int NSRC = 8, CALLS = 24;
  std::vector<char *> src(NSRC);
  for (auto &p : src) { p = (char *)malloc(total); memset(p, 1, total); }
  char *dev; cudaMalloc(&dev, total);
  char *pin_full; cudaMallocHost(&pin_full, total);
  const size_t ring_bytes = (size_t)64 << 20;
  char *pin_ring; cudaMallocHost(&pin_ring, ring_bytes);
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  std::vector<cudaEvent_t> ev(4096);
  for (auto &e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  std::vector<std::thread> th;
  for (int i = 0; i < nthreads; ++i) th.emplace_back(worker);

  auto run = [&](const char *name, size_t chunk, int ring, bool nt) {
    double best = 1e9, sum = 0;
    for (int c = 0; c < CALLS; ++c) {
      const int nchunk = (int)((total + chunk - 1) / chunk);
      Job job(nchunk);
      job.src = src[c % NSRC]; job.dst_base = ring ? pin_ring : pin_full; job.chunk = chunk; job.total = total; job.nchunk = nchunk;
      job.ring = ring; job.nt = nt;
      for (auto &d : job.done) d.store(0);
      job.freed.store(ring ? ring : nchunk);
      const double t0 = now_ms();
      g_job.store(&job);
      g_gen.fetch_add(1, std::memory_order_release);
      int complete = 0;   // DMAs known complete
      for (int i = 0; i < nchunk; ++i) {
        while (!job.done[i].load(std::memory_order_acquire)) {
          if (ring && complete < i && cudaEventQuery(ev[complete]) == cudaSuccess) { ++complete; job.freed.store(ring + complete, std::memory_order_release); }
          else _mm_pause();
        }
        const size_t off = (size_t)i * chunk, n = std::min(chunk, total - off);
        cudaMemcpyAsync(dev + off, ring ? pin_ring + (size_t)(i % ring) * chunk : pin_full + off, n, cudaMemcpyHostToDevice, st);
        if (ring) {
          cudaEventRecord(ev[i], st);
          while (complete <= i && cudaEventQuery(ev[complete]) == cudaSuccess) { ++complete; job.freed.store(ring + complete, std::memory_order_release); }
        }
      }
      cudaStreamSynchronize(st);
      const double dt = now_ms() - t0;
      while (job.exited.load(std::memory_order_acquire) < nthreads) _mm_pause();   // no worker still looks at this job
      g_job.store(nullptr);
      if (c >= 4) { sum += dt; if (dt < best) best = dt; }
    }
    printf("%-58s %2d thr  chunk %4zu KB  ring %3d : mean %.3f ms  best %.3f ms  (%.1f GB/s)\n", name, nthreads, chunk >> 10, ring,
           sum / (CALLS - 4), best, total / (sum / (CALLS - 4)) / 1e6);
    fflush(stdout);
  };
  run("A  full-size staging, non-temporal stores", (size_t)2 << 20, 0, true);
  run("A' full-size staging, memcpy", (size_t)2 << 20, 0, false);
  for (size_t ck : {(size_t)512 << 10, (size_t)1 << 20, (size_t)2 << 20})
    for (int ring : {8, 16, 32}) {
      if ((size_t)ring * ck > ring_bytes) continue;
      run("B  ring, memcpy", ck, ring, false);
      run("B' ring, non-temporal stores", ck, ring, true);
    }
  g_quit.store(1);
  g_gen.fetch_add(1);
  for (auto &t : th) t.join();
  return 0;
}
