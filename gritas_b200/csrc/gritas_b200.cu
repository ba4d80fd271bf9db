// gritas_b200.cu -- host side of the C ABI declared in include/gritas_b200.h.
//
// One context per process/GPU.  The count / sum / sum-of-squares accumulators live in HBM for
// the whole run (the reference keeps them in host arrays passed inout to every
// GrITAS_Accum_ call, m_gritas_grids.f:307-311); observation columns are staged through
// double-buffered pinned + device slots so the copy of synoptic time n+1 overlaps the kernel of
// time n (gritas.f:389-808 is the loop this sits in).  There is no CPU implementation of any
// entry point in this file: without a CUDA device create() fails.
#include "../../include/gritas_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <cub/device/device_radix_sort.cuh>

#include "kernels.cuh"

using namespace gb200;

namespace {

constexpr int NSLOT = 2;   // staging slots (double buffering)
constexpr int NWK = 8;     // max worker streams of the tiled path (GRITAS_B200_WORKERS, default 6): the K1 launches
                           // of consecutive calls overlap
constexpr int NCOL = 18;   // staged columns per slot
enum Col { C_LAT, C_LON, C_LEV, C_D0, C_D1, C_D2, C_OBS, C_XVEC, C_KX, C_KT, C_FLAG, C_TIME, C_B0, C_B1, C_B2, C_C0, C_C1, C_C2 };

// --- NCCL through dlopen: the library has no link-time dependency on it ------------------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct NcclApi {
  void *so = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Reduce)(const void *, void *, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool load() {
    if (so) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      so = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (so) break;
    }
    if (!so) return false;
    GetUniqueId = (decltype(GetUniqueId))dlsym(so, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(so, "ncclCommInitRank");
    AllReduce = (decltype(AllReduce))dlsym(so, "ncclAllReduce");
    Reduce = (decltype(Reduce))dlsym(so, "ncclReduce");
    CommDestroy = (decltype(CommDestroy))dlsym(so, "ncclCommDestroy");
    GetErrorString = (decltype(GetErrorString))dlsym(so, "ncclGetErrorString");
    return GetUniqueId && CommInitRank && AllReduce && Reduce && CommDestroy;
  }
};
NcclApi g_nccl;
constexpr int NCCL_FLOAT64 = 8;  // ncclDouble
constexpr int NCCL_SUM = 0;      // ncclSum

// --- host copy pool: pageable caller arrays <-> pinned staging, chunk by chunk, on a few threads ---------------
// An untouched gritas.f hands GrITAS_Accum pageable allocatables (ODS_Get allocates them per synoptic time, so a
// cudaHostRegister cache would never hit) and expects GrITAS_Norm to fill pageable grids.  One thread's memcpy runs
// at a fraction of PCIe; a handful of threads copying 1 MB chunks keep the DMA engine fed while it already moves
// the chunks that are done.
constexpr size_t COPY_CHUNK = (size_t)2 << 20;
constexpr size_t COPY_CHUNK_SMALL = (size_t)256 << 10;   // the first COPY_SMALL_FIRST chunks of a call (see stage())
constexpr int COPY_SMALL_FIRST = 8;

// Large host copies with non-temporal stores: the destination (pinned staging read next by the DMA engine, or a grid
// far larger than the caches) is not going to be read by this core, so the copy should not pull it into the cache
// first.  Measured on the bench box (tools/hostcopy_bench.c): 60 GB/s with 8 threads against 49 GB/s for memcpy.
#if defined(__x86_64__)
#include <immintrin.h>
__attribute__((target("avx2"))) void copy_stream(void *dst, const void *src, size_t n) {
  char *d = (char *)dst;
  const char *s = (const char *)src;
  const size_t head = ((uintptr_t)d & 31u) ? 32u - ((uintptr_t)d & 31u) : 0u;
  if (n < 4096 + head) { memcpy(d, s, n); return; }
  if (head) { memcpy(d, s, head); d += head; s += head; n -= head; }
  const size_t blocks = n / 128;
  for (size_t i = 0; i < blocks; ++i) {
    const __m256i a = _mm256_loadu_si256((const __m256i *)(s) + 0), b = _mm256_loadu_si256((const __m256i *)(s) + 1),
                  c = _mm256_loadu_si256((const __m256i *)(s) + 2), e = _mm256_loadu_si256((const __m256i *)(s) + 3);
    _mm256_stream_si256((__m256i *)(d) + 0, a);
    _mm256_stream_si256((__m256i *)(d) + 1, b);
    _mm256_stream_si256((__m256i *)(d) + 2, c);
    _mm256_stream_si256((__m256i *)(d) + 3, e);
    s += 128;
    d += 128;
  }
  _mm_sfence();
  if (n % 128) memcpy(d, s, n % 128);
}
bool have_avx2() { static const bool v = __builtin_cpu_supports("avx2"); return v; }
void copy_large(void *dst, const void *src, size_t n) {
  if (have_avx2()) copy_stream(dst, src, n);
  else memcpy(dst, src, n);
}
#else
void copy_large(void *dst, const void *src, size_t n) { memcpy(dst, src, n); }
#endif

class CopyPool {
 public:
  struct Task { void *dst; const void *src; size_t n; std::atomic<int> *done; };
  static CopyPool &get() {
    static CopyPool *pool = new CopyPool();   // lives for the process: no destructor race at exit
    return *pool;
  }
  // all chunks of a call at once: one lock, one wake-up (a futex wake per task costs more than the copy in a VM)
  void submit_batch(const Task *t, size_t count) {
    if (!count) return;
    {
      std::lock_guard<std::mutex> lk(m_);
      for (size_t i = 0; i < count; ++i) q_.push_back(t[i]);
      queued_.fetch_add((int)count, std::memory_order_release);
    }
    cv_.notify_all();
  }
  // one large copy as `parts` tasks: *done starts at 1 - parts and reaches 1 when the last part has landed
  void submit_split(void *dst, const void *src, size_t n, std::atomic<int> *done, size_t part = (size_t)512 << 10) {
    const size_t parts = std::max<size_t>(1, (n + part - 1) / part);
    done->store(1 - (int)parts, std::memory_order_relaxed);
    std::vector<Task> t(parts);
    for (size_t i = 0; i < parts; ++i) t[i] = Task{(char *)dst + i * part, (const char *)src + i * part, std::min(part, n - i * part), done};
    submit_batch(t.data(), parts);
  }
  // the waiting thread copies too instead of spinning
  void wait(std::atomic<int> *done) {
    while (done->load(std::memory_order_acquire) < 1)
      if (!run_one()) std::this_thread::yield();
  }
 private:
  CopyPool() {
    int n = 0;
    if (const char *e = getenv("GRITAS_B200_COPY_THREADS")) n = atoi(e);
    else {
      const unsigned hw = std::thread::hardware_concurrency();
      n = (int)std::min(7u, std::max(1u, hw / 2));         // + the calling thread
    }
    for (int i = 0; i < n; ++i) th_.emplace_back([this] { run(); });
    for (auto &t : th_) t.detach();
  }
  bool pop(Task &t) {
    if (queued_.load(std::memory_order_acquire) <= 0) return false;
    std::lock_guard<std::mutex> lk(m_);
    if (q_.empty()) return false;
    t = q_.front();
    q_.pop_front();
    queued_.fetch_sub(1, std::memory_order_relaxed);
    return true;
  }
  bool run_one() {
    Task t;
    if (!pop(t)) return false;
    copy_large(t.dst, t.src, t.n);
    t.done->fetch_add(1, std::memory_order_release);   // wait() returns when the counter reaches 1 (or above: see submit_split)
    return true;
  }
  void run() {
    for (;;) {
      if (run_one()) continue;
      // nothing queued: poll for a little while (the next call of a file loop follows within microseconds), then sleep
      bool got = false;
      for (int spin = 0; spin < 20000 && !got; ++spin) {
        if (queued_.load(std::memory_order_acquire) > 0) got = true;
        else if ((spin & 63) == 63) std::this_thread::yield();
      }
      if (got) continue;
      std::unique_lock<std::mutex> lk(m_);
      cv_.wait(lk, [this] { return !q_.empty(); });
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<Task> q_;
  std::atomic<int> queued_{0};
};

struct PendingCopy {   // one chunk of a pageable column: copied into the pinned slot buffer by the pool, then DMA'd
  void *dev; const void *pin; size_t n; std::atomic<int> *done;
  const void *src;
  bool submitted;
};

struct Slot {
  std::vector<PendingCopy> pending;
  std::vector<std::atomic<int>> flags;
  size_t nflags = 0;
  void *dev[NCOL] = {};
  void *pin[NCOL] = {};
  size_t dev_cap[NCOL] = {};
  size_t pin_cap[NCOL] = {};
  int *flag_out = nullptr;
  size_t flag_out_cap = 0;
  void *flag_pin = nullptr;        // early ob_flag write-back: pinned landing buffer when the caller's array is pageable
  size_t flag_pin_cap = 0;
  cudaEvent_t ints_copied = nullptr, flags_ready = nullptr;
  std::atomic<int> flag_done{1};
  cudaEvent_t consumed = nullptr;  // kernel that read this slot has finished
  cudaEvent_t copied = nullptr;    // H2D copies into this slot have finished
  bool in_use = false;             // staging buffers of this slot are referenced by a queued kernel
  bool direct_dma = false;         // this call DMAs straight from a pinned array of the caller
};

}  // namespace

struct gritas_b200_ctx {
  int device = 0;
  int mode = 0, flags = 0;
  GridDesc g = {};
  size_t nrec = 0;
  int *d_table = nullptr;
  double *d_p1 = nullptr, *d_p2 = nullptr;
  unsigned short *d_lut = nullptr;
  Status *d_status = nullptr;
  Status *h_status = nullptr;  // pinned
  cudaStream_t compute = nullptr, copy = nullptr;
  cudaStream_t early_st = nullptr;   // early ob_flag write-back (k_early_flags + its D2H), next to the upload of the fp64 columns
  bool own_compute = true;
  Slot slot[NSLOT];
  int next_slot = 0;
  unsigned call_seq = 0;
  // deterministic path: keys + sort of call n+1 run on a sort stream while the in-order segment sums of call n (which
  // must stay serial: they continue the running sums in the records) run on `compute`.  Two buffer sets alternate.
  struct DetBuf {
    unsigned *keys = nullptr, *idx = nullptr, *keys2 = nullptr, *idx2 = nullptr;
    size_t cap = 0;
    void *tmp = nullptr;
    size_t tmp_cap = 0;
    cudaStream_t st = nullptr;
    cudaEvent_t sorted = nullptr, consumed = nullptr;
    bool busy = false;          // `consumed` has been recorded at least once
  } det[2];
  int next_det = 0;
  void *cub_tmp = nullptr;      // pre-sort scratch
  size_t cub_cap = 0;
  int key_bits = 32;
  // tiled fast path: staging lists + flush bookkeeping
  StageDesc sg = {};
  bool staging = false;        // tiled path available (buffers allocated)
  int tiled_choice = -1;       // -1 undecided (probe the first large call), 0 direct, 1 tiled
  bool minmax = false;         // GrITAS_MinMax handle: records {n, min key, max key, pad}
  double *d_mask = nullptr;    // gritas_maskout fraction field (im x jm), optional
  bool acc_fresh = true;       // records untouched since create / reset (apart from direct RED, see StageDesc::dirty)
  size_t stage_slots = 0;      // ntiles * cap
  size_t staged_upper = 0;     // upper bound of entries appended since the last flush
  size_t tf_smem = 0;            // shared memory of the tile kernel (K2f / K2t)
  int tf_ctas = 148, tf_per_sm = 1;
  int k1_per_sm = 2;             // CTAs per SM of one K1 launch (GRITAS_B200_K1_CTAS_PER_SM; its launch bounds keep 4 resident: two launches fill an SM)
  bool fused_norm = true;        // Norm consumes the tile lists itself (K2f) instead of flush (K2t) + K3
  bool det_tiled = false;        // deterministic mode on the tiled path (k_tile_det): entries carry sequence numbers
  size_t td_smem = 0;            // shared memory of k_tile_det
  int td_ctas = 148;
  unsigned long long det_calls = 0;    // Accum calls since the reset
  unsigned *h_ctrl = nullptr;          // pinned [NWK][4]: the control words after the fill / dirty arrays, per worker stream
  bool ctrl_valid[NWK] = {};
  unsigned long long call_order = 0;   // order of the next Accum call (deterministic mode): calls since the reset, or what
  bool call_order_set = false;         // gritas_b200_set_call_order gave for it (multi-rank: the global file index)

  double *red = nullptr;         // multi-rank Norm: compact {n, sums} planes per tile, the operand of ncclReduce
  size_t red_cap = 0;
  // Worker streams: the K1 launch of one Accum call (1.2 M observations, two waves of short CTAs) cannot fill
  // the GPU; consecutive calls run on different streams (own level-miss record) and overlap.  Everything
  // that reads or resets the accumulators runs on `compute` after join_workers().
  cudaStream_t wk[NWK] = {};
  cudaEvent_t wk_done[NWK] = {};
  bool wk_pending[NWK] = {};
  int next_wk = 0;
  int nwk = 6;                    // measured on the cfg2 month: 3 streams x 4 CTAs/SM 3.09 ms, 6 x 2 and 8 x 2 3.03 ms, 2 x 4 3.43 ms (profiles/r2b_k1_k2f_experiments.txt)
  cudaEvent_t main_ev = nullptr;  // last accumulator-wide operation queued on `compute`
  // reduce + finalize pipeline (multi-GPU): chunk c is tile-accumulated on `compute`, reduced to the root on
  // `comm_stream`, finalized on `fin_stream` while later chunks are still being accumulated / reduced
  static constexpr int NCHUNK = 16;   // capacity; the pipeline uses GRITAS_B200_PIPE_CHUNKS of them (default 8)
  cudaStream_t comm_stream = nullptr, fin_stream = nullptr;
  cudaEvent_t ev_acc[NCHUNK] = {}, ev_red[NCHUNK] = {}, ev_fin = nullptr;
  // pre-sort scratch (gritas_b200_presort)
  void *ps_cols = nullptr, *ps_key[2] = {}, *ps_idx[2] = {}, *ps_sorted = nullptr;
  size_t ps_cols_cap = 0, ps_key_cap[2] = {}, ps_idx_cap[2] = {}, ps_sorted_cap = 0;
  // finalize scratch (reference-layout planes on the device)
  double *out = nullptr;
  size_t out_cap = 0;
  // multi-GPU
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  // sharded Norm over peer memory (gritas_b200_peer_attach): every rank's lists / fill words / records
  int np = 1, prank = 0;
  void *peer_lists[MAX_PEERS] = {}, *peer_fill[MAX_PEERS] = {}, *peer_acc[MAX_PEERS] = {}, *peer_red[MAX_PEERS] = {};
  bool peer_ipc[MAX_PEERS] = {};      // pointers of rank s came from cudaIpcOpenMemHandle (close them on destroy)
  unsigned *fill_snap = nullptr;      // local snapshot of every rank's fill / dirty / records-touched words
  size_t fill_snap_cap = 0;
  unsigned long long *slab_prefix = nullptr;   // scratch of k_slab_bounds
  unsigned *d_bounds = nullptr, *h_bounds = nullptr;   // slab boundaries of the last sharded Norm (device / pinned)
  cudaEvent_t bounds_ev = nullptr;
  bool bounds_valid = false;
  bool export_valid = false;          // h->red holds the planes of the current accumulation state (gritas_b200_export_planes)
  double *bar_buf = nullptr;          // operand of the stream-ordered barrier (a one-element all-reduce)
  // device -> pageable host: ring of pinned bounce buffers (copy_out_pageable)
  static constexpr int NBOUNCE = 8;
  static constexpr size_t BOUNCE_BYTES = (size_t)4 << 20;
  void *bounce[NBOUNCE] = {};
  cudaEvent_t bounce_ev[NBOUNCE] = {};
  std::atomic<int> bounce_done[NBOUNCE];
  // counters
  uint64_t launches = 0, h2d = 0, d2h = 0;
  char err[512] = {0};
};

namespace {

int fail(gritas_b200_handle h, int rc, const char *fmt, ...) {
  if (h) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(h->err, sizeof(h->err), fmt, ap);
    va_end(ap);
  }
  return rc;
}

#define CU(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess)                                                                           \
      return fail(h, GRITAS_B200_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

// Where does a caller pointer live?  0 pageable host, 1 pinned host, 2 device / managed.
// cudaPointerGetAttributes costs about a microsecond and an Accum call passes nine columns, so the
// answers are kept in a small direct-mapped cache (drivers reuse their column buffers).  Device and
// host addresses never alias under UVA; a pageable/pinned mix-up after a free + realloc at the same
// address only changes how the copy is staged, never its result.
int mem_kind_query(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return 2;
  if (a.type == cudaMemoryTypeHost) return 1;
  return 0;
}

struct KindCache {
  static constexpr int N = 4096;
  const void *ptr[N] = {};
  signed char kind[N] = {};
};
thread_local KindCache g_kinds;

int mem_kind(const void *p) {
  if (!p) return 0;
  const size_t slot = (((uintptr_t)p >> 4) * 0x9E3779B97F4A7C15ull >> 40) % KindCache::N;
  if (g_kinds.ptr[slot] == p) return g_kinds.kind[slot];
  const int k = mem_kind_query(p);
  g_kinds.ptr[slot] = p;
  g_kinds.kind[slot] = (signed char)k;
  return k;
}

int grow_dev(gritas_b200_handle h, void **p, size_t *cap, size_t bytes) {
  if (*cap >= bytes) return 0;
  if (*p) CU(cudaFree(*p));
  *p = nullptr;
  *cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  CU(cudaMalloc(p, want));
  *cap = want;
  return 0;
}

int grow_pin(gritas_b200_handle h, void **p, size_t *cap, size_t bytes) {
  if (*cap >= bytes) return 0;
  if (*p) CU(cudaFreeHost(*p));
  *p = nullptr;
  *cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  CU(cudaMallocHost(p, want));
  *cap = want;
  return 0;
}

int flush_pending(gritas_b200_handle h, Slot &s, size_t hook_at = (size_t)-1, const std::function<int()> *hook = nullptr);

// Makes column `c` of this call available on the device; returns the device pointer in *out.
// Device-resident inputs are used in place.  Pinned inputs are DMA'd straight from the caller's
// array; pageable inputs go through the slot's pinned buffer.
int stage(gritas_b200_handle h, Slot &s, int c, const void *src, size_t bytes, const void **out, bool *any_copy) {
  // (s.direct_dma: a copy of this call reads the caller's own pinned array -- an ASYNC call then has to wait for the DMA
  // before it returns; pageable arrays are consumed once the pool has copied them into the slot's pinned buffer)
  if (!src || bytes == 0) { *out = src ? src : nullptr; return 0; }
  const int kind = mem_kind(src);
  if (kind == 2) { *out = src; return 0; }
  if (s.in_use) {  // first overwrite of this slot in this call: the kernel that last read it must be done
    CU(cudaEventSynchronize(s.consumed));
    s.in_use = false;
  }
  int rc = grow_dev(h, &s.dev[c], &s.dev_cap[c], bytes);
  if (rc) return rc;
  if (kind == 0) {
    // pageable: the pool copies chunk by chunk into the slot's pinned buffer; flush_pending() issues each chunk's DMA as
    // soon as it has landed there (the other chunks are still being copied)
    rc = grow_pin(h, &s.pin[c], &s.pin_cap[c], bytes);
    if (rc) return rc;
    // the first chunks of a call are small: the pool's threads all start at once, and the DMA engine idles until the
    // FIRST chunk has landed in pinned memory (2 MB at one thread's share of the copy rate is ~0.3 ms of a 1.2 ms upload)
    for (size_t off = 0, n = 0; off < bytes; off += n) {
      n = std::min(s.pending.size() < (size_t)COPY_SMALL_FIRST ? COPY_CHUNK_SMALL : COPY_CHUNK, bytes - off);
      if (s.nflags >= s.flags.size() && (rc = flush_pending(h, s))) return rc;   // (a very large call: DMA what is there, reuse the flags)
      std::atomic<int> *done = &s.flags[s.nflags++];
      done->store(0, std::memory_order_relaxed);
      s.pending.push_back(PendingCopy{(char *)s.dev[c] + off, (const char *)s.pin[c] + off, n, done, (const char *)src + off, false});
    }
  } else {
    CU(cudaMemcpyAsync(s.dev[c], src, bytes, cudaMemcpyHostToDevice, h->copy));
    s.direct_dma = true;
  }
  h->h2d += bytes;
  *any_copy = true;
  *out = s.dev[c];
  return 0;
}

// DMA of the pageable chunks in the order they were submitted (each as soon as the pool has copied it).
// `hook` runs once the DMAs of the first `hook_at` chunks are queued (the early ob_flag write-back: after the integer columns).
int flush_pending(gritas_b200_handle h, Slot &s, size_t hook_at, const std::function<int()> *hook) {
  {   // hand every chunk of the call to the pool in one go
    std::vector<CopyPool::Task> batch;
    batch.reserve(s.pending.size());
    for (PendingCopy &pc : s.pending)
      if (!pc.submitted) {
        batch.push_back(CopyPool::Task{const_cast<void *>(pc.pin), pc.src, pc.n, pc.done});
        pc.submitted = true;
      }
    CopyPool::get().submit_batch(batch.data(), batch.size());
  }
  size_t issued = 0;
  for (const PendingCopy &pc : s.pending) {
    if (hook && issued == hook_at) { if (int hr = (*hook)()) return hr; }
    CopyPool::get().wait(pc.done);
    CU(cudaMemcpyAsync(pc.dev, pc.pin, pc.n, cudaMemcpyHostToDevice, h->copy));
    ++issued;
  }
  if (hook && issued <= hook_at) { if (int hr = (*hook)()) return hr; }
  s.pending.clear();
  s.nflags = 0;
  return 0;
}

// Device -> pageable host array: DMA into a ring of pinned bounce buffers, the pool copies them out while the next
// chunks are in flight (cudaMemcpyAsync to pageable memory would stage through the driver on one thread).
// Synchronous: the data is in `user` on return.  Stream order is kept (the DMAs are queued on `sm`).
int copy_out_pageable(gritas_b200_handle h, void *user, const void *dev, size_t bytes, cudaStream_t sm) {
  constexpr int NB = gritas_b200_ctx::NBOUNCE, LAG = 2;
  constexpr size_t CB = gritas_b200_ctx::BOUNCE_BYTES;
  if (!h->bounce[0]) {
    for (int b = 0; b < NB; ++b) {
      CU(cudaMallocHost(&h->bounce[b], CB));
      CU(cudaEventCreateWithFlags(&h->bounce_ev[b], cudaEventDisableTiming));
      h->bounce_done[b].store(1);
    }
  }
  const size_t nchunk = (bytes + CB - 1) / CB;
  auto retire = [&](size_t j) -> int {      // chunk j has landed in its bounce buffer: hand it to the pool
    const int b = (int)(j % NB);
    CU(cudaEventSynchronize(h->bounce_ev[b]));
    const size_t off = j * CB, n = std::min(CB, bytes - off);
    CopyPool::get().submit_split((char *)user + off, h->bounce[b], n, &h->bounce_done[b]);
    return 0;
  };
  for (size_t i = 0; i < nchunk; ++i) {
    const int b = (int)(i % NB);
    CopyPool::get().wait(&h->bounce_done[b]);      // the copy out of chunk i - NB is complete: the buffer is free
    const size_t off = i * CB, n = std::min(CB, bytes - off);
    CU(cudaMemcpyAsync(h->bounce[b], (const char *)dev + off, n, cudaMemcpyDeviceToHost, sm));
    CU(cudaEventRecord(h->bounce_ev[b], sm));
    if (i >= (size_t)LAG) { int rc = retire(i - LAG); if (rc) return rc; }
  }
  for (size_t j = nchunk > (size_t)LAG ? nchunk - LAG : 0; j < nchunk; ++j) { int rc = retire(j); if (rc) return rc; }
  for (int b = 0; b < NB; ++b) CopyPool::get().wait(&h->bounce_done[b]);
  h->d2h += bytes;
  return 0;
}

bool aligned16(const void *p) { return (((uintptr_t)p) & 15u) == 0; }

// GRITAS_B200_TRACE=1: host-clock phase times of every Accum call on stderr (where does a synchronous call spend its time?)
struct CallTrace {
  bool on;
  double t0, last;
  char buf[256];
  int len = 0;
  static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
  CallTrace() {
    static const bool v = getenv("GRITAS_B200_TRACE") && atoi(getenv("GRITAS_B200_TRACE")) != 0;
    on = v;
    buf[0] = 0;
    t0 = last = on ? now() : 0.0;
  }
  void mark(const char *what) {
    if (!on || len >= (int)sizeof(buf) - 32) return;
    const double t = now();
    len += snprintf(buf + len, sizeof(buf) - (size_t)len, " %s %.3f", what, t - last);
    last = t;
  }
  ~CallTrace() {
    if (on) fprintf(stderr, "[gritas_b200 trace] total %.3f ms:%s\n", now() - t0, buf);
  }
};
thread_local CallTrace *g_trace = nullptr;
inline void trace_mark(const char *what) { if (g_trace) g_trace->mark(what); }

int clear_status(gritas_b200_handle h) {
  Status st[NWK + 1];
  for (Status &q : st) { q.first_bad = BAD_NONE; q.bad_lev = 0.0; q.near_pairs = 0; q.pairs = 0; }
  CU(cudaMemcpyAsync(h->d_status, st, sizeof(st), cudaMemcpyHostToDevice, h->compute));
  CU(cudaStreamSynchronize(h->compute));
  return 0;
}

// `compute` waits for every K1 launch queued on the worker streams (no host synchronisation).
int join_workers(gritas_b200_handle h) {
  for (int w = 0; w < NWK; ++w)
    if (h->wk_pending[w]) {
      CU(cudaStreamWaitEvent(h->compute, h->wk_done[w], 0));
      h->wk_pending[w] = false;
    }
  return 0;
}

// Marks an accumulator-wide operation on `compute` (reset, flush, add_raw): later K1 launches on the
// worker streams must not start before it.
int mark_main(gritas_b200_handle h) {
  if (h->main_ev) CU(cudaEventRecord(h->main_ev, h->compute));
  return 0;
}

// The records are about to receive sums (flush, direct accumulate, add_raw, reduction): they are no longer "all zero
// apart from dirty tiles".  The device copy of that fact (the word after the dirty words) is what a peer's sharded Norm reads.
int touch_records(gritas_b200_handle h) {
  if (!h->acc_fresh) return 0;
  h->acc_fresh = false;
  if (h->staging) CU(cudaMemsetAsync(h->sg.fill + 2 * (size_t)h->sg.ntiles, 1, sizeof(unsigned), h->compute));
  return 0;
}

// Waits for the compute stream and converts a recorded level miss into the return code.
int drain(gritas_b200_handle h, double *bad_lev) {
  int jr = join_workers(h);
  if (jr) return jr;
  CU(cudaMemcpyAsync(h->h_status, h->d_status, sizeof(Status) * (NWK + 1), cudaMemcpyDeviceToHost, h->compute));
  if (h->det_tiled)
    CU(cudaMemcpyAsync(h->h_ctrl + 4 * NWK, h->sg.fill + 2 * (size_t)h->sg.ntiles, 4 * sizeof(unsigned), cudaMemcpyDeviceToHost, h->compute));
  CU(cudaStreamSynchronize(h->compute));
  if (h->det_tiled && h->h_ctrl[4 * NWK + 1] != 0u) {
    const unsigned why = h->h_ctrl[4 * NWK + 1];
    CU(cudaMemsetAsync(h->sg.fill + 2 * (size_t)h->sg.ntiles + 1, 0, sizeof(unsigned), h->compute));
    CU(cudaStreamSynchronize(h->compute));
    return fail(h, GRITAS_B200_ERR_ORDER,
                why == 2u ? "deterministic mode: one box holds more parked observations than a chunk of the tile kernel; its sums were "
                            "formed out of order (complete, not reproducible).  Use GRITAS_B200_FLAG_DET_SORT for such data"
                          : "deterministic mode: a tile list overflowed; the observations that did not fit were added out of order "
                            "(complete, not reproducible).  Raise GRITAS_B200_STAGE_SLOTS or use GRITAS_B200_FLAG_DET_SORT");
  }
  int best = 0;  // earliest (call, observation) over the streams
  for (int q = 1; q <= NWK; ++q)
    if (h->h_status[q].first_bad < h->h_status[best].first_bad) best = q;
  if (h->h_status[best].first_bad != BAD_NONE) {
    const double lev = h->h_status[best].bad_lev;
    const unsigned long long fb = h->h_status[best].first_bad;
    if (bad_lev) *bad_lev = lev;
    int rc = clear_status(h);
    if (rc) return rc;
    return fail(h, GRITAS_B200_ERR_LEVEL, "Accum: obs level is %.17g hPa; Accum: could not find level (call %u, obs %u)",
                lev, (unsigned)(fb >> 32), (unsigned)fb + 1u);
  }
  return 0;
}

int grid_for(size_t work, int threads, int per_sm) {
  size_t b = (work + threads - 1) / threads;
  size_t cap = (size_t)148 * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

inline int filter_mode(const FilterDesc &f) { return !f.ods ? 0 : ((f.use_time || f.use_lon) ? 2 : 1); }

template <int NR, int FMODE>
void launch_fast_m(gritas_b200_handle h, const ObsCols &c, const FilterDesc &f, int nblk, size_t smem, bool probe) {
  if (probe)
    k_accum_fast<NR, true, FMODE, true><<<nblk, ACC_THREADS, smem, h->compute>>>(h->g, c, f, h->d_status, h->call_seq);
  else if (h->flags & GRITAS_B200_FLAG_NO_AGGREGATE)
    k_accum_fast<NR, false, FMODE><<<nblk, ACC_THREADS, smem, h->compute>>>(h->g, c, f, h->d_status, h->call_seq);
  else
    k_accum_fast<NR, true, FMODE><<<nblk, ACC_THREADS, smem, h->compute>>>(h->g, c, f, h->d_status, h->call_seq);
}

template <int NR>
void launch_fast(gritas_b200_handle h, const ObsCols &c, const FilterDesc &f, int nblk, size_t smem, bool probe) {
  switch (filter_mode(f)) {
    case 0: launch_fast_m<NR, 0>(h, c, f, nblk, smem, probe); break;
    case 1: launch_fast_m<NR, 1>(h, c, f, nblk, smem, probe); break;
    default: launch_fast_m<NR, 2>(h, c, f, nblk, smem, probe); break;
  }
}

// Tiled fast path, K1 of one Accum call on worker stream w: every accepted observation -> one entry in its tile's list.
template <int NR>
void launch_lists(gritas_b200_handle h, int w, const ObsCols &c, const FilterDesc &f) {
  const int nvec = (c.n + OBS_PER_THREAD - 1) / OBS_PER_THREAD;
  // one resident wave (4 CTAs per SM, launch bounds of k_list_obs), grid-stride over the file
  const int g1 = grid_for((size_t)nvec, K1_THREADS, h->k1_per_sm);
  const size_t smem = levtab_smem_bytes(h->g);
  Status *st = h->d_status + 1 + w;
  if (h->det_tiled) {
    const unsigned long long sb = h->call_order << SEQ_IDX_BITS;
    switch (filter_mode(f)) {
      case 0: k_list_obs<NR, 0, true><<<g1, K1_THREADS, smem, h->wk[w]>>>(h->g, c, f, st, h->call_seq, h->sg, sb); break;
      case 1: k_list_obs<NR, 1, true><<<g1, K1_THREADS, smem, h->wk[w]>>>(h->g, c, f, st, h->call_seq, h->sg, sb); break;
      default: k_list_obs<NR, 2, true><<<g1, K1_THREADS, smem, h->wk[w]>>>(h->g, c, f, st, h->call_seq, h->sg, sb); break;
    }
    return;
  }
  switch (filter_mode(f)) {
    case 0: k_list_obs<NR, 0><<<g1, K1_THREADS, smem, h->wk[w]>>>(h->g, c, f, st, h->call_seq, h->sg); break;
    case 1: k_list_obs<NR, 1><<<g1, K1_THREADS, smem, h->wk[w]>>>(h->g, c, f, st, h->call_seq, h->sg); break;
    default: k_list_obs<NR, 2><<<g1, K1_THREADS, smem, h->wk[w]>>>(h->g, c, f, st, h->call_seq, h->sg); break;
  }
}

// K2d launches (deterministic tiled path).
template <int NR, int MODE>
void launch_det(gritas_b200_handle h, cudaStream_t sm, unsigned t0, unsigned t1, const OutPlanes &o2, const OutPlanes &o3, int nrmlz, int ntr,
                int raw2, int raw3, const PeerSet &ps) {
  if (t1 <= t0) return;
  const int ctas = (int)std::min((unsigned)h->td_ctas, t1 - t0);
  k_tile_det<NR, MODE><<<ctas, TileDet<NR>::THREADS, h->td_smem, sm>>>(h->g, h->sg, (unsigned)h->nrec, o2, o3, nrmlz, ntr, raw2, raw3,
                                                                      h->acc_fresh ? 1 : 0, t0, t1, ps);
}
template <int MODE>
void launch_det_nr(gritas_b200_handle h, cudaStream_t sm, unsigned t0, unsigned t1, const OutPlanes &o2, const OutPlanes &o3, int nrmlz, int ntr,
                   int raw2, int raw3, const PeerSet &ps) {
  switch (h->g.nr) {
    case 1: launch_det<1, MODE>(h, sm, t0, t1, o2, o3, nrmlz, ntr, raw2, raw3, ps); break;
    case 2: launch_det<2, MODE>(h, sm, t0, t1, o2, o3, nrmlz, ntr, raw2, raw3, ps); break;
    default: launch_det<3, MODE>(h, sm, t0, t1, o2, o3, nrmlz, ntr, raw2, raw3, ps); break;
  }
}

// K2t over tiles [t0, t1): the lists are added to the records and emptied.
template <int NR>
void launch_tiles(gritas_b200_handle h, unsigned t0, unsigned t1, int free_sms) {
  if (t1 <= t0) return;
  const OutPlanes none = {nullptr, nullptr, nullptr, nullptr, nullptr, 0};
  // the grid is persistent (every CTA resident): when a collective runs next to it, leave SMs for its channels
  const int avail = std::max(h->tf_per_sm, h->tf_ctas - free_sms * h->tf_per_sm);
  const int ctas = (int)std::min((unsigned)avail, t1 - t0);
  k_tile_final<NR, TILE_FLUSH><<<ctas, TileFinal<NR>::THREADS, h->tf_smem, h->compute>>>(h->g, h->sg, (unsigned)h->nrec, none, none, 0, 0,
                                                                                         0, 0, h->acc_fresh ? 1 : 0, t0, t1, nullptr, PeerSet{});
}

void launch_tiles_nr(gritas_b200_handle h, unsigned t0, unsigned t1, int free_sms = 0) {
  switch (h->g.nr) {
    case 1: launch_tiles<1>(h, t0, t1, free_sms); break;
    case 2: launch_tiles<2>(h, t0, t1, free_sms); break;
    default: launch_tiles<3>(h, t0, t1, free_sms); break;
  }
}

int flush_stage(gritas_b200_handle h, unsigned min_fill = 0) {
  if (!h->staging || h->staged_upper == 0) return 0;
  int jr = join_workers(h);
  if (jr) return jr;
  if (h->det_tiled) {
    const OutPlanes none = {nullptr, nullptr, nullptr, nullptr, nullptr, 0};
    launch_det_nr<DET_FLUSH>(h, h->compute, 0u, h->sg.ntiles, none, none, 0, 0, (int)min_fill, 0, PeerSet{});
    if (min_fill > 0) {
      // only the long lists (hot tiles of clustered data) went to their records; the tile kernel marked those tiles dirty,
      // everything else is still parked and the records of the other tiles are still zero
      CU(cudaGetLastError());
      h->launches += 1;
      return mark_main(h);
    }
  } else {
    launch_tiles_nr(h, 0u, h->sg.ntiles);
  }
  CU(cudaGetLastError());
  h->launches += 1;
  h->staged_upper = 0;
  if ((jr = touch_records(h))) return jr;
  return mark_main(h);
}

// K2f: Norm / get_raw straight from the tile lists (+ the records if anything was flushed into them).
template <int NR>
void launch_tile_final(gritas_b200_handle h, cudaStream_t sm, const OutPlanes &o2, const OutPlanes &o3, int nrmlz, int ntr,
                       int raw2, int raw3) {
  k_tile_final<NR, TILE_FINALIZE><<<h->tf_ctas, TileFinal<NR>::THREADS, h->tf_smem, sm>>>(h->g, h->sg, (unsigned)h->nrec, o2, o3,
                                                                                         nrmlz, ntr, raw2, raw3, h->acc_fresh ? 1 : 0,
                                                                                         0u, h->sg.ntiles, nullptr, PeerSet{});
}

// Multi-rank Norm on the tiled path: K2x exports the sums of tiles [t0, t1) as compact planes (the reduction operand),
// K3x finalizes them once they are reduced.
template <int NR>
void launch_tile_export(gritas_b200_handle h, cudaStream_t sm, unsigned t0, unsigned t1) {
  if (t1 <= t0) return;
  const OutPlanes none = {nullptr, nullptr, nullptr, nullptr, nullptr, 0};
  const int ctas = (int)std::min((unsigned)h->tf_ctas, t1 - t0);
  k_tile_final<NR, TILE_EXPORT><<<ctas, TileFinal<NR>::THREADS, h->tf_smem, sm>>>(h->g, h->sg, (unsigned)h->nrec, none, none, 0, 0, 0, 0,
                                                                                   h->acc_fresh ? 1 : 0, t0, t1, h->red, PeerSet{});
}
template <int NR>
void launch_tile_import(gritas_b200_handle h, cudaStream_t sm, const OutPlanes &o2, const OutPlanes &o3, int nrmlz, int ntr,
                        int raw2, int raw3, unsigned t0, unsigned t1) {
  if (t1 <= t0) return;
  const int ctas = (int)std::min((unsigned)h->tf_ctas, t1 - t0);
  k_tile_final<NR, TILE_IMPORT><<<ctas, TileFinal<NR>::THREADS, h->tf_smem, sm>>>(h->g, h->sg, (unsigned)h->nrec, o2, o3, nrmlz, ntr,
                                                                                   raw2, raw3, 1, t0, t1, h->red, PeerSet{});
}

bool fused_norm_ok(gritas_b200_handle h) {
  return h->staging && h->fused_norm && h->staged_upper > 0 && !h->minmax;
}

// Host side of every reader of the accumulators: copies done, lists flushed unless the reader consumes them itself
// (K2f), level misses reported.
int settle(gritas_b200_handle h, bool keep_lists, double *bad_lev) {
  CU(cudaStreamSynchronize(h->copy));
  if (!keep_lists) {
    int rc = flush_stage(h);
    if (rc) return rc;
  }
  return drain(h, bad_lev);
}

template <int NR>
void launch_segsum(gritas_b200_handle h, const ObsCols &c, const FilterDesc &f, const unsigned *keys,
                   const unsigned *idx, int n) {
  k_segment_sum<NR><<<(n + ACC_THREADS - 1) / ACC_THREADS, ACC_THREADS, 0, h->compute>>>(h->g, c, f, keys, idx, n,
                                                                                         (unsigned)h->nrec);
}

// The common back half of accum / accum_ods once the column pointers are on the device.
int run_accum(gritas_b200_handle h, ObsCols c, const FilterDesc &f, Slot &s, bool any_copy, int32_t *flag_host_or_dev,
              double *bad_lev) {
  const int n = c.n;
  // ob_flag write-back target
  bool flag_to_host = false;
  if (flag_host_or_dev) {
    if (mem_kind(flag_host_or_dev) == 2) {
      c.flag_out = flag_host_or_dev;
    } else {
      int rc = grow_dev(h, (void **)&s.flag_out, &s.flag_out_cap, sizeof(int) * (size_t)n);
      if (rc) return rc;
      c.flag_out = s.flag_out;
      flag_to_host = true;
    }
  }
  c.flag_rw = (c.flag && c.flag == c.flag_out) ? 1 : 0;   // device-resident inout ob_flag: read with coherent loads
  c.aligned = aligned16(c.lat) && aligned16(c.lon) && aligned16(c.lev) && aligned16(c.kx) && aligned16(c.kt) &&
              aligned16(c.flag) && aligned16(c.time) && aligned16(c.flag_out);
  for (int ir = 0; ir < h->g.nr; ++ir) c.aligned = c.aligned && aligned16(c.d[ir]);

  // Fast mode picks between the tiled and the direct accumulate by looking at the data: the first call with at
  // least 64 Ki observations is classified by a locality probe first.
  if ((h->mode & 0xff) == GRITAS_B200_MODE_FAST && h->staging && h->tiled_choice < 0 && !h->minmax && n >= 65536) {
    // the probe looks at the first 256 Ki observations of this call (classification only, nothing is accumulated) ...
    if (any_copy) {
      if (int pr = flush_pending(h, s)) return pr;
      CU(cudaEventRecord(s.copied, h->copy));
      CU(cudaStreamWaitEvent(h->compute, s.copied, 0));
    }
    ObsCols cp = c;
    cp.n = std::min(n, 1 << 18);
    cp.flag_out = nullptr;
    const int nvec = (cp.n + OBS_PER_THREAD - 1) / OBS_PER_THREAD;
    const int nblk = grid_for((size_t)nvec, ACC_THREADS, 8);
    const size_t psmem = levtab_smem_bytes(h->g);
    switch (h->g.nr) {
      case 1: launch_fast<1>(h, cp, f, nblk, psmem, true); break;
      case 2: launch_fast<2>(h, cp, f, nblk, psmem, true); break;
      default: launch_fast<3>(h, cp, f, nblk, psmem, true); break;
    }
    CU(cudaGetLastError());
    h->launches += 1;
    CU(cudaMemcpyAsync(h->h_status, h->d_status, sizeof(Status), cudaMemcpyDeviceToHost, h->compute));
    CU(cudaStreamSynchronize(h->compute));
    const unsigned pairs = h->h_status[0].pairs, nearp = h->h_status[0].near_pairs;
    h->tiled_choice = (pairs > 0 && (double)nearp < 0.5 * (double)pairs) ? 1 : 0;
    // ... and the whole call then runs on the path it picked
  }
  const bool probe = false;
  const bool tiled = h->staging && (h->det_tiled || ((h->mode & 0xff) == GRITAS_B200_MODE_FAST && h->tiled_choice == 1));
  if (h->det_tiled) {
    if (!h->call_order_set) h->call_order = h->det_calls;
    h->call_order_set = false;
    h->det_calls += 1;
    if (h->call_order >= (1ull << 20) || (unsigned long long)n >= (1ull << SEQ_IDX_BITS))
      return fail(h, GRITAS_B200_ERR_ARG, "deterministic mode: at most 2^20 calls between resets and 2^28 observations per call");
    // a tile list about to fill up (clustered data)?  flush in time: what does not fit a list is added out of order
    unsigned seen = 0;
    for (int q = 0; q < h->nwk; ++q)
      if (h->ctrl_valid[q] && cudaEventQuery(h->wk_done[q]) == cudaSuccess) { seen = std::max(seen, h->h_ctrl[4 * q + 2]); h->ctrl_valid[q] = false; }
    cudaGetLastError();
    if (seen > h->sg.cap / 2) {
      int rc = flush_stage(h, h->sg.cap / 4);
      if (rc) return rc;
      CU(cudaMemsetAsync(h->sg.fill + 2 * (size_t)h->sg.ntiles + 2, 0, sizeof(unsigned), h->compute));
      if ((rc = mark_main(h))) return rc;
    }
  }
  const int w = h->next_wk;
  cudaStream_t run = tiled ? h->wk[w] : h->compute;
  if (tiled) {
    h->next_wk = (h->next_wk + 1) % h->nwk;
    CU(cudaStreamWaitEvent(run, h->main_ev, 0));   // not before the last reset / flush
  }
  if (any_copy) {
    if (int pr = flush_pending(h, s)) return pr;
    CU(cudaEventRecord(s.copied, h->copy));
    CU(cudaStreamWaitEvent(run, s.copied, 0));
    trace_mark("flush2");
  }
  const size_t smem = levtab_smem_bytes(h->g);
  if (h->minmax) {
    const int nvec = (n + OBS_PER_THREAD - 1) / OBS_PER_THREAD;
    const int nblk = grid_for((size_t)nvec, ACC_THREADS, 8);
    switch (filter_mode(f)) {
      case 0: k_minmax<0><<<nblk, ACC_THREADS, smem, h->compute>>>(h->g, c, f, h->d_status, h->call_seq); break;
      case 1: k_minmax<1><<<nblk, ACC_THREADS, smem, h->compute>>>(h->g, c, f, h->d_status, h->call_seq); break;
      default: k_minmax<2><<<nblk, ACC_THREADS, smem, h->compute>>>(h->g, c, f, h->d_status, h->call_seq); break;
    }
    h->launches += 1;
  } else if (tiled) {
    // tiled path: append this call's accepted observations to the per-tile lists.  Keep the expected fill at or
    // below ~half the capacity; what still overflows (skewed data) is accumulated directly by K1
    if (h->staged_upper + (size_t)n > h->stage_slots / 2) {
      int rc = flush_stage(h);
      if (rc) return rc;
      CU(cudaStreamWaitEvent(run, h->main_ev, 0));
    }
    h->staged_upper += (size_t)n;
    switch (h->g.nr) {
      case 1: launch_lists<1>(h, w, c, f); break;
      case 2: launch_lists<2>(h, w, c, f); break;
      default: launch_lists<3>(h, w, c, f); break;
    }
    h->launches += 1;
  } else if ((h->mode & 0xff) == GRITAS_B200_MODE_FAST) {
    if (int tr = touch_records(h)) return tr;
    const int nvec = (n + OBS_PER_THREAD - 1) / OBS_PER_THREAD;
    const int nblk = grid_for((size_t)nvec, ACC_THREADS, 8);
    switch (h->g.nr) {
      case 1: launch_fast<1>(h, c, f, nblk, smem, probe); break;
      case 2: launch_fast<2>(h, c, f, nblk, smem, probe); break;
      default: launch_fast<3>(h, c, f, nblk, smem, probe); break;
    }
    h->launches += 1;
  } else {
    if (int tr = touch_records(h)) return tr;
    // keys -> stable LSD radix sort of (key, obs index) on a sort stream -> in-order segment sums on `compute`
    gritas_b200_ctx::DetBuf &db = h->det[h->next_det];
    h->next_det ^= 1;
    if (!db.st) {
      CU(cudaStreamCreateWithFlags(&db.st, cudaStreamNonBlocking));
      CU(cudaEventCreateWithFlags(&db.sorted, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&db.consumed, cudaEventDisableTiming));
    }
    if (db.busy) CU(cudaStreamWaitEvent(db.st, db.consumed, 0));   // the segment sums that read this buffer set are done
    if (db.cap < (size_t)n) {
      if (db.busy) CU(cudaEventSynchronize(db.consumed));
      for (unsigned **p : {&db.keys, &db.idx, &db.keys2, &db.idx2}) {
        if (*p) CU(cudaFree(*p));
        *p = nullptr;
      }
      size_t want = (size_t)n + (size_t)n / 8 + 1024;
      CU(cudaMalloc(&db.keys, want * 4));
      CU(cudaMalloc(&db.idx, want * 4));
      CU(cudaMalloc(&db.keys2, want * 4));
      CU(cudaMalloc(&db.idx2, want * 4));
      db.cap = want;
    }
    size_t need = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, need, db.keys, db.keys2, db.idx, db.idx2, n, 0, h->key_bits, db.st);
    if (db.tmp_cap < need) {
      if (db.busy) CU(cudaEventSynchronize(db.consumed));
      if (grow_dev(h, &db.tmp, &db.tmp_cap, need)) return GRITAS_B200_ERR_CUDA;
    }
    if (any_copy) CU(cudaStreamWaitEvent(db.st, s.copied, 0));
    k_box_keys<<<grid_for((size_t)n, ACC_THREADS, 8), ACC_THREADS, smem, db.st>>>(
        h->g, c, f, h->d_status, h->call_seq, db.keys, db.idx, (unsigned)h->nrec);
    size_t tmp = db.tmp_cap;
    CU(cub::DeviceRadixSort::SortPairs(db.tmp, tmp, db.keys, db.keys2, db.idx, db.idx2, n, 0, h->key_bits, db.st));
    CU(cudaEventRecord(db.sorted, db.st));
    CU(cudaStreamWaitEvent(h->compute, db.sorted, 0));
    switch (h->g.nr) {
      case 1: launch_segsum<1>(h, c, f, db.keys2, db.idx2, n); break;
      case 2: launch_segsum<2>(h, c, f, db.keys2, db.idx2, n); break;
      default: launch_segsum<3>(h, c, f, db.keys2, db.idx2, n); break;
    }
    CU(cudaEventRecord(db.consumed, h->compute));
    db.busy = true;
    h->launches += 2 + (uint64_t)((h->key_bits + 7) / 8) * 2 + 1;  // keys, segsum, sort passes (approx.)
  }
  CU(cudaGetLastError());
  if (any_copy || flag_to_host) {
    CU(cudaEventRecord(s.consumed, run));
    s.in_use = true;
  }

  if (flag_to_host) {
    if (mem_kind(flag_host_or_dev) == 0 && (size_t)n * sizeof(int) >= ((size_t)1 << 20)) {
      if (int cr = copy_out_pageable(h, flag_host_or_dev, s.flag_out, sizeof(int) * (size_t)n, run)) return cr;
    } else {
      CU(cudaMemcpyAsync(flag_host_or_dev, s.flag_out, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, run));
      h->d2h += sizeof(int) * (size_t)n;
    }
  }
  if (tiled) {
    if (h->det_tiled) {   // the fullest list so far, for the flush decision of a later call
      CU(cudaMemcpyAsync(h->h_ctrl + 4 * w, h->sg.fill + 2 * (size_t)h->sg.ntiles, 4 * sizeof(unsigned), cudaMemcpyDeviceToHost, run));
      h->ctrl_valid[w] = true;
    }
    CU(cudaEventRecord(h->wk_done[w], run));
    h->wk_pending[w] = true;
  }
  const bool async = (h->flags & GRITAS_B200_FLAG_ASYNC) != 0;
  trace_mark("launch");
  if (!async || flag_to_host) {
    const int dr = drain(h, bad_lev);
    trace_mark("drain");
    return dr;
  }
  // async: the caller's arrays must be reusable on return -> wait for the copies, not the kernel
  if (any_copy && s.direct_dma && !(h->flags & GRITAS_B200_FLAG_HOLD_INPUTS)) CU(cudaEventSynchronize(s.copied));
  return 0;
}

int check_handle(gritas_b200_handle h) {
  if (!h) return GRITAS_B200_ERR_ARG;
  if (cudaSetDevice(h->device) != cudaSuccess) return fail(h, GRITAS_B200_ERR_CUDA, "cudaSetDevice(%d) failed", h->device);
  return 0;
}

}  // namespace

// =================================================================================================
extern "C" {

int gritas_b200_version(void) { return 100; }

const char *gritas_b200_last_error(gritas_b200_handle h) { return h ? h->err : "null handle"; }

int gritas_b200_create(gritas_b200_handle *hp, int im, int jm, int km, int nlevs, int l2d, int l3d, int nr,
                       const double *p1, const double *p2, const int32_t *kxkt_table, int kxmax, int ktmax, int mode,
                       int device) {
  if (!hp) return GRITAS_B200_ERR_ARG;
  *hp = nullptr;
  if (im < 1 || jm < 2 || km < 1 || nlevs < 1 || nlevs > km || l2d < 0 || l3d < 0 || !p1 || !p2 || !kxkt_table ||
      kxmax < 1 || ktmax < 1)
    return GRITAS_B200_ERR_ARG;
  if (nr < 1) return GRITAS_B200_ERR_ARG;
  if (nr > 3) return GRITAS_B200_ERR_NR;  // m_gritas_grids.f:395-396
  if ((mode & GRITAS_B200_FLAG_MINMAX) && nr > 1) return GRITAS_B200_ERR_NR;  // 'MinMax: could not handle more then 1 field' (:1072)
  const int m = mode & 0xff;
  if (m != GRITAS_B200_MODE_FAST && m != GRITAS_B200_MODE_DETERMINISTIC) return GRITAS_B200_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    cudaGetLastError();
    return GRITAS_B200_ERR_CUDA;  // no CPU fallback
  }
  if (device < 0) {
    const char *lr = getenv("LOCAL_RANK");
    if (lr && *lr) device = atoi(lr) % ndev;
    else if (cudaGetDevice(&device) != cudaSuccess) device = 0;
  }
  if (device >= ndev) return GRITAS_B200_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return GRITAS_B200_ERR_CUDA;

  const size_t nbox3 = (size_t)im * jm * km * l3d, nbox2 = (size_t)im * jm * l2d;
  if (nbox3 + nbox2 >= (size_t)0xFFFFFFF0u) return GRITAS_B200_ERR_ARG;  // 32-bit record keys

  gritas_b200_handle h = new (std::nothrow) gritas_b200_ctx();
  if (!h) return GRITAS_B200_ERR_ARG;
  h->device = device;
  h->mode = m;
  h->flags = mode & ~0xff;
  h->minmax = (mode & GRITAS_B200_FLAG_MINMAX) != 0;
  GridDesc &g = h->g;
  g.im = im; g.jm = jm; g.km = km; g.nlevs = nlevs; g.l2d = l2d; g.l3d = l3d; g.nr = nr;
  g.rs = (nr == 1) ? 4 : 8;
  g.kxmax = kxmax; g.ktmax = ktmax;
  g.dlon = 360.0 / (double)im;        // m_gritas_grids.f:128
  g.dlat = 180.0 / (double)(jm - 1);  // m_gritas_grids.f:129
  g.nbox3 = (unsigned)nbox3; g.nbox2 = (unsigned)nbox2;
  h->nrec = nbox3 + nbox2;
  // GrITAS_Pint tables are contiguous (p1(k+1) is the same expression as p2(k)) and descending:
  // then the branch-free search is exact.  Anything else takes the literal linear scan.
  g.contiguous = 1;
  for (int k = 0; k + 1 < nlevs; ++k)
    if (!(p1[k + 1] == p2[k]) || !(p2[k] > p2[k + 1])) g.contiguous = 0;
  if (!(p1[0] > p2[0])) g.contiguous = 0;
  for (int k = 0; k < nlevs; ++k)
    if (p1[k] != p1[k] || p2[k] != p2[k]) g.contiguous = 0;
  g.p1_top = p1[0];
  g.rdlon = 1.0 / g.dlon;
  g.rdlat = 1.0 / g.dlat;
  g.rim = 1.0 / (double)im;
  g.rjm = 1.0 / (double)jm;
  // Level LUT for the branch-free lookup: cells of 2^-m octave on the high word of lev, chosen so that
  // no cell holds more than LUT_STEPS interfaces (kernels.cuh find_level).
  std::vector<unsigned short> lut;
  g.lut_ncell = 0;
  g.lut_shift = 0;
  g.lut_cell0 = 0;
  if (g.contiguous) {
    lut.resize(16384);
    g.lut_ncell = build_level_lut(nlevs, p1, p2, lut.data(), (int)lut.size(), &g.lut_shift, &g.lut_cell0);
  }
  g.p2_in_smem = (sizeof(double) * (size_t)nlevs + 2 * (size_t)g.lut_ncell + 8 <= 40 * 1024) ? 1 : 0;
  h->key_bits = 1;
  while (h->key_bits < 32 && ((size_t)1 << h->key_bits) <= h->nrec) h->key_bits++;

#define CUC(call)                                                                                       \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) {                                                                            \
      fprintf(stderr, "gritas_b200_create: %s: %s\n", #call, cudaGetErrorString(e_));                   \
      gritas_b200_destroy(h);                                                                           \
      return GRITAS_B200_ERR_CUDA;                                                                      \
    }                                                                                                   \
  } while (0)
  CUC(cudaStreamCreateWithFlags(&h->compute, cudaStreamNonBlocking));
  CUC(cudaStreamCreateWithFlags(&h->copy, cudaStreamNonBlocking));
  for (int i = 0; i < NSLOT; ++i) {
    h->slot[i].flags = std::vector<std::atomic<int>>(4096);
    CUC(cudaEventCreateWithFlags(&h->slot[i].consumed, cudaEventDisableTiming));
    CUC(cudaEventCreateWithFlags(&h->slot[i].copied, cudaEventDisableTiming));
  }
  CUC(cudaMalloc(&h->d_table, sizeof(int) * (size_t)kxmax * ktmax));
  CUC(cudaMalloc(&h->d_p1, sizeof(double) * nlevs));
  CUC(cudaMalloc(&h->d_p2, sizeof(double) * nlevs));
  CUC(cudaMalloc(&h->d_status, sizeof(Status) * (NWK + 1)));   // [0] compute stream, [1 + w] worker w
  CUC(cudaMallocHost(&h->h_status, sizeof(Status) * (NWK + 1)));
  const size_t acc_bytes = sizeof(double) * (h->nrec > 0 ? h->nrec : 1) * g.rs;
  CUC(cudaMalloc(&g.acc, acc_bytes));
  CUC(cudaMemcpy(h->d_table, kxkt_table, sizeof(int) * (size_t)kxmax * ktmax, cudaMemcpyDefault));
  CUC(cudaMemcpy(h->d_p1, p1, sizeof(double) * nlevs, cudaMemcpyDefault));
  CUC(cudaMemcpy(h->d_p2, p2, sizeof(double) * nlevs, cudaMemcpyDefault));
  if (g.lut_ncell > 0) {
    CUC(cudaMalloc(&h->d_lut, sizeof(unsigned short) * (size_t)g.lut_ncell));
    CUC(cudaMemcpy(h->d_lut, lut.data(), sizeof(unsigned short) * (size_t)g.lut_ncell, cudaMemcpyHostToDevice));
  }
  g.lut = h->d_lut;
  if (h->minmax) k_minmax_init<<<grid_for(h->nrec, 256, 8), 256, 0, h->compute>>>(g.acc, h->nrec);  // gritas.f:364-366
  else CUC(cudaMemsetAsync(g.acc, 0, acc_bytes, h->compute));  // gritas.f:352-363
  g.table = h->d_table; g.p1 = h->d_p1; g.p2 = h->d_p2;
  Status st[NWK + 1];
  for (Status &q : st) { q.first_bad = BAD_NONE; q.bad_lev = 0.0; q.near_pairs = 0; q.pairs = 0; }
  CUC(cudaMemcpyAsync(h->d_status, st, sizeof(st), cudaMemcpyHostToDevice, h->compute));
  CUC(cudaStreamSynchronize(h->compute));
  CUC(cudaEventCreateWithFlags(&h->main_ev, cudaEventDisableTiming));

  // ---- tiled fast path --------------------------------------------------
  // Worth it when the observations of a file scatter over a grid far larger than L2 (conventional data); data that
  // walks the records in order (radiance footprints) stays on the direct path -- decided by the locality probe.
  {
    const char *e_on = getenv("GRITAS_B200_STAGE"), *e_shift = getenv("GRITAS_B200_TILE_SHIFT"),
               *e_slots = getenv("GRITAS_B200_STAGE_SLOTS"), *e_max = getenv("GRITAS_B200_STAGE_MAX_GRID_MB");
    // (with K2f a Norm no longer moves the records, so large grids profit too: cfg4's 32 GB of records 28 -> 19 ms/step)
    const size_t max_grid = (e_max ? (size_t)atoll(e_max) : (size_t)65536) << 20;
    const bool det_tiles = m == GRITAS_B200_MODE_DETERMINISTIC && !(h->flags & GRITAS_B200_FLAG_DET_SORT);
    bool want = (m == GRITAS_B200_MODE_FAST || det_tiles) && !h->minmax && !(h->flags & GRITAS_B200_FLAG_NO_STAGING) && acc_bytes <= max_grid &&
                h->nrec >= 4096;
    if (e_on && atoi(e_on) == 0) want = false;
    if (want) {
      int shift = (nr == 1) ? 11 : 10;     // tile = 2048 / 1024 records (measured best of 256..2048 at cfg2): K2f keeps 2 CTAs per SM
      if (e_shift) shift = atoi(e_shift);
      const size_t T = (size_t)1 << shift;
      // the tile kernel (K2f / K2t) keeps the tile's sum planes and one sort chunk of its list in shared memory
      const size_t smem_f = nr == 1 ? TileFinal<1>::smem_bytes(T) : (nr == 2 ? TileFinal<2>::smem_bytes(T) : TileFinal<3>::smem_bytes(T));
      const size_t smem_d = nr == 1 ? TileDet<1>::smem_bytes(T) : (nr == 2 ? TileDet<2>::smem_bytes(T) : TileDet<3>::smem_bytes(T));
      const size_t smem = std::max(smem_f, det_tiles ? smem_d : (size_t)0);
      size_t free_b = 0, total_b = 0;
      CUC(cudaMemGetInfo(&free_b, &total_b));
      const size_t entry = (nr == 1) ? ListEntry<1>::BYTES : ListEntry<2>::BYTES;
      size_t slots = e_slots ? (size_t)atoll(e_slots) : std::min((size_t)384 << 20, std::max((size_t)8 << 20, 20 * h->nrec));
      slots = std::min(slots, free_b / 4 / entry);
      const size_t ntiles = (h->nrec + T - 1) / T;
      size_t cap = (slots / ntiles) & ~(size_t)31;
      if (cap >= 256 && smem <= (size_t)227 * 1024 && shift >= 8 && shift <= 15) {
        h->sg.tile_shift = shift;
        h->sg.det_r0 = 192u;
        if (const char *e_r0 = getenv("GRITAS_B200_DET_R0")) h->sg.det_r0 = (unsigned)std::max(32, atoi(e_r0));
        h->sg.ntiles = (unsigned)ntiles;
        h->sg.cap = (unsigned)std::min(cap, (size_t)0x7fffffe0);
        h->stage_slots = ntiles * (size_t)h->sg.cap;
        CUC(cudaMalloc(&h->sg.fill, sizeof(unsigned) * (2 * ntiles + 4)));   // tickets, the per-tile dirty words, the records-touched word
        CUC(cudaMemset(h->sg.fill, 0, sizeof(unsigned) * (2 * ntiles + 4)));
        h->sg.dirty = h->sg.fill + ntiles;
        CUC(cudaMalloc(&h->sg.lists, entry * h->stage_slots));
        if (const char *e_w = getenv("GRITAS_B200_WORKERS")) h->nwk = std::max(1, std::min(NWK, atoi(e_w)));
        if (const char *e_k = getenv("GRITAS_B200_K1_CTAS_PER_SM")) h->k1_per_sm = std::max(1, atoi(e_k));
        for (int w = 0; w < h->nwk; ++w) {                // one stream, event and level-miss record per worker
          CUC(cudaStreamCreateWithFlags(&h->wk[w], cudaStreamNonBlocking));
          CUC(cudaEventCreateWithFlags(&h->wk_done[w], cudaEventDisableTiming));
        }
        int nsm = 148, per_sm = 1;
        CUC(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device));
#define TF_SETUP(NR_)                                                                                                        \
  do {                                                                                                                        \
    CUC(cudaFuncSetAttribute(k_tile_final<NR_, TILE_FINALIZE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    CUC(cudaFuncSetAttribute(k_tile_final<NR_, TILE_FLUSH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
    CUC(cudaFuncSetAttribute(k_tile_final<NR_, TILE_EXPORT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    CUC(cudaFuncSetAttribute(k_tile_final<NR_, TILE_IMPORT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    CUC(cudaFuncSetAttribute(k_tile_final<NR_, TILE_PEERS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
    CUC(cudaFuncSetAttribute(k_tile_final<NR_, TILE_IMPORT_PEERS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tile_final<NR_, TILE_FINALIZE>, TileFinal<NR_>::THREADS, smem)); \
  } while (0)
        if (nr == 1) TF_SETUP(1);
        else if (nr == 2) TF_SETUP(2);
        else TF_SETUP(3);
#undef TF_SETUP
        h->tf_smem = smem_f;
        h->tf_per_sm = std::max(per_sm, 1);
        if (det_tiles) {
          int dps = 1;
#define TD_SETUP(NR_)                                                                                                           \
  do {                                                                                                                           \
    CUC(cudaFuncSetAttribute(k_tile_det<NR_, DET_FINALIZE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d));         \
    CUC(cudaFuncSetAttribute(k_tile_det<NR_, DET_FLUSH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d));            \
    CUC(cudaFuncSetAttribute(k_tile_det<NR_, DET_PEERS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d));            \
    CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&dps, k_tile_det<NR_, DET_FINALIZE>, TileDet<NR_>::THREADS, smem_d));     \
  } while (0)
          if (nr == 1) TD_SETUP(1);
          else if (nr == 2) TD_SETUP(2);
          else TD_SETUP(3);
#undef TD_SETUP
          h->td_smem = smem_d;
          h->td_ctas = (int)std::min((size_t)nsm * (size_t)std::max(dps, 1), ntiles);
          h->det_tiled = true;
          h->tiled_choice = 1;
          CUC(cudaMallocHost(&h->h_ctrl, sizeof(unsigned) * 4 * (NWK + 1)));
          memset(h->h_ctrl, 0, sizeof(unsigned) * 4 * (NWK + 1));
        }
        h->tf_ctas = (int)std::min((size_t)nsm * (size_t)h->tf_per_sm, ntiles);   // persistent: all CTAs resident
        // GRITAS_B200_FUSED_NORM=0: Norm flushes the lists into the records (K2t) and finalizes them with K3
        const char *e_f = getenv("GRITAS_B200_FUSED_NORM");
        h->fused_norm = !(e_f && atoi(e_f) == 0);
        h->staging = true;
        if (h->flags & GRITAS_B200_FLAG_FORCE_TILED) h->tiled_choice = 1;
        if (e_on && atoi(e_on) == 1) h->tiled_choice = 1;
      }
    }
  }
#undef CUC
  *hp = h;
  return GRITAS_B200_OK;
}

void gritas_b200_destroy(gritas_b200_handle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->compute) cudaStreamSynchronize(h->compute);
  if (h->copy) cudaStreamSynchronize(h->copy);
  if (h->early_st) { cudaStreamSynchronize(h->early_st); cudaStreamDestroy(h->early_st); }
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  for (int i = 0; i < NSLOT; ++i) {
    Slot &s = h->slot[i];
    for (int c = 0; c < NCOL; ++c) {
      if (s.dev[c]) cudaFree(s.dev[c]);
      if (s.pin[c]) cudaFreeHost(s.pin[c]);
    }
    if (s.flag_out) cudaFree(s.flag_out);
    if (s.flag_pin) cudaFreeHost(s.flag_pin);
    if (s.ints_copied) cudaEventDestroy(s.ints_copied);
    if (s.flags_ready) cudaEventDestroy(s.flags_ready);
    if (s.consumed) cudaEventDestroy(s.consumed);
    if (s.copied) cudaEventDestroy(s.copied);
  }
  for (int w = 0; w < NWK; ++w) {
    if (h->wk[w]) { cudaStreamSynchronize(h->wk[w]); cudaStreamDestroy(h->wk[w]); }
    if (h->wk_done[w]) cudaEventDestroy(h->wk_done[w]);
  }
  if (h->main_ev) cudaEventDestroy(h->main_ev);
  if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
  if (h->fin_stream) cudaStreamDestroy(h->fin_stream);
  for (int c = 0; c < gritas_b200_ctx::NCHUNK; ++c) {
    if (h->ev_acc[c]) cudaEventDestroy(h->ev_acc[c]);
    if (h->ev_red[c]) cudaEventDestroy(h->ev_red[c]);
  }
  if (h->ev_fin) cudaEventDestroy(h->ev_fin);
  for (int b = 0; b < gritas_b200_ctx::NBOUNCE; ++b) {
    if (h->bounce[b]) cudaFreeHost(h->bounce[b]);
    if (h->bounce_ev[b]) cudaEventDestroy(h->bounce_ev[b]);
  }
  for (int r = 0; r < MAX_PEERS; ++r)
    if (h->peer_ipc[r])
      for (void *p : {h->peer_lists[r], h->peer_fill[r], h->peer_acc[r], h->peer_red[r]})
        if (p) cudaIpcCloseMemHandle(p);
  for (void *p : {(void *)h->sg.fill, (void *)h->sg.lists, (void *)h->red, (void *)h->fill_snap, (void *)h->bar_buf, (void *)h->slab_prefix,
                  (void *)h->d_bounds})
    if (p) cudaFree(p);
  if (h->h_bounds) cudaFreeHost(h->h_bounds);
  if (h->bounds_ev) cudaEventDestroy(h->bounds_ev);
  for (void *p : {h->ps_cols, h->ps_key[0], h->ps_key[1], h->ps_idx[0], h->ps_idx[1], h->ps_sorted})
    if (p) cudaFree(p);
  for (auto &db : h->det) {
    if (db.st) { cudaStreamSynchronize(db.st); cudaStreamDestroy(db.st); }
    if (db.sorted) cudaEventDestroy(db.sorted);
    if (db.consumed) cudaEventDestroy(db.consumed);
    for (void *p : {(void *)db.keys, (void *)db.idx, (void *)db.keys2, (void *)db.idx2, db.tmp})
      if (p) cudaFree(p);
  }
  if (h->d_lut) cudaFree(h->d_lut);
  if (h->d_mask) cudaFree(h->d_mask);
  for (void *p : {(void *)h->d_table, (void *)h->d_p1, (void *)h->d_p2, (void *)h->d_status, (void *)h->g.acc,
                  h->cub_tmp, (void *)h->out})
    if (p) cudaFree(p);
  if (h->h_status) cudaFreeHost(h->h_status);
  if (h->h_ctrl) cudaFreeHost(h->h_ctrl);
  if (h->compute && h->own_compute) cudaStreamDestroy(h->compute);
  if (h->copy) cudaStreamDestroy(h->copy);
  cudaGetLastError();
  delete h;
}

// Residual columns as expressions of ODS columns (gritas.f:588-704): del(:,ir) = sa*a[ir] [+/- b[ir]] [+ c[ir]].
struct ResidualSpec {
  const double *b[3] = {nullptr, nullptr, nullptr}, *c[3] = {nullptr, nullptr, nullptr};
  signed char sa[3] = {1, 1, 1}, sb[3] = {0, 0, 0};
  int expr = 0, post = 0;
};

// Early ob_flag write-back (see accum_common).  begin: once the integer columns' DMAs are queued -- k_early_flags and the
// download of its result on a stream of their own; end: the flags are in the caller's array.
static int early_flags_begin(gritas_b200_handle h, Slot &s, const ObsCols &c, bool any_copy, int32_t *flag_user) {
  const size_t nb4 = sizeof(int) * (size_t)c.n;
  if (!h->early_st) CU(cudaStreamCreateWithFlags(&h->early_st, cudaStreamNonBlocking));
  if (!s.ints_copied) {
    CU(cudaEventCreateWithFlags(&s.ints_copied, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&s.flags_ready, cudaEventDisableTiming));
  }
  if (any_copy) {
    CU(cudaEventRecord(s.ints_copied, h->copy));
    CU(cudaStreamWaitEvent(h->early_st, s.ints_copied, 0));
  }
  int rc = grow_dev(h, (void **)&s.flag_out, &s.flag_out_cap, nb4);
  if (rc) return rc;
  k_early_flags<<<grid_for((size_t)c.n, 256, 4), 256, 0, h->early_st>>>(h->g, c.kx, c.kt, c.flag, s.flag_out, c.n);
  CU(cudaGetLastError());
  h->launches += 1;
  void *dst = flag_user;
  if (mem_kind(flag_user) == 0) {                        // pageable: land in a pinned buffer, the copy pool finishes the job
    if ((rc = grow_pin(h, &s.flag_pin, &s.flag_pin_cap, nb4))) return rc;
    dst = s.flag_pin;
  }
  CU(cudaMemcpyAsync(dst, s.flag_out, nb4, cudaMemcpyDeviceToHost, h->early_st));
  CU(cudaEventRecord(s.flags_ready, h->early_st));
  return 0;
}

static int early_flags_end(gritas_b200_handle h, Slot &s, int32_t *flag_user, int nobs) {
  const size_t nb4 = sizeof(int) * (size_t)nobs;
  CU(cudaEventSynchronize(s.flags_ready));
  if (mem_kind(flag_user) == 0) {
    CopyPool::get().submit_split(flag_user, s.flag_pin, nb4, &s.flag_done);
    CopyPool::get().wait(&s.flag_done);
  }
  h->d2h += nb4;
  return 0;
}

static int accum_common(gritas_b200_handle h, int nobs, const double *lat, const double *lon, const double *lev,
                        const double *const dcol[3], const double *obs, const double *xvec, const double *omf_for_scale,
                        const int32_t *kx, const int32_t *kt, const int32_t *flag_in, const int32_t *time,
                        FilterDesc f, int32_t *flag_out, double *bad_lev, const ResidualSpec *rs = nullptr) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (nobs < 0) return fail(h, GRITAS_B200_ERR_ARG, "nobs < 0");
  h->call_seq += 1;
  h->export_valid = false;
  if (nobs == 0) return GRITAS_B200_OK;
  if (!lat || !lon || !lev || !kx || !kt) return fail(h, GRITAS_B200_ERR_ARG, "null column");
  for (int ir = 0; ir < h->g.nr; ++ir)
    if (!dcol[ir]) return fail(h, GRITAS_B200_ERR_ARG, "null residual column %d", ir + 1);

  Slot &s = h->slot[h->next_slot];
  h->next_slot = (h->next_slot + 1) % NSLOT;
  s.direct_dma = false;
  CallTrace trace;
  struct TraceScope { CallTrace *prev; TraceScope(CallTrace *t) : prev(g_trace) { g_trace = t->on ? t : nullptr; } ~TraceScope() { g_trace = prev; } } trace_scope(&trace);

  struct PendingGuard {   // an error return must not leave pool tasks reading the caller's arrays
    Slot &s;
    ~PendingGuard() {
      for (const PendingCopy &pc : s.pending)
        if (pc.submitted) CopyPool::get().wait(pc.done);
      s.pending.clear();
      s.nflags = 0;
    }
  } guard{s};
  ObsCols c = {};
  c.n = nobs;
  bool any = false;
  const size_t nb8 = sizeof(double) * (size_t)nobs, nb4 = sizeof(int) * (size_t)nobs;
  // a source column may appear in several roles: stage it once per call
  const void *seen_src[NCOL] = {}, *seen_dev[NCOL] = {};
  int nseen = 0;
  auto stage1 = [&](int col, const void *src, size_t bytes, const void **out) -> int {
    for (int q = 0; q < nseen; ++q)
      if (seen_src[q] == src) { *out = seen_dev[q]; return 0; }
    const int r = stage(h, s, col, src, bytes, out, &any);
    if (!r && nseen < NCOL) { seen_src[nseen] = src; seen_dev[nseen] = *out; ++nseen; }
    return r;
  };
  // The integer columns go first.  GrITAS_Accum's ob_flag (inout) depends on them alone (m_gritas_grids.f:406-416), so for
  // a host caller the flags are computed and sent back while the fp64 columns are still on their way up (early write-back):
  // the call no longer ends with a download that nothing overlaps.
  if ((rc = stage1(C_KX, kx, nb4, (const void **)&c.kx))) return rc;
  if ((rc = stage1(C_KT, kt, nb4, (const void **)&c.kt))) return rc;
  if (flag_in && (rc = stage1(C_FLAG, flag_in, nb4, (const void **)&c.flag))) return rc;
  static const bool early_off = getenv("GRITAS_B200_EARLY_FLAGS") && atoi(getenv("GRITAS_B200_EARLY_FLAGS")) == 0;
  const bool early = flag_out && flag_in == flag_out && !f.ods && !h->g.mask && !h->minmax && !early_off && nobs >= 4096 && nobs <= (1 << 27) &&
                     mem_kind(flag_out) != 2;
  const std::function<int()> early_hook = [&]() -> int { return early_flags_begin(h, s, c, any, flag_out); };
  const size_t n_int_chunks = s.pending.size();
  if (early && n_int_chunks == 0) {                      // pinned (or device) integer columns: their DMAs are queued already
    if ((rc = early_hook())) return rc;
  }
  if ((rc = stage1(C_LAT, lat, nb8, (const void **)&c.lat))) return rc;
  if ((rc = stage1(C_LON, lon, nb8, (const void **)&c.lon))) return rc;
  if ((rc = stage1(C_LEV, lev, nb8, (const void **)&c.lev))) return rc;
  for (int ir = 0; ir < h->g.nr; ++ir)
    if ((rc = stage1(C_D0 + ir, dcol[ir], nb8, (const void **)&c.d[ir]))) return rc;
  for (int ir = 0; ir < 3; ++ir) { c.sa[ir] = 1; c.sb[ir] = 0; }
  if (rs && rs->expr) {
    f.expr = 1;
    f.post = rs->post;
    for (int ir = 0; ir < h->g.nr; ++ir) {
      c.sa[ir] = rs->sa[ir];
      c.sb[ir] = rs->sb[ir];
      if (rs->b[ir] && (rc = stage1(C_B0 + ir, rs->b[ir], nb8, (const void **)&c.db[ir]))) return rc;
      if (rs->c[ir] && (rc = stage1(C_C0 + ir, rs->c[ir], nb8, (const void **)&c.dc[ir]))) return rc;
    }
  }
  if ((f.scale_res == 1 || f.post) && (rc = stage1(C_XVEC, xvec, nb8, (const void **)&c.xvec))) return rc;
  if (f.scale_res >= 2 && (rc = stage1(C_OBS, obs, nb8, (const void **)&c.obs))) return rc;
  if (f.scale_res == 3 && (rc = stage1(C_D2, omf_for_scale, nb8, (const void **)&c.omf))) return rc;   // obs - omf
  if (f.use_time && (rc = stage1(C_TIME, time, nb4, (const void **)&c.time))) return rc;
  trace_mark("stage");
  if (!early) return run_accum(h, c, f, s, any, flag_out, bad_lev);
  if (n_int_chunks > 0 && (rc = flush_pending(h, s, n_int_chunks, &early_hook))) return rc;
  trace_mark("flush");
  // every upload is queued; the flags came down long ago: hand them to the caller while the last DMAs and K1 run
  if ((rc = early_flags_end(h, s, flag_out, nobs))) return rc;
  trace_mark("flags");
  return run_accum(h, c, f, s, any, nullptr, bad_lev);
}

int gritas_b200_accum(gritas_b200_handle h, int nobs, const double *lat, const double *lon, const double *lev,
                      const double *del, const int32_t *kx, const int32_t *kt, int32_t *ob_flag, double *bad_lev) {
  if (!h) return GRITAS_B200_ERR_ARG;
  if (nobs > 0 && !del) return fail(h, GRITAS_B200_ERR_ARG, "null del");
  const double *d[3] = {del, h->g.nr > 1 ? del + (size_t)nobs : nullptr, h->g.nr > 2 ? del + 2 * (size_t)nobs : nullptr};
  FilterDesc f = {};
  int32_t *flag_out = (h->flags & GRITAS_B200_FLAG_NO_WRITEBACK) ? nullptr : ob_flag;
  return accum_common(h, nobs, lat, lon, lev, d, nullptr, nullptr, nullptr, kx, kt, ob_flag, nullptr, f, flag_out, bad_lev);
}

int gritas_b200_accum_ods(gritas_b200_handle h, int nobs, const double *lat, const double *lon, const double *lev,
                          const int32_t *time, const int32_t *kt, const int32_t *kx, const int32_t *qcexcl,
                          const double *omf, const double *oma, const double *obs, const double *xvec, int iopt,
                          int scale_res, int ioptn, int t1, int t2, double lona, double lonb, int x_passive,
                          int x_pre_bad, int32_t *ob_flag_out, double *bad_lev) {
  if (!h) return GRITAS_B200_ERR_ARG;
  const double *d[3] = {nullptr, nullptr, nullptr};
  const int a = iopt < 0 ? -iopt : iopt;
  int want_nr = 1;
  if (a == 0 || a == GRITAS_B200_IOPT_OMF) d[0] = omf;                       // gritas.f:562-576
  else if (a == 2 || a == GRITAS_B200_IOPT_OMA) d[0] = oma;                  // gritas.f:575-587
  else if (a == GRITAS_B200_IOPT_OMF_OMA) { d[0] = omf; d[1] = oma; want_nr = 2; }
  else return fail(h, GRITAS_B200_ERR_ARG, "iopt %d not supported by the fused entry", iopt);
  if (want_nr != h->g.nr) return fail(h, GRITAS_B200_ERR_ARG, "iopt %d needs nr=%d, handle has nr=%d", iopt, want_nr, h->g.nr);
  if (scale_res < 0 || scale_res > 3) return fail(h, GRITAS_B200_ERR_ARG, "scale_res out of range");
  if (nobs > 0 && !qcexcl) return fail(h, GRITAS_B200_ERR_ARG, "null qcexcl");
  if (nobs > 0 && scale_res == 1 && !xvec) return fail(h, GRITAS_B200_ERR_ARG, "scale_res=1 needs xvec");
  if (nobs > 0 && scale_res >= 2 && !obs) return fail(h, GRITAS_B200_ERR_ARG, "scale_res=2,3 needs obs");
  if (nobs > 0 && scale_res == 3 && !omf) return fail(h, GRITAS_B200_ERR_ARG, "scale_res=3 needs omf");
  FilterDesc f = {};
  // the filter only runs for odd iopt (gritas.f:710); otherwise ob_flag stays 0 (gritas.f:551)
  const bool filtered = (a & 1) != 0;
  f.ods = filtered ? 1 : 0;
  f.ioptn = ioptn;
  f.t1 = t1; f.t2 = t2;
  f.use_time = (!filtered || t2 < t1) ? 0 : 1;                       // gritas.f:717-723
  f.use_lon = (!filtered || (lona < -180.0 && lonb > 180.0)) ? 0 : 1;  // gritas.f:724-730
  f.lona = lona; f.lonb = lonb;
  f.x_passive = x_passive; f.x_pre_bad = x_pre_bad;
  f.scale_res = scale_res;
  if (nobs > 0 && f.use_time && !time) return fail(h, GRITAS_B200_ERR_ARG, "time window needs the time column");
  return accum_common(h, nobs, lat, lon, lev, d, obs, xvec, omf, kx, kt, filtered ? qcexcl : nullptr, time, f,
                      ob_flag_out, bad_lev);
}

// Fused entry for every residual option of gritas.f:562-704 that needs one ODS structure only
// (no -meanres second file): |iopt| 0..25, 28..33 and GRITAS_B200_IOPT_OMF_OMA.
int gritas_b200_accum_odsx(gritas_b200_handle h, int nobs, const gritas_b200_ods *o, int iopt, int options, int scale_res,
                           int ioptn, int t1, int t2, double lona, double lonb, int x_passive, int x_pre_bad,
                           int32_t *ob_flag_out, double *bad_lev) {
  if (!h || !o) return GRITAS_B200_ERR_ARG;
  const int a = iopt < 0 ? -iopt : iopt;
  const bool tch = (options & GRITAS_B200_OPT_TCH) != 0, self = (options & GRITAS_B200_OPT_SELF) != 0;
  const double *A[3] = {nullptr, nullptr, nullptr};
  ResidualSpec rs;
  int nr = 1;
  auto col = [&](int ir, const double *pa, int sa, const double *pb = nullptr, int sb = 0, const double *pc = nullptr) {
    A[ir] = pa; rs.sa[ir] = (signed char)sa; rs.b[ir] = pb; rs.sb[ir] = (signed char)sb; rs.c[ir] = pc;
    if (sa < 0 || pb || pc) rs.expr = 1;
  };
  const double *omf = o->omf, *oma = o->oma, *obs = o->obs, *xvec = o->xvec, *xm = o->xm;
  switch (a) {
    case 0: case 1: col(0, omf, 1); break;                                              // gritas.f:562-574
    case 2: case 3: col(0, oma, 1); break;                                              // :575-587
    case 4: case 5: col(0, obs, 1); break;                                              // :588-589
    case 6: case 7: col(0, xvec, 1); break;                                             // :590-591 prescribed sigo
    case 8: case 9: col(0, xm, 1); break;                                               // :592-593 observation bias
    case 10: case 11:                                                                   // :594-611 <AmF,OmF>
      if (tch) { col(0, oma, 1, omf, -1); col(1, omf, -1); col(2, oma, -1); nr = 3; }
      else { col(0, omf, 1, oma, -1); col(1, omf, 1); nr = 2; }
      break;
    case 12: case 13:                                                                   // :612-624 <AmF,OmA>
      if (tch) { col(0, oma, -1); col(1, omf, 1, oma, -1); col(2, omf, 1); nr = 3; }
      else { col(0, omf, 1, oma, -1); col(1, oma, 1); nr = 2; }
      break;
    case 14: case 15:                                                                   // :625-640 <OmA,OmF>
      if (tch && self) { col(0, obs, 1); col(1, obs, 1, omf, -1); col(2, obs, 1, oma, -1); nr = 3; }
      else if (tch) { col(0, omf, 1); col(1, oma, 1); col(2, oma, 1, omf, -1); nr = 3; }
      else { col(0, oma, 1); col(1, omf, 1); nr = 2; }
      break;
    case 16: case 17: col(0, omf, 1); col(1, xvec, 1); nr = 2; if (a == 17) { rs.post = 2; rs.expr = 1; } break;   // :641-644, :738
    case 18: case 19: col(0, oma, 1); col(1, xvec, 1); nr = 2; if (a == 19) { rs.post = 2; rs.expr = 1; } break;   // :645-648, :738
    case 20: case 21: col(0, obs, 1, xm, 1); break;                                     // :649-650 original observation
    case 22: case 23: col(0, xvec, 1); break;                                           // :651-652 observation impacts
    case 24: case 25: col(0, obs, 1, omf, -1); break;                                   // :653-664 (no -meanres)
    case 28: case 29: col(0, obs, 1, omf, -1, xm); break;                               // :678-679 h(xf)
    case 30: case 31: col(0, obs, 1, oma, -1); break;                                   // :680-691 (no -meanres)
    case 32: case 33: col(0, omf, 1, oma, -1); col(1, oma, 1); nr = 2; rs.post = 1; rs.expr = 1; break;   // :692-701
    case GRITAS_B200_IOPT_OMF_OMA: col(0, omf, 1); col(1, oma, 1); nr = 2; break;
    default: return fail(h, GRITAS_B200_ERR_ARG, "iopt %d: not valid / needs -meanres (gritas.f:702-703)", iopt);
  }
  if (nr != h->g.nr) return fail(h, GRITAS_B200_ERR_ARG, "iopt %d needs nr=%d, handle has nr=%d", iopt, nr, h->g.nr);
  if (scale_res < 0 || scale_res > 3) return fail(h, GRITAS_B200_ERR_ARG, "scale_res out of range");
  if (a > 3) scale_res = 0;                                   // the reference scales only the omf / oma options
  if (nobs > 0) {
    for (int ir = 0; ir < nr; ++ir)
      if (!A[ir] || (rs.sb[ir] && !rs.b[ir])) return fail(h, GRITAS_B200_ERR_ARG, "iopt %d: a column it needs is null", iopt);
    if (!o->qcexcl) return fail(h, GRITAS_B200_ERR_ARG, "null qcexcl");
    if ((scale_res == 1 || rs.post) && !xvec) return fail(h, GRITAS_B200_ERR_ARG, "xvec needed");
    if (scale_res >= 2 && !obs) return fail(h, GRITAS_B200_ERR_ARG, "scale_res=2,3 needs obs");
    if (scale_res == 3 && !omf) return fail(h, GRITAS_B200_ERR_ARG, "scale_res=3 needs omf");
  }
  FilterDesc f = {};
  const bool filtered = (a & 1) != 0;                         // gritas.f:710: the filter runs for odd iopt only
  f.ods = filtered ? 1 : 0;
  f.ioptn = ioptn;
  f.t1 = t1; f.t2 = t2;
  f.use_time = (!filtered || t2 < t1) ? 0 : 1;
  f.use_lon = (!filtered || (lona < -180.0 && lonb > 180.0)) ? 0 : 1;
  f.lona = lona; f.lonb = lonb;
  f.x_passive = x_passive; f.x_pre_bad = x_pre_bad;
  f.scale_res = scale_res;
  if (nobs > 0 && f.use_time && !o->time) return fail(h, GRITAS_B200_ERR_ARG, "time window needs the time column");
  return accum_common(h, nobs, o->lat, o->lon, o->lev, A, obs, xvec, omf, o->kx, o->kt, filtered ? o->qcexcl : nullptr,
                      o->time, f, ob_flag_out, bad_lev, &rs);
}

// The -meanres variants of gritas.f:653-677 and 680-691 (a member ODS against the ensemble-mean ODS, both matched to the same
// order): |iopt| 24/25 del = <omf> - omf, 26/27 del(:,1) = <omf> - omf, del(:,2) = xvec (nr = 2), 30/31 del = <oma> - oma.
// Where the observation value differs between the two files the reference makes the observation passive and leaves del
// unset (:656-662); here del is the same difference, which only matters if passive data are kept (ioptn > 0).
int gritas_b200_accum_odsx_meanres(gritas_b200_handle h, int nobs, const gritas_b200_ods *o, const double *mean_obs,
                                   const double *mean_omf, const double *mean_oma, int iopt, int ioptn, int t1, int t2, double lona,
                                   double lonb, int x_passive, int x_pre_bad, int32_t *ob_flag_out, double *bad_lev) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (!o || nobs < 0) return fail(h, GRITAS_B200_ERR_ARG, "accum_odsx_meanres: bad arguments");
  const int a = iopt < 0 ? -iopt : iopt;
  const bool f24 = a == 24 || a == 25, f26 = a == 26 || a == 27, f30 = a == 30 || a == 31;
  if (!f24 && !f26 && !f30) return fail(h, GRITAS_B200_ERR_ARG, "iopt %d has no -meanres variant (gritas.f:653-691)", iopt);
  const int nr = f26 ? 2 : 1;
  if (nr != h->g.nr) return fail(h, GRITAS_B200_ERR_ARG, "iopt %d needs nr=%d, handle has nr=%d", iopt, nr, h->g.nr);
  if (nobs == 0) {
    const double *none[3] = {nullptr, nullptr, nullptr};
    FilterDesc f0 = {};
    return accum_common(h, 0, nullptr, nullptr, nullptr, none, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, f0, nullptr, bad_lev);
  }
  const double *mres = f30 ? mean_oma : mean_omf, *res = f30 ? o->oma : o->omf;
  if (!o->obs || !mean_obs || !mres || !res || !o->qcexcl || (f26 && !o->xvec)) return fail(h, GRITAS_B200_ERR_ARG, "accum_odsx_meanres: a column it needs is null");
  const size_t n = (size_t)nobs;
  // qcexcl with the mismatches made passive, on the device (scratch of the pre-sort)
  if ((rc = join_workers(h))) return rc;
  CU(cudaStreamSynchronize(h->compute));
  const size_t f8 = (8 * n + 255) & ~(size_t)255, i4 = (4 * n + 255) & ~(size_t)255;
  if ((rc = grow_dev(h, &h->ps_cols, &h->ps_cols_cap, 2 * f8 + 2 * i4 + 256))) return rc;
  char *scr = (char *)h->ps_cols;
  const void *src[3] = {o->obs, mean_obs, o->qcexcl};
  const size_t bytes[3] = {8 * n, 8 * n, 4 * n}, slot[3] = {f8, f8, i4};
  const void *dev[3];
  size_t off = 0;
  for (int q = 0; q < 3; ++q) {
    if (mem_kind(src[q]) == 2) { dev[q] = src[q]; continue; }
    CU(cudaMemcpyAsync(scr + off, src[q], bytes[q], cudaMemcpyHostToDevice, h->compute));
    h->h2d += bytes[q];
    dev[q] = scr + off;
    off += slot[q];
  }
  int *qc2 = (int *)(scr + 2 * f8 + i4);
  k_meanres_qc<<<grid_for(n, 256, 8), 256, 0, h->compute>>>((const double *)dev[0], (const double *)dev[1], (const int *)dev[2], x_passive,
                                                             qc2, nobs);
  h->launches++;
  CU(cudaGetLastError());
  if ((rc = mark_main(h))) return rc;
  const double *A[3] = {mres, f26 ? o->xvec : nullptr, nullptr};
  ResidualSpec rs;
  rs.b[0] = res;            // del(:,1) = <res> - res
  rs.sb[0] = -1;
  rs.expr = 1;
  FilterDesc f = {};
  const bool filtered = (a & 1) != 0;                         // gritas.f:710: the filter runs for odd iopt only
  f.ods = filtered ? 1 : 0;
  f.ioptn = ioptn;
  f.t1 = t1; f.t2 = t2;
  f.use_time = (!filtered || t2 < t1) ? 0 : 1;
  f.use_lon = (!filtered || (lona < -180.0 && lonb > 180.0)) ? 0 : 1;
  f.lona = lona; f.lonb = lonb;
  f.x_passive = x_passive; f.x_pre_bad = x_pre_bad;
  if (f.use_time && !o->time) return fail(h, GRITAS_B200_ERR_ARG, "time window needs the time column");
  return accum_common(h, nobs, o->lat, o->lon, o->lev, A, o->obs, o->xvec, o->omf, o->kx, o->kt, filtered ? qc2 : nullptr, o->time, f,
                      ob_flag_out, bad_lev, &rs);
}

int gritas_b200_sync(gritas_b200_handle h, double *bad_lev) {
  int rc = check_handle(h);
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->copy));
  if ((rc = flush_stage(h))) return rc;
  return drain(h, bad_lev);
}

int gritas_b200_wait(gritas_b200_handle h, double *bad_lev) {
  int rc = check_handle(h);
  if (rc) return rc;
  return settle(h, true, bad_lev);
}

int gritas_b200_set_call_order(gritas_b200_handle h, uint64_t order) {
  int rc = check_handle(h);
  if (rc) return rc;
  h->call_order = order;
  h->call_order_set = true;
  return 0;
}

int gritas_b200_flush(gritas_b200_handle h) {
  int rc = check_handle(h);
  if (rc) return rc;
  return flush_stage(h);
}

int gritas_b200_reset(gritas_b200_handle h) {
  int rc = check_handle(h);
  if (rc) return rc;
  if ((rc = join_workers(h))) return rc;
  if (h->minmax) {
    k_minmax_init<<<grid_for(h->nrec, 256, 8), 256, 0, h->compute>>>(h->g.acc, h->nrec);
  } else if (h->staging && h->acc_fresh) {
    // nothing was flushed into the records since they were last zeroed: only a list overflow can have touched them
    k_zero_dirty_tiles<<<148 * 8, 256, 0, h->compute>>>(h->g.acc, h->nrec * (size_t)h->g.rs, h->g.rs, h->sg);
    h->launches++;
  } else {
    CU(cudaMemsetAsync(h->g.acc, 0, sizeof(double) * h->nrec * h->g.rs, h->compute));
  }
  h->acc_fresh = true;
  h->export_valid = false;
  h->det_calls = 0;
  h->call_order_set = false;
  for (bool &v : h->ctrl_valid) v = false;
  if (h->staging) {
    CU(cudaMemsetAsync(h->sg.fill, 0, sizeof(unsigned) * (2 * (size_t)h->sg.ntiles + 4), h->compute));   // tickets, dirty words, records-touched word
    h->staged_upper = 0;
  }
  return mark_main(h);
}

int gritas_b200_barrier(gritas_b200_handle h) {
  int rc = check_handle(h);
  if (rc) return rc;
  return join_workers(h);
}

// K3 launches for 3-D rows [jl0, jl1) (jl = j + jm*l) and, if with2d, the surface boxes, on stream `sm`.
static int launch_finalize(gritas_b200_handle h, cudaStream_t sm, double *const dst[10], int raw, bool raw3, int xpl, int nrmlz,
                           int ntr, int jl0, int jl1, bool with2d, size_t b2_0 = 0, size_t b2_1 = 0) {
  const GridDesc &g = h->g;
  const int nr = g.nr;
  const size_t n2 = g.nbox2;
  if (n2 && with2d) {
    OutPlanes o = {dst[0], dst[1], dst[2], dst[3], dst[4], xpl};
    const int nb = grid_for(n2, 256, 8);
    if (nr == 1) k_finalize2<1><<<nb, 256, 0, sm>>>(g, o, nrmlz, ntr, raw, b2_0, b2_1 ? b2_1 : n2);
    else if (nr == 2) k_finalize2<2><<<nb, 256, 0, sm>>>(g, o, nrmlz, ntr, raw, b2_0, b2_1 ? b2_1 : n2);
    else k_finalize2<3><<<nb, 256, 0, sm>>>(g, o, nrmlz, ntr, raw, b2_0, b2_1 ? b2_1 : n2);
    h->launches++;
  }
  if (g.nbox3 && jl1 > jl0) {
    OutPlanes o = {dst[5], dst[6], dst[7], dst[8], dst[9], (raw3 && (raw == 0 || raw == 3)) ? 1 : xpl};
    dim3 blk(32, 32), grd((g.km + 31) / 32, (g.im + 31) / 32, 1);
    const int z = jl1 - jl0;
    grd.z = (unsigned)(z > 65535 ? 65535 : z);
    const int r3 = raw == 2 ? 2 : (raw3 ? 1 : raw);
    const size_t fsm = sizeof(double) * (size_t)(4 * nr + 1) * 32 * 33;
    if (nr == 1) {
      CU(cudaFuncSetAttribute(k_finalize3<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
      k_finalize3<1><<<grd, blk, fsm, sm>>>(g, o, nrmlz, ntr, r3, jl0, jl1);
    } else if (nr == 2) {
      CU(cudaFuncSetAttribute(k_finalize3<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
      k_finalize3<2><<<grd, blk, fsm, sm>>>(g, o, nrmlz, ntr, r3, jl0, jl1);
    } else {
      CU(cudaFuncSetAttribute(k_finalize3<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
      k_finalize3<3><<<grd, blk, fsm, sm>>>(g, o, nrmlz, ntr, r3, jl0, jl1);
    }
    h->launches++;
  }
  CU(cudaGetLastError());
  return 0;
}

// GRITAS_B200_FORCE_PIPELINE=1 (tests): a single rank runs the multi-rank Norm pipeline, minus the collective.
static bool force_pipeline() {
  const char *e = getenv("GRITAS_B200_FORCE_PIPELINE");
  return e && atoi(e) != 0;
}

// Finalize (raw=0) or export the raw sums (raw=1) into reference-layout planes.
// reduce_root < 0: the accumulators are final (flushed, all-reduced if wanted) -- finalize them.
// reduce_root >= 0 (multi-rank): tile flush, ncclReduce to the root and finalize are pipelined chunk by
// chunk; only the root finalizes and writes outputs.
static int finalize_to(gritas_b200_handle h, int raw, int nrmlz, int ntresh, double *const user[10], const double **dev_out,
                       int reduce_root = -1) {
  // order: bias_2d rtms_2d stdv_2d xcov_2d nobs_2d bias_3d rtms_3d stdv_3d xcov_3d nobs_3d
  const GridDesc &g = h->g;
  const int nr = g.nr;
  if (!raw && nr > 2) return fail(h, GRITAS_B200_ERR_NR, "Accum: could not handle more then 2 residuals");  // :555
  if (raw == 3 && nr != 3) return fail(h, GRITAS_B200_ERR_NR, "GrITAS_3CH_Norm_ could not handle more then 2 residuals");  // :1240
  const size_t n2 = g.nbox2, n3 = g.nbox3;
  if (!raw && h->minmax) return fail(h, GRITAS_B200_ERR_ARG, "Norm on a MinMax handle: use gritas_b200_get_minmax (gritas.f:884 skips Norm)");
  const int xpl = (raw == 1 || raw == 2) ? 1 : nr;    // raw == 3: GrITAS_3CH_Norm_, three 3CH planes in xcov
  const size_t sz[10] = {n2 * nr, n2 * nr, n2 * nr, n2 * xpl, n2, n3 * nr, n3 * nr, n3 * nr, n3 * xpl, n3};
  // nlevs<2: the reference returns before the 3-D loop (m_gritas_grids.f:607) -> 3-D sums stay raw
  const bool raw3 = raw == 1 || raw == 2 || g.nlevs < 2;   // (also m_gritas_grids.f:1292 for the 3CH variant)
  const bool piped = reduce_root >= 0 && ((h->comm != nullptr && h->nranks > 1) || force_pipeline());
  const bool writer = !piped || h->rank == reduce_root;
  double *dst[10];
  bool bounce[10];
  size_t need = 0;
  for (int a = 0; a < 10; ++a) {
    bounce[a] = false;
    dst[a] = nullptr;
    const bool skip = !writer || (!user[a] && !dev_out) || sz[a] == 0 || (raw == 1 && (a == 2 || a == 7)) ||
                      (raw == 2 && (a == 1 || a == 3 || a == 6 || a == 8)) || (!raw && raw3 && a == 7) ||
                      (raw == 3 && raw3 && (a == 7 || a == 8));
    if (skip) continue;
    if (!dev_out && mem_kind(user[a]) == 2) dst[a] = user[a];
    else { bounce[a] = true; need += sz[a]; }
  }
  if (need) {
    int rc = grow_dev(h, (void **)&h->out, &h->out_cap, sizeof(double) * need);
    if (rc) return rc;
    size_t off = 0;
    for (int a = 0; a < 10; ++a)
      if (bounce[a]) { dst[a] = h->out + off; off += sz[a]; }
  }
  const int ntr = ntresh < 0 ? 0 : ntresh;
  const int nrows = n3 ? g.jm * g.l3d : 0;
  if (!piped && raw != 2 && fused_norm_ok(h)) {
    int rc = join_workers(h);
    if (rc) return rc;
    const OutPlanes o2 = {dst[0], dst[1], dst[2], dst[3], dst[4], xpl};
    const OutPlanes o3 = {dst[5], dst[6], dst[7], dst[8], dst[9], (raw3 && (raw == 0 || raw == 3)) ? 1 : xpl};
    const int r3 = raw3 ? 1 : raw;
    if (h->det_tiled) {
      launch_det_nr<DET_FINALIZE>(h, h->compute, 0u, h->sg.ntiles, o2, o3, nrmlz, ntr, raw, r3, PeerSet{});
    } else {
      switch (nr) {
        case 1: launch_tile_final<1>(h, h->compute, o2, o3, nrmlz, ntr, raw, r3); break;
        case 2: launch_tile_final<2>(h, h->compute, o2, o3, nrmlz, ntr, raw, r3); break;
        default: launch_tile_final<3>(h, h->compute, o2, o3, nrmlz, ntr, raw, r3); break;
      }
    }
    CU(cudaGetLastError());
    h->launches++;
    if ((rc = mark_main(h))) return rc;     // later K1 launches append to the lists this kernel reads
  } else if (!piped) {
    int rc = launch_finalize(h, h->compute, dst, raw, raw3, xpl, nrmlz, ntr, 0, nrows, true);
    if (rc) return rc;
  } else {
    int rc = join_workers(h);
    if (rc) return rc;
    if (!h->comm_stream) {
      CU(cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
      CU(cudaStreamCreateWithFlags(&h->fin_stream, cudaStreamNonBlocking));
      for (int c = 0; c < gritas_b200_ctx::NCHUNK; ++c) {
        CU(cudaEventCreateWithFlags(&h->ev_acc[c], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_red[c], cudaEventDisableTiming));
      }
      CU(cudaEventCreateWithFlags(&h->ev_fin, cudaEventDisableTiming));
    }
    static const int NC = getenv("GRITAS_B200_PIPE_CHUNKS") ? std::min((int)gritas_b200_ctx::NCHUNK, std::max(1, atoi(getenv("GRITAS_B200_PIPE_CHUNKS")))) : 8;
    static const int pipe_free_sms = getenv("GRITAS_B200_PIPE_FREE_SMS") ? std::max(0, atoi(getenv("GRITAS_B200_PIPE_FREE_SMS"))) : 0;
    const bool staged = h->staging && h->staged_upper > 0;
    const size_t T = h->staging ? ((size_t)1 << h->sg.tile_shift) : 1;
    static const bool compact_off = getenv("GRITAS_B200_COMPACT_REDUCE") && atoi(getenv("GRITAS_B200_COMPACT_REDUCE")) == 0;
    if (h->staging && h->fused_norm && raw != 2 && !compact_off) {
      // (The choice must be the same on every rank -- it decides the size of the collective -- so it depends on the
      // handle's configuration only, not on whether this rank has anything parked: a rank whose data went to the
      // records, or that has none, exports those.)
      // Tiled path: the month is parked in the tile lists.  Per chunk of tiles: K2x accumulates the lists and exports
      // {n, sums} as compact planes (48 bytes per box at nr = 2 instead of a 64-byte record; records that already hold
      // part of a tile are folded in), ncclReduce sums the planes to the root, K3x finalizes them there.  The records and
      // the lists are not touched: like the one-rank Norm, this does not change the accumulation state.
      const int nv1 = 1 + (nr == 1 ? 2 : (nr == 2 ? 5 : 6));
      const size_t per_tile = (size_t)nv1 * T;
      if ((rc = grow_dev(h, (void **)&h->red, &h->red_cap, sizeof(double) * per_tile * h->sg.ntiles))) return rc;
      const OutPlanes o2 = {dst[0], dst[1], dst[2], dst[3], dst[4], xpl};
      const OutPlanes o3 = {dst[5], dst[6], dst[7], dst[8], dst[9], (raw3 && (raw == 0 || raw == 3)) ? 1 : xpl};
      const int r3 = raw3 ? 1 : raw;
      for (int c = 0; c < NC; ++c) {
        const unsigned t0 = (unsigned)((unsigned long long)h->sg.ntiles * c / NC), t1 = (unsigned)((unsigned long long)h->sg.ntiles * (c + 1) / NC);
        switch (nr) {
          case 1: launch_tile_export<1>(h, h->compute, t0, t1); break;
          case 2: launch_tile_export<2>(h, h->compute, t0, t1); break;
          default: launch_tile_export<3>(h, h->compute, t0, t1); break;
        }
        h->launches++;
        CU(cudaEventRecord(h->ev_acc[c], h->compute));
        CU(cudaStreamWaitEvent(h->comm_stream, h->ev_acc[c], 0));
        if (t1 > t0 && h->nranks > 1) {
          double *seg = h->red + (size_t)t0 * per_tile;
          int e = g_nccl.Reduce(seg, seg, (size_t)(t1 - t0) * per_tile, NCCL_FLOAT64, NCCL_SUM, reduce_root, h->comm, h->comm_stream);
          if (e != 0) return fail(h, GRITAS_B200_ERR_NCCL, "ncclReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "?");
        }
        CU(cudaEventRecord(h->ev_red[c], h->comm_stream));
        if (writer) {
          CU(cudaStreamWaitEvent(h->fin_stream, h->ev_red[c], 0));
          switch (nr) {
            case 1: launch_tile_import<1>(h, h->fin_stream, o2, o3, nrmlz, ntr, raw, r3, t0, t1); break;
            case 2: launch_tile_import<2>(h, h->fin_stream, o2, o3, nrmlz, ntr, raw, r3, t0, t1); break;
            default: launch_tile_import<3>(h, h->fin_stream, o2, o3, nrmlz, ntr, raw, r3, t0, t1); break;
          }
          h->launches++;
        }
      }
      CU(cudaGetLastError());
      CU(cudaEventRecord(h->ev_fin, writer ? h->fin_stream : h->comm_stream));
      CU(cudaStreamWaitEvent(h->compute, h->ev_fin, 0));
      if ((rc = mark_main(h))) return rc;
    } else {
    const size_t row_recs = (size_t)g.im * g.km;
    unsigned tprev = 0;
    size_t rprev = 0;
    for (int c = 0; c < NC; ++c) {
      const int r0 = (int)((long long)nrows * c / NC), r1 = (int)((long long)nrows * (c + 1) / NC);
      const bool last = c == NC - 1;
      const size_t rec_end = last ? h->nrec : (size_t)r1 * row_recs;   // the last chunk also carries the 2-D boxes
      if (staged) {
        const unsigned t1 = last ? h->sg.ntiles : (unsigned)std::min((size_t)h->sg.ntiles, (rec_end + T - 1) / T);
        launch_tiles_nr(h, tprev, t1, pipe_free_sms);   // includes a tile straddling into the next chunk: complete before its reduce
        tprev = t1;
        h->launches++;
      }
      CU(cudaEventRecord(h->ev_acc[c], h->compute));
      CU(cudaStreamWaitEvent(h->comm_stream, h->ev_acc[c], 0));
      if (rec_end > rprev && h->nranks > 1) {
        double *seg = g.acc + rprev * g.rs;
        int e = g_nccl.Reduce(seg, seg, (rec_end - rprev) * g.rs, NCCL_FLOAT64, NCCL_SUM, reduce_root, h->comm, h->comm_stream);
        if (e != 0) return fail(h, GRITAS_B200_ERR_NCCL, "ncclReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "?");
      }
      rprev = rec_end;
      CU(cudaEventRecord(h->ev_red[c], h->comm_stream));
      if (writer) {
        CU(cudaStreamWaitEvent(h->fin_stream, h->ev_red[c], 0));
        rc = launch_finalize(h, h->fin_stream, dst, raw, raw3, xpl, nrmlz, ntr, r0, r1, last);
        if (rc) return rc;
      }
    }
    CU(cudaGetLastError());
    h->staged_upper = 0;
    if ((rc = touch_records(h))) return rc;
    CU(cudaEventRecord(h->ev_fin, writer ? h->fin_stream : h->comm_stream));
    CU(cudaStreamWaitEvent(h->compute, h->ev_fin, 0));
    if ((rc = mark_main(h))) return rc;
    }
  }
  for (int a = 0; a < 10; ++a) {
    if (dev_out) dev_out[a] = dst[a];
    else if (bounce[a]) {
      if (mem_kind(user[a]) == 0 && sz[a] * sizeof(double) >= ((size_t)1 << 20)) {
        int rc = copy_out_pageable(h, user[a], dst[a], sizeof(double) * sz[a], h->compute);
        if (rc) return rc;
      } else {
        CU(cudaMemcpyAsync(user[a], dst[a], sizeof(double) * sz[a], cudaMemcpyDeviceToHost, h->compute));
        h->d2h += sizeof(double) * sz[a];
      }
    }
  }
  if (piped || h->det_tiled) return drain(h, nullptr);     // (deterministic tiled path: the tile kernel may have lost the order of a box)
  CU(cudaStreamSynchronize(h->compute));
  return 0;
}

int gritas_b200_norm(gritas_b200_handle h, int nrmlz, int ntresh, double *bias_2d, double *rtms_2d, double *stdv_2d,
                     double *xcov_2d, double *nobs_2d, double *bias_3d, double *rtms_3d, double *stdv_3d,
                     double *xcov_3d, double *nobs_3d) {
  int rc = check_handle(h);
  if (rc) return rc;
  if ((rc = settle(h, fused_norm_ok(h), nullptr))) return rc;
  double *const u[10] = {bias_2d, rtms_2d, stdv_2d, xcov_2d, nobs_2d, bias_3d, rtms_3d, stdv_3d, xcov_3d, nobs_3d};
  return finalize_to(h, 0, nrmlz ? 1 : 0, ntresh, u, nullptr);
}

// GrITAS_3CH_Norm_ (m_gritas_grids.f:1178-1341; gritas.f:813-818, -tch): three residual series; var_* receive the variances
// (no square root), xcov_* the three-cornered-hat estimates.
int gritas_b200_norm_3ch(gritas_b200_handle h, int nrmlz, int ntresh, double *bias_2d, double *rtms_2d, double *var_2d,
                         double *xcov_2d, double *nobs_2d, double *bias_3d, double *rtms_3d, double *var_3d,
                         double *xcov_3d, double *nobs_3d) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (h->g.nr != 3) return fail(h, GRITAS_B200_ERR_NR, "GrITAS_3CH_Norm_ needs three residual series (nr = %d)", h->g.nr);
  if ((rc = settle(h, fused_norm_ok(h), nullptr))) return rc;
  double *const u[10] = {bias_2d, rtms_2d, var_2d, xcov_2d, nobs_2d, bias_3d, rtms_3d, var_3d, xcov_3d, nobs_3d};
  return finalize_to(h, 3, nrmlz ? 1 : 0, ntresh, u, nullptr);
}

int gritas_b200_norm_device(gritas_b200_handle h, int nrmlz, int ntresh, const double **bias_2d, const double **rtms_2d,
                            const double **stdv_2d, const double **xcov_2d, const double **nobs_2d,
                            const double **bias_3d, const double **rtms_3d, const double **stdv_3d,
                            const double **xcov_3d, const double **nobs_3d) {
  int rc = check_handle(h);
  if (rc) return rc;
  if ((rc = settle(h, fused_norm_ok(h), nullptr))) return rc;
  double *const u[10] = {};
  const double *d[10] = {};
  rc = finalize_to(h, 0, nrmlz ? 1 : 0, ntresh, u, d);
  if (rc) return rc;
  const double **outp[10] = {bias_2d, rtms_2d, stdv_2d, xcov_2d, nobs_2d, bias_3d, rtms_3d, stdv_3d, xcov_3d, nobs_3d};
  for (int a = 0; a < 10; ++a)
    if (outp[a]) *outp[a] = d[a];
  return 0;
}

// Multi-rank Norm: tile flush, reduction of the partial accumulators to `root` and finalize, pipelined in
// chunks; only the root's output arrays are written.  Collective over the communicator.
int gritas_b200_norm_reduce(gritas_b200_handle h, int root, int nrmlz, int ntresh, double *bias_2d, double *rtms_2d,
                            double *stdv_2d, double *xcov_2d, double *nobs_2d, double *bias_3d, double *rtms_3d,
                            double *stdv_3d, double *xcov_3d, double *nobs_3d) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (root < 0 || root >= h->nranks) return fail(h, GRITAS_B200_ERR_ARG, "bad root");
  CU(cudaStreamSynchronize(h->copy));
  double *const u[10] = {bias_2d, rtms_2d, stdv_2d, xcov_2d, nobs_2d, bias_3d, rtms_3d, stdv_3d, xcov_3d, nobs_3d};
  if ((h->nranks == 1 || !h->comm) && !force_pipeline()) {
    if ((rc = settle(h, fused_norm_ok(h), nullptr))) return rc;
    return finalize_to(h, 0, nrmlz ? 1 : 0, ntresh, u, nullptr);
  }
  return finalize_to(h, 0, nrmlz ? 1 : 0, ntresh, u, nullptr, root);
}

int gritas_b200_norm_reduce_device(gritas_b200_handle h, int root, int nrmlz, int ntresh, const double **outs10) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (root < 0 || root >= h->nranks || !outs10) return fail(h, GRITAS_B200_ERR_ARG, "bad root / outs");
  CU(cudaStreamSynchronize(h->copy));
  double *const u[10] = {};
  if ((h->nranks == 1 || !h->comm) && !force_pipeline()) {
    if ((rc = settle(h, fused_norm_ok(h), nullptr))) return rc;
    return finalize_to(h, 0, nrmlz ? 1 : 0, ntresh, u, outs10);
  }
  return finalize_to(h, 0, nrmlz ? 1 : 0, ntresh, u, outs10, root);
}

int gritas_b200_get_raw(gritas_b200_handle h, double *bias_2d, double *rtms_2d, double *xcov_2d, double *nobs_2d,
                        double *bias_3d, double *rtms_3d, double *xcov_3d, double *nobs_3d) {
  int rc = check_handle(h);
  if (rc) return rc;
  if ((rc = settle(h, fused_norm_ok(h), nullptr))) return rc;
  double *const u[10] = {bias_2d, rtms_2d, nullptr, xcov_2d, nobs_2d, bias_3d, rtms_3d, nullptr, xcov_3d, nobs_3d};
  return finalize_to(h, 1, 0, 0, u, nullptr);
}

// Scorecard of one synoptic time (scores1_ / scores2_ / ave_nomiss, m_gritas_grids.f:1526-1826, minus the CSV writing of the
// un-vendored m_obs_stats): Norm with nrmlz = .true. (:1632-1634) on the device, then the regional averages; only the
// (mregs, km, l2d + l3d, 3) real(4) table crosses PCIe, the finalized grids stay on the device.
int gritas_b200_scores(gritas_b200_handle h, int nregs, const int32_t *regindx, int lon0, int lon1, float *collect_stats) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (h->g.nr != 1) return fail(h, GRITAS_B200_ERR_NR, "GrITAS_scores_: incorrec number of residuals, aborting ...");   // :1704-1706
  if (nregs < 1 || nregs > 64 || !regindx || !collect_stats) return fail(h, GRITAS_B200_ERR_ARG, "scores: bad arguments");
  for (int r = 0; r < nregs; ++r)
    if (regindx[2 * r] < 1 || regindx[2 * r + 1] > h->g.jm) return fail(h, GRITAS_B200_ERR_ARG, "scores: region %d outside the grid", r + 1);
  if (lon0 < 1 || lon1 > h->g.im) return fail(h, GRITAS_B200_ERR_ARG, "scores: longitude range outside the grid");
  if ((rc = settle(h, fused_norm_ok(h), nullptr))) return rc;
  double *const u[10] = {};
  const double *d[10] = {};
  if ((rc = finalize_to(h, 0, 1, -1, u, d))) return rc;
  const size_t total = (size_t)nregs * h->g.km * (h->g.l2d + h->g.l3d) * 3;
  // scratch after the output planes: region table + result
  void *scr = nullptr;
  CU(cudaMalloc(&scr, sizeof(int) * 2 * nregs + sizeof(float) * total + 64));
  int *d_reg = (int *)scr;
  float *d_out = (float *)((char *)scr + ((sizeof(int) * 2 * nregs + 15) & ~(size_t)15));
  CU(cudaMemcpyAsync(d_reg, regindx, sizeof(int) * 2 * nregs, cudaMemcpyHostToDevice, h->compute));
  if (mem_kind(collect_stats) == 2) CU(cudaMemcpyAsync(d_out, collect_stats, sizeof(float) * total, cudaMemcpyDeviceToDevice, h->compute));
  else CU(cudaMemcpyAsync(d_out, collect_stats, sizeof(float) * total, cudaMemcpyHostToDevice, h->compute));   // untouched entries keep the caller's values
  const Planes3 a2 = {{d[0], d[1], d[2]}}, a3 = {{d[5], d[6], d[7]}};
  k_scores<<<(int)((total + 127) / 128), 128, 0, h->compute>>>(h->g, a2, a3, d[4], d[9], d_reg, nregs, lon0, lon1, d_out);
  h->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(collect_stats, d_out, sizeof(float) * total, cudaMemcpyDefault, h->compute));
  h->d2h += sizeof(float) * total;
  CU(cudaStreamSynchronize(h->compute));
  CU(cudaFree(scr));
  return 0;
}

// Region rows of the scorecard as init_ derives them (m_gritas_grids.f:163-175): for every region the last latitude row
// at or below its southern bound and the last one at or below its northern bound (glat(j) = -90 + (j-1) dLat).
int gritas_b200_score_regions(int jm, int nregs, const double *regbounds, int32_t *regindx) {
  if (jm < 2 || nregs < 1 || !regbounds || !regindx) return 1;
  const double dlat = 180.0 / (double)(jm - 1);
  for (int i = 0; i < nregs; ++i) {
    regindx[2 * i] = 0;
    regindx[2 * i + 1] = 0;
    for (int j = 1; j <= jm; ++j) {
      const double glat = -90.0 + (double)(j - 1) * dlat;
      if (glat <= regbounds[2 * i]) regindx[2 * i] = j;
      if (glat > 0.0 && glat <= regbounds[2 * i + 1]) regindx[2 * i + 1] = j;
      if (glat < 0.0 && glat <= regbounds[2 * i + 1]) regindx[2 * i + 1] = j;
    }
  }
  return 0;
}

// GrITAS_MinMax results (m_gritas_grids.f:1019-1167): minv (AMISS where nothing fell), maxv (-AMISS where
// nothing fell; gritas.f:997-1000 turns those into AMISS afterwards) and the counts, reference layout.
int gritas_b200_get_minmax(gritas_b200_handle h, double *minv_2d, double *maxv_2d, double *nobs_2d, double *minv_3d,
                           double *maxv_3d, double *nobs_3d) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (!h->minmax) return fail(h, GRITAS_B200_ERR_ARG, "not a MinMax handle (GRITAS_B200_FLAG_MINMAX)");
  if ((rc = gritas_b200_sync(h, nullptr))) return rc;
  double *const u[10] = {minv_2d, nullptr, maxv_2d, nullptr, nobs_2d, minv_3d, nullptr, maxv_3d, nullptr, nobs_3d};
  return finalize_to(h, 2, 0, 0, u, nullptr);
}

// gritas_maskout fused into the classify step: observations whose box has maskfld(i,j) < 0.99 are flagged
// (m_gritas_masks.F90:164-183).  maskfld: (im, jm) column-major on the handle's grid; NULL removes the mask.
int gritas_b200_set_mask(gritas_b200_handle h, const double *maskfld) {
  int rc = check_handle(h);
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->compute));
  if ((rc = join_workers(h))) return rc;
  CU(cudaStreamSynchronize(h->compute));
  if (!maskfld) {
    h->g.mask = nullptr;
    return 0;
  }
  const size_t bytes = sizeof(double) * (size_t)h->g.im * h->g.jm;
  if (!h->d_mask) CU(cudaMalloc(&h->d_mask, bytes));
  CU(cudaMemcpy(h->d_mask, maskfld, bytes, cudaMemcpyDefault));
  h->g.mask = h->d_mask;
  return 0;
}

int gritas_b200_add_raw(gritas_b200_handle h, const double *bias_2d, const double *rtms_2d, const double *xcov_2d,
                        const double *nobs_2d, const double *bias_3d, const double *rtms_3d, const double *xcov_3d,
                        const double *nobs_3d) {
  int rc = check_handle(h);
  if (rc) return rc;
  if ((rc = flush_stage(h))) return rc;
  if ((rc = join_workers(h))) return rc;
  if ((rc = touch_records(h))) return rc;
  const GridDesc &g = h->g;
  const int nr = g.nr;
  const size_t n2 = g.nbox2, n3 = g.nbox3;
  const double *src[8] = {bias_2d, rtms_2d, xcov_2d, nobs_2d, bias_3d, rtms_3d, xcov_3d, nobs_3d};
  const size_t sz[8] = {n2 * nr, n2 * nr, n2, n2, n3 * nr, n3 * nr, n3, n3};
  const double *dev[8] = {};
  size_t need = 0;
  for (int a = 0; a < 8; ++a)
    if (src[a] && sz[a] && mem_kind(src[a]) != 2) need += sz[a];
  if (need) {
    rc = grow_dev(h, (void **)&h->out, &h->out_cap, sizeof(double) * need);
    if (rc) return rc;
  }
  size_t off = 0;
  for (int a = 0; a < 8; ++a) {
    if (!src[a] || !sz[a]) continue;
    if (mem_kind(src[a]) == 2) { dev[a] = src[a]; continue; }
    CU(cudaMemcpyAsync(h->out + off, src[a], sizeof(double) * sz[a], cudaMemcpyHostToDevice, h->compute));
    h->h2d += sizeof(double) * sz[a];
    dev[a] = h->out + off;
    off += sz[a];
  }
  for (int is3d = 0; is3d < 2; ++is3d) {
    const size_t nb = is3d ? n3 : n2;
    if (!nb) continue;
    const double *const *d = dev + 4 * is3d;
    if (!d[0] && !d[1] && !d[2] && !d[3]) continue;
    OutPlanes in = {(double *)d[0], (double *)d[1], nullptr, (double *)d[2], (double *)d[3], 1};
    const int blocks = grid_for(nb, 256, 8);
    if (nr == 1) k_add_raw<1><<<blocks, 256, 0, h->compute>>>(g, in, is3d);
    else if (nr == 2) k_add_raw<2><<<blocks, 256, 0, h->compute>>>(g, in, is3d);
    else k_add_raw<3><<<blocks, 256, 0, h->compute>>>(g, in, is3d);
    h->launches++;
  }
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->compute));
  return mark_main(h);
}

// IndexSet + the eight stable sorts on device-resident columns (order: qcexcl lat lon lev time kt ks kx); *fin = the
// permutation (0-based) in one of the handle's index buffers.
static int presort_core(gritas_b200_handle h, int nobs, const void *const dev[8], int **fin_out) {
  const size_t n = (size_t)nobs;
  const size_t esz[8] = {4, 8, 8, 8, 4, 4, 4, 4};
  cudaStream_t sm = h->compute;
  int rc;
  for (int b = 0; b < 2; ++b) {
    if ((rc = grow_dev(h, &h->ps_key[b], &h->ps_key_cap[b], 8 * n))) return rc;
    if ((rc = grow_dev(h, &h->ps_idx[b], &h->ps_idx_cap[b], 4 * n))) return rc;
  }
  size_t t32 = 0, t64 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t32, (unsigned *)nullptr, (unsigned *)nullptr, (int *)nullptr, (int *)nullptr, nobs, 0, 32, sm);
  cub::DeviceRadixSort::SortPairs(nullptr, t64, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int *)nullptr,
                                  (int *)nullptr, nobs, 0, 64, sm);
  if ((rc = grow_dev(h, &h->cub_tmp, &h->cub_cap, std::max(t32, t64)))) return rc;
  const int nblk = grid_for(n, 256, 8);
  int cur = 0;
  k_iota<<<nblk, 256, 0, sm>>>((int *)h->ps_idx[0], nobs);
  for (int p = 0; p < 8; ++p) {
    int *idx = (int *)h->ps_idx[cur], *idx2 = (int *)h->ps_idx[cur ^ 1];
    size_t tmp = h->cub_cap;
    if (esz[p] == 4) {
      k_sortkey_i32<<<nblk, 256, 0, sm>>>((const int *)dev[p], idx, (unsigned *)h->ps_key[0], nobs);
      CU(cub::DeviceRadixSort::SortPairs(h->cub_tmp, tmp, (unsigned *)h->ps_key[0], (unsigned *)h->ps_key[1], idx, idx2, nobs, 0, 32, sm));
    } else {
      k_sortkey_f64<<<nblk, 256, 0, sm>>>((const double *)dev[p], idx, (unsigned long long *)h->ps_key[0], nobs);
      CU(cub::DeviceRadixSort::SortPairs(h->cub_tmp, tmp, (unsigned long long *)h->ps_key[0], (unsigned long long *)h->ps_key[1], idx,
                                         idx2, nobs, 0, 64, sm));
    }
    cur ^= 1;
    h->launches += 1 + (esz[p] == 4 ? 9 : 17);   // key kernel + radix passes (approx.)
  }
  *fin_out = (int *)h->ps_idx[cur];
  return 0;
}

// ---- pre-sort (gritas.f:499-508) ------------------------------------------------------------------
// IndexSet followed by eight stable ascending IndexSorts -- qcexcl, lat, lon, lev, time, kt, ks, kx, in this order, so
// that the final order is (kx, ks, kt, time, lev, lon, lat, qcexcl, original position).  The reference spends seconds
// per synoptic time here (index merge sorts on the host); on the device it is eight LSD radix-sort passes over
// (order-preserving key, index) pairs.  `is` receives the permutation (index_base 0 or 1: Fortran's `is` is 1-based);
// the column gathers of gritas.f:510-523 stay with the caller.  Columns and `is` may live on the host or the device.
// The result is the unique stable order, hence identical to the host sort for any input without NaNs (the merge
// compares with `<`: -0.0 and +0.0 are equal there and here; the place of NaNs is unspecified in both).
int gritas_b200_presort(gritas_b200_handle h, int nobs, const int32_t *qcexcl, const double *lat, const double *lon,
                        const double *lev, const int32_t *time, const int32_t *kt, const int32_t *ks, const int32_t *kx,
                        int index_base, int32_t *is) {
  // gritas.f sorts while reading a file, before the first GrITAS_Accum has created the handle: a NULL handle uses a
  // process-wide scratch context (a stream and buffers on the current device)
  static gritas_b200_handle scratch = nullptr;
  if (!h) {
    if (!scratch) {
      int dev = 0;
      if (cudaGetDevice(&dev) != cudaSuccess) return GRITAS_B200_ERR_CUDA;
      scratch = new gritas_b200_ctx();
      scratch->device = dev;
      if (cudaStreamCreateWithFlags(&scratch->compute, cudaStreamNonBlocking) != cudaSuccess) {
        delete scratch;
        scratch = nullptr;
        return GRITAS_B200_ERR_CUDA;
      }
    }
    h = scratch;
  }
  int rc = check_handle(h);
  if (rc) return rc;
  if (nobs < 0 || !is || !qcexcl || !lat || !lon || !lev || !time || !kt || !ks || !kx || (index_base != 0 && index_base != 1))
    return fail(h, GRITAS_B200_ERR_ARG, "presort: bad arguments");
  if (nobs == 0) return 0;
  const size_t n = (size_t)nobs;
  cudaStream_t sm = h->compute;
  // staging of host columns: 3 fp64 + 5 int32 columns
  const void *src[8] = {qcexcl, lat, lon, lev, time, kt, ks, kx};
  const size_t esz[8] = {4, 8, 8, 8, 4, 4, 4, 4};
  size_t need = 0;
  for (int c = 0; c < 8; ++c)
    if (mem_kind(src[c]) != 2) need += (esz[c] * n + 15) & ~(size_t)15;
  if (need && (rc = grow_dev(h, &h->ps_cols, &h->ps_cols_cap, need))) return rc;
  const void *dev[8];
  size_t off = 0;
  for (int c = 0; c < 8; ++c) {
    if (mem_kind(src[c]) == 2) { dev[c] = src[c]; continue; }
    void *d = (char *)h->ps_cols + off;
    CU(cudaMemcpyAsync(d, src[c], esz[c] * n, cudaMemcpyHostToDevice, sm));
    h->h2d += esz[c] * n;
    dev[c] = d;
    off += (esz[c] * n + 15) & ~(size_t)15;
  }
  int *fin = nullptr;
  if ((rc = presort_core(h, nobs, dev, &fin))) return rc;
  const int cur = (fin == (int *)h->ps_idx[0]) ? 0 : 1;
  const int nblk = grid_for(n, 256, 8);
  if (mem_kind(is) == 2) {
    k_add_base<<<nblk, 256, 0, sm>>>(fin, is, index_base, nobs);
  } else {
    int *out = (int *)h->ps_idx[cur ^ 1];
    k_add_base<<<nblk, 256, 0, sm>>>(fin, out, index_base, nobs);
    CU(cudaMemcpyAsync(is, out, 4 * n, cudaMemcpyDeviceToHost, sm));
    h->d2h += 4 * n;
  }
  h->launches += 2;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(sm));
  return 0;
}

// The whole per-file preamble of gritas.f on the device: the 8-key pre-sort (gritas.f:499-508), the column gathers
// (gritas.f:510-523), residual selection (gritas.f:562-704), the QC filter (gritas.f:710-739) and GrITAS_Accum_.  Columns come
// in as ODS_Get left them (unsorted, host or device); they are copied to the device once, sorted and gathered there, and the
// accumulate kernels read the gathered copies.  In deterministic mode the sums are then formed in the reference's order.
// ob_flag_out (may be NULL) is in SORTED order, as the reference's ob_flag is; is_out (may be NULL) receives the 1-based
// permutation.
int gritas_b200_accum_odsx_sorted(gritas_b200_handle h, int nobs, const gritas_b200_ods *o, const int32_t *ks, int iopt, int options,
                                  int scale_res, int ioptn, int t1, int t2, double lona, double lonb, int x_passive, int x_pre_bad,
                                  int32_t *ob_flag_out, int32_t *is_out, double *bad_lev) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (!o || nobs < 0) return fail(h, GRITAS_B200_ERR_ARG, "accum_odsx_sorted: bad arguments");
  if (nobs == 0) return gritas_b200_accum_odsx(h, 0, o, iopt, options, scale_res, ioptn, t1, t2, lona, lonb, x_passive, x_pre_bad, ob_flag_out, bad_lev);
  if (!o->lat || !o->lon || !o->lev || !o->time || !o->kt || !o->kx || !o->qcexcl || !ks)
    return fail(h, GRITAS_B200_ERR_ARG, "accum_odsx_sorted: the pre-sort needs qcexcl, lat, lon, lev, time, kt, ks, kx");
  const size_t n = (size_t)nobs;
  // the accumulate of an earlier call may still read the gathered columns of that call
  if ((rc = join_workers(h))) return rc;
  CU(cudaStreamSynchronize(h->compute));
  const void *fsrc[8] = {o->lat, o->lon, o->lev, o->obs, o->omf, o->oma, o->xvec, o->xm};
  const void *isrc[5] = {o->time, o->kt, o->kx, o->qcexcl, ks};
  const size_t f8 = (8 * n + 255) & ~(size_t)255, i4 = (4 * n + 255) & ~(size_t)255;
  size_t need_in = 0, need_out = 0;
  for (const void *p : fsrc) if (p) { need_out += f8; if (mem_kind(p) != 2) need_in += f8; }
  for (const void *p : isrc) if (p) { need_out += i4; if (mem_kind(p) != 2) need_in += i4; }
  if ((rc = grow_dev(h, &h->ps_cols, &h->ps_cols_cap, need_in + 256))) return rc;
  if ((rc = grow_dev(h, &h->ps_sorted, &h->ps_sorted_cap, need_out + 256))) return rc;
  const void *fdev[8] = {}, *idev[5] = {};
  size_t off = 0;
  auto upload = [&](const void *p, size_t bytes, size_t slot, const void **dev) -> int {
    if (!p) return 0;
    if (mem_kind(p) == 2) { *dev = p; return 0; }
    void *d = (char *)h->ps_cols + off;
    CU(cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, h->compute));
    h->h2d += bytes;
    *dev = d;
    off += slot;
    return 0;
  };
  for (int q = 0; q < 8; ++q) if ((rc = upload(fsrc[q], 8 * n, f8, &fdev[q]))) return rc;
  for (int q = 0; q < 5; ++q) if ((rc = upload(isrc[q], 4 * n, i4, &idev[q]))) return rc;
  const void *keys[8] = {idev[3], fdev[0], fdev[1], fdev[2], idev[0], idev[1], idev[4], idev[2]};   // qcexcl lat lon lev time kt ks kx
  int *fin = nullptr;
  if ((rc = presort_core(h, nobs, keys, &fin))) return rc;
  GatherCols g = {};
  double *fs[8] = {};
  int *is5[5] = {};
  size_t so = 0;
  for (int q = 0; q < 8; ++q)
    if (fdev[q]) { fs[q] = (double *)((char *)h->ps_sorted + so); so += f8; g.f_src[g.nf] = (const double *)fdev[q]; g.f_dst[g.nf] = fs[q]; ++g.nf; }
  for (int q = 0; q < 5; ++q)
    if (idev[q]) { is5[q] = (int *)((char *)h->ps_sorted + so); so += i4; g.i_src[g.ni] = (const int *)idev[q]; g.i_dst[g.ni] = is5[q]; ++g.ni; }
  const int nblk = grid_for(n, 256, 8);
  k_gather_cols<<<nblk, 256, 0, h->compute>>>(g, fin, nobs);
  h->launches++;
  if (is_out) {
    const int cur = (fin == (int *)h->ps_idx[0]) ? 0 : 1;
    if (mem_kind(is_out) == 2) {
      k_add_base<<<nblk, 256, 0, h->compute>>>(fin, is_out, 1, nobs);
    } else {
      int *out = (int *)h->ps_idx[cur ^ 1];
      k_add_base<<<nblk, 256, 0, h->compute>>>(fin, out, 1, nobs);
      CU(cudaMemcpyAsync(is_out, out, 4 * n, cudaMemcpyDeviceToHost, h->compute));
      h->d2h += 4 * n;
    }
    h->launches++;
  }
  CU(cudaGetLastError());
  if ((rc = mark_main(h))) return rc;           // the accumulate (possibly on a worker stream) starts after the gathers
  gritas_b200_ods sorted = {fs[0], fs[1], fs[2], fs[3], fs[4], fs[5], fs[6], fs[7], is5[0], is5[1], is5[2], is5[3]};
  return gritas_b200_accum_odsx(h, nobs, &sorted, iopt, options, scale_res, ioptn, t1, t2, lona, lonb, x_passive, x_pre_bad,
                                ob_flag_out, bad_lev);
}

// ---- multi-GPU ----------------------------------------------------------------------------------
int gritas_b200_nccl_unique_id(void *id128) {
  if (!id128) return GRITAS_B200_ERR_ARG;
  if (!g_nccl.load()) return GRITAS_B200_ERR_NCCL;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != 0) return GRITAS_B200_ERR_NCCL;
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int gritas_b200_comm_init(gritas_b200_handle h, int nranks, int rank, const void *id128) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (nranks < 1 || rank < 0 || rank >= nranks || !id128) return fail(h, GRITAS_B200_ERR_ARG, "bad rank/nranks");
  if (!g_nccl.load()) return fail(h, GRITAS_B200_ERR_NCCL, "libnccl.so.2 not loadable: %s", dlerror());
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  int e = g_nccl.CommInitRank(&h->comm, nranks, id, rank);
  if (e != 0) return fail(h, GRITAS_B200_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "?");
  h->nranks = nranks;
  h->rank = rank;
  // What shapes the collectives of the multi-rank Norm (tiled or not, tile size, list capacity, pipeline settings from the
  // GRITAS_B200_* environment) must be the same on every rank: a max-reduction of (sig, -sig) tells every rank alike.
  if (nranks > 1) {
    const char *e_c = getenv("GRITAS_B200_PIPE_CHUNKS"), *e_r = getenv("GRITAS_B200_COMPACT_REDUCE");
    const double sig = (double)((h->staging ? 1 : 0) + 2 * (h->fused_norm ? 1 : 0) + 4 * (h->det_tiled ? 1 : 0)) + 8.0 * h->g.nr +
                       64.0 * (h->staging ? h->sg.tile_shift : 0) + 4096.0 * (e_c ? atoi(e_c) : 8) + 65536.0 * (e_r ? atoi(e_r) : 1) +
                       1048576.0 * (double)((h->staging ? h->sg.cap : 0) % 1000003u) + 1.0e13 * (double)(h->nrec % 100003u);
    double host[2] = {sig, -sig};
    if (!h->bar_buf) CU(cudaMalloc(&h->bar_buf, 64));
    CU(cudaMemcpyAsync(h->bar_buf, host, sizeof(host), cudaMemcpyHostToDevice, h->compute));
    e = g_nccl.AllReduce(h->bar_buf, h->bar_buf, 2, NCCL_FLOAT64, /*ncclMax*/ 2, h->comm, h->compute);
    if (e != 0) return fail(h, GRITAS_B200_ERR_NCCL, "ncclAllReduce (settings check): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "?");
    CU(cudaMemcpyAsync(host, h->bar_buf, sizeof(host), cudaMemcpyDeviceToHost, h->compute));
    CU(cudaStreamSynchronize(h->compute));
    CU(cudaMemsetAsync(h->bar_buf, 0, 64, h->compute));
    if (host[0] != -host[1])
      return fail(h, GRITAS_B200_ERR_ARG, "comm_init: the ranks were created with different grids, modes or GRITAS_B200_* settings "
                                          "(the collectives of the multi-rank Norm would not match)");
  }
  return 0;
}

int gritas_b200_allreduce(gritas_b200_handle h) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (h->nranks == 1 && !h->comm) return 0;
  if (!h->comm) return fail(h, GRITAS_B200_ERR_NCCL, "comm_init not called");
  if (h->minmax) return fail(h, GRITAS_B200_ERR_ARG, "MinMax handles are not summed across ranks");
  CU(cudaStreamSynchronize(h->copy));
  if ((rc = flush_stage(h))) return rc;
  if ((rc = join_workers(h))) return rc;      // flush_stage returns early when nothing is parked: K1 launches may still be queued
  if ((rc = touch_records(h))) return rc;
  int e = g_nccl.AllReduce(h->g.acc, h->g.acc, h->nrec * h->g.rs, NCCL_FLOAT64, NCCL_SUM, h->comm, h->compute);
  if (e != 0) return fail(h, GRITAS_B200_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "?");
  return mark_main(h);                        // later K1 launches (direct RED on a full list) must not overlap the in-place reduction
}

// ---- sharded multi-rank Norm over peer memory --------------------------------------------------------
// The month is sharded f mod nranks (gritas.f:389-393 is the file loop each rank runs on its share); every rank parks
// its accepted observations in its own tile lists.  GrITAS_Norm (gritas.f:885-887) then needs the sums over all ranks.
// Instead of reducing the sum grids (1.3 GB at cfg2 whatever the rank count) rank r takes 1/nranks of the rows and its
// tile kernel pulls the parked entries of those tiles from every rank's lists through NVLink peer mappings
// (3.9 GB / nranks per rank), finalizes them and writes its slab of the output arrays.  No collective on the data path;
// the two stream-ordered barriers are one-element all-reduces.
namespace {

struct PeerBlob {
  uint64_t magic;
  int32_t pid, device, staging, nr, tile_shift, mode;
  uint32_t ntiles, cap;
  uint64_t nrec;
  uint64_t raw_lists, raw_fill, raw_acc, raw_red;
  int32_t ipc_ok, pad;
  cudaIpcMemHandle_t ipc_lists, ipc_fill, ipc_acc, ipc_red;
  char host[64];
};
static_assert(sizeof(PeerBlob) <= GRITAS_B200_PEER_BLOB_BYTES, "PeerBlob must fit the ABI blob");
constexpr uint64_t PEER_MAGIC = 0x6772697461733262ull;

struct Slab {
  size_t lo, hi;                    // records [lo, hi)
  int jl3_0, jl3_1, jl2_0, jl2_1;   // rows jl = j + jm*l of the 3-D / 2-D grids
};

// Rows are never split: the slab of rank r starts at the row boundary at or below nrec * r / n.
size_t slab_bound(const GridDesc &g, size_t nrec, int r, int n) {
  if (r <= 0) return 0;
  if (r >= n) return nrec;
  const size_t x = nrec * (size_t)r / (size_t)n;
  const size_t row3 = (size_t)g.im * g.km;
  if (x <= (size_t)g.nbox3) return x / row3 * row3;
  return (size_t)g.nbox3 + (x - g.nbox3) / g.im * g.im;
}

Slab slab_from(const GridDesc &g, size_t lo, size_t hi) {
  Slab s;
  s.lo = lo;
  s.hi = hi;
  const size_t row3 = (size_t)g.im * g.km, n3 = g.nbox3;
  s.jl3_0 = (int)(std::min(s.lo, n3) / row3);
  s.jl3_1 = (int)(std::min(s.hi, n3) / row3);
  s.jl2_0 = (int)((std::max(s.lo, n3) - n3) / g.im);
  s.jl2_1 = (int)((std::max(s.hi, n3) - n3) / g.im);
  return s;
}

Slab slab_of(const GridDesc &g, size_t nrec, int r, int n) {
  Slab s;
  s.lo = slab_bound(g, nrec, r, n);
  s.hi = slab_bound(g, nrec, r + 1, n);
  const size_t row3 = (size_t)g.im * g.km, n3 = g.nbox3;
  s.jl3_0 = (int)(std::min(s.lo, n3) / row3);
  s.jl3_1 = (int)(std::min(s.hi, n3) / row3);
  s.jl2_0 = (int)((std::max(s.lo, n3) - n3) / g.im);
  s.jl2_1 = (int)((std::max(s.hi, n3) - n3) / g.im);
  return s;
}

// Every rank has queued everything before this point of its compute stream (stream-ordered; no host wait).
int stream_barrier(gritas_b200_handle h) {
  if (!h->comm || h->nranks < 2) return 0;
  if (!h->bar_buf) {
    CU(cudaMalloc(&h->bar_buf, 64));
    CU(cudaMemsetAsync(h->bar_buf, 0, 64, h->compute));
  }
  int e = g_nccl.AllReduce(h->bar_buf, h->bar_buf, 1, NCCL_FLOAT64, NCCL_SUM, h->comm, h->compute);
  if (e != 0) return fail(h, GRITAS_B200_ERR_NCCL, "ncclAllReduce (barrier): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "?");
  return 0;
}

// Device -> host (or in place) copy of this rank's slab of reference-layout array `a` (see finalize_to for the order).
int copy_slab(gritas_b200_handle h, const Slab &sl, int a, int planes, double *user, const double *dev) {
  const GridDesc &g = h->g;
  const bool is3 = a >= 5;
  if (a == 4 || a == 9) planes = 1;
  if (is3) {
    const size_t pitch = sizeof(double) * (size_t)g.im * g.jm;
    for (int ir = 0; ir < planes; ++ir)
      for (int l = sl.jl3_0 / g.jm; l <= (sl.jl3_1 - 1) / g.jm && sl.jl3_1 > sl.jl3_0; ++l) {
        const int ja = std::max(sl.jl3_0, l * g.jm) - l * g.jm, jb = std::min(sl.jl3_1, (l + 1) * g.jm) - l * g.jm;
        if (jb <= ja) continue;
        const size_t off = (size_t)g.im * ((size_t)ja + (size_t)g.jm * ((size_t)g.km * ((size_t)l + (size_t)g.l3d * ir)));
        const size_t width = sizeof(double) * (size_t)g.im * (size_t)(jb - ja);
        CU(cudaMemcpy2DAsync(user + off, pitch, dev + off, pitch, width, (size_t)g.km, cudaMemcpyDeviceToHost, h->compute));
        h->d2h += width * g.km;
      }
  } else {
    for (int ir = 0; ir < planes; ++ir) {
      if (sl.jl2_1 <= sl.jl2_0) continue;
      const size_t off = (size_t)g.im * (size_t)sl.jl2_0 + (size_t)g.nbox2 * ir;
      const size_t bytes = sizeof(double) * (size_t)g.im * (size_t)(sl.jl2_1 - sl.jl2_0);
      CU(cudaMemcpyAsync(user + off, dev + off, bytes, cudaMemcpyDeviceToHost, h->compute));
      h->d2h += bytes;
    }
  }
  return 0;
}

int finalize_sharded(gritas_b200_handle h, int nrmlz, int ntresh, double *const user[10], const double **dev_out) {
  const GridDesc &g = h->g;
  const int nr = g.nr;
  if (nr > 2) return fail(h, GRITAS_B200_ERR_NR, "Accum: could not handle more then 2 residuals");  // :555
  if (h->minmax) return fail(h, GRITAS_B200_ERR_ARG, "Norm on a MinMax handle");
  const size_t n2 = g.nbox2, n3 = g.nbox3;
  const int xpl = nr;
  const size_t sz[10] = {n2 * nr, n2 * nr, n2 * nr, n2 * xpl, n2, n3 * nr, n3 * nr, n3 * nr, n3 * xpl, n3};
  const bool raw3 = g.nlevs < 2;   // m_gritas_grids.f:607: the 3-D sums stay raw
  Slab sl = slab_of(g, h->nrec, h->prank, h->np);
  double *dst[10];
  bool bounce[10];
  size_t need = 0;
  for (int a = 0; a < 10; ++a) {
    bounce[a] = false;
    dst[a] = nullptr;
    if ((!user[a] && !dev_out) || sz[a] == 0 || (raw3 && a == 7)) continue;
    if (!dev_out && mem_kind(user[a]) == 2) dst[a] = user[a];
    else { bounce[a] = true; need += sz[a]; }
  }
  int rc;
  if (need) {
    if ((rc = grow_dev(h, (void **)&h->out, &h->out_cap, sizeof(double) * need))) return rc;
    size_t off = 0;
    for (int a = 0; a < 10; ++a)
      if (bounce[a]) { dst[a] = h->out + off; off += sz[a]; }
  }
  const int ntr = ntresh < 0 ? 0 : ntresh;
  if ((rc = join_workers(h))) return rc;
  if ((rc = stream_barrier(h))) return rc;            // every rank's accumulate work is complete
  PeerSet ps = {};
  ps.n = h->np;
  ps.own_lo = (unsigned)sl.lo;
  ps.own_hi = (unsigned)sl.hi;
  for (int s = 0; s < h->np; ++s) {
    ps.lists[s] = (const unsigned char *)h->peer_lists[s];
    ps.fill[s] = (const unsigned *)h->peer_fill[s];
    ps.acc[s] = (const double *)h->peer_acc[s];
    ps.red[s] = (const double *)h->peer_red[s];
  }
  if (h->staging) {
    const unsigned nwords = 2u * h->sg.ntiles + 4u;
    if ((rc = grow_dev(h, (void **)&h->fill_snap, &h->fill_snap_cap, sizeof(unsigned) * (size_t)nwords * h->np))) return rc;
    k_gather_words<<<148, 256, 0, h->compute>>>(h->fill_snap, ps, nwords);
    for (int s = 0; s < h->np; ++s) ps.fill[s] = h->fill_snap + (size_t)s * nwords;
    // slabs balanced by parked entries (every rank computes the same boundaries from the same snapshot)
    static const bool balance = !(getenv("GRITAS_B200_SLAB_BALANCE") && atoi(getenv("GRITAS_B200_SLAB_BALANCE")) == 0);
    if (balance) {
      if (!h->d_bounds) {
        CU(cudaMalloc(&h->d_bounds, sizeof(unsigned) * (MAX_PEERS + 2)));
        CU(cudaMallocHost(&h->h_bounds, sizeof(unsigned) * (MAX_PEERS + 2)));
        CU(cudaMalloc(&h->slab_prefix, sizeof(unsigned long long) * ((size_t)h->sg.ntiles + 2)));
        CU(cudaEventCreateWithFlags(&h->bounds_ev, cudaEventDisableTiming));
      }
      k_slab_bounds<<<1, 1024, 0, h->compute>>>(h->fill_snap, nwords, h->np, h->sg, (unsigned)h->nrec, g.nbox3, (unsigned)(g.im * g.km),
                                              (unsigned)g.im, h->slab_prefix, h->d_bounds);
      CU(cudaMemcpyAsync(h->h_bounds, h->d_bounds, sizeof(unsigned) * (h->np + 2), cudaMemcpyDeviceToHost, h->compute));
      CU(cudaEventRecord(h->bounds_ev, h->compute));
      ps.bounds = h->d_bounds;
      ps.self = h->prank;
      h->launches += 1;
    }
    const OutPlanes o2 = {dst[0], dst[1], dst[2], dst[3], dst[4], xpl};
    const OutPlanes o3 = {dst[5], dst[6], dst[7], dst[8], dst[9], raw3 ? 1 : xpl};
    const int r3 = raw3 ? 1 : 0;
    const unsigned T = 1u << h->sg.tile_shift;
    unsigned t0 = (unsigned)(sl.lo >> h->sg.tile_shift), t1 = (unsigned)((sl.hi + T - 1) >> h->sg.tile_shift);
    if (balance) { t0 = 0u; t1 = h->sg.ntiles; }        // the kernel takes its tile range from the device-side boundaries
    // Entry pull or planes exchange?  Per rank the entry pull moves 32 B x (entries the rank parked) x (N-1)/N over NVLink, the
    // planes exchange 48 B x boxes x (N-1)/N (nr = 2) plus one more pass over the rank's lists and the planes (K2x).  Measured
    // at cfg2: the pull wins at 2.25 entries per box and rank (N = 2 month: 1.9 against 2.7 ms), the planes win at 4.5 (eight
    // months on 8 ranks).  Every rank sees the same total, hence decides alike.
    bool planes = false;
    if (balance && !h->det_tiled && h->peer_red[h->prank] && h->red) {
      const double thr = getenv("GRITAS_B200_PLANES_ABOVE") ? atof(getenv("GRITAS_B200_PLANES_ABOVE")) : 3.0;
      CU(cudaEventSynchronize(h->bounds_ev));
      planes = (double)h->h_bounds[h->np + 1] > thr * (double)h->nrec * (double)h->np;
      for (int s2 = 0; s2 < h->np; ++s2) planes = planes && h->peer_red[s2] != nullptr;
    }
    if (planes) {
      // K2x: this rank's lists -> its export planes (all tiles); barrier; K3p: the slab's planes summed over the ranks
      if (!h->comm && !h->export_valid)
        return fail(h, GRITAS_B200_ERR_ARG, "norm_sharded without a communicator (several handles in one process): the ranks hold "
                                            "enough entries for the planes exchange -- call gritas_b200_export_planes on every handle first");
      if (!h->export_valid) {
        switch (nr) {
          case 1: launch_tile_export<1>(h, h->compute, 0u, h->sg.ntiles); break;
          default: launch_tile_export<2>(h, h->compute, 0u, h->sg.ntiles); break;
        }
        h->launches += 1;
      }
      if ((rc = stream_barrier(h))) return rc;
      const int ctas = h->tf_ctas;
#define PIMP_LAUNCH(NR_)                                                                                                              \
  k_tile_final<NR_, TILE_IMPORT_PEERS><<<ctas, TileFinal<NR_>::THREADS, h->tf_smem, h->compute>>>(g, h->sg, (unsigned)h->nrec, o2, o3, nrmlz, \
                                                                                                 ntr, 0, r3, 1, t0, t1, nullptr, ps)
      if (nr == 1) PIMP_LAUNCH(1);
      else PIMP_LAUNCH(2);
#undef PIMP_LAUNCH
      h->launches += 1;
    } else if (t1 > t0 && h->det_tiled) {
      launch_det_nr<DET_PEERS>(h, h->compute, t0, t1, o2, o3, nrmlz, ntr, 0, r3, ps);
    } else if (t1 > t0) {
      const int ctas = (int)std::min((unsigned)h->tf_ctas, t1 - t0);
#define PEER_LAUNCH(NR_)                                                                                                      \
  k_tile_final<NR_, TILE_PEERS><<<ctas, TileFinal<NR_>::THREADS, h->tf_smem, h->compute>>>(g, h->sg, (unsigned)h->nrec, o2, o3, nrmlz, \
                                                                                          ntr, 0, r3, 0, t0, t1, nullptr, ps)
      if (nr == 1) PEER_LAUNCH(1);
      else PEER_LAUNCH(2);
#undef PEER_LAUNCH
    }
    h->launches += 2;
    CU(cudaGetLastError());
    if (balance) {
      CU(cudaEventSynchronize(h->bounds_ev));             // (long done: the tile kernel is still running)
      sl = slab_from(g, h->h_bounds[h->prank], h->h_bounds[h->prank + 1]);
      h->bounds_valid = true;
    }
  } else {
    h->bounds_valid = false;
    // no lists on this path (deterministic mode, GRITAS_B200_FLAG_NO_STAGING): the records of this rank's rows are summed
    // over the ranks in rank order, then finalized by K3.  One-shot: this rank's records then hold the total of its rows.
    const size_t d0 = sl.lo * (size_t)g.rs, d1 = sl.hi * (size_t)g.rs;
    if (d1 > d0) {
      k_sum_peer_records<<<grid_for((d1 - d0), 256, 8), 256, 0, h->compute>>>(g.acc, ps, h->prank, d0, d1);
      h->launches += 1;
    }
    if ((rc = touch_records(h))) return rc;
    const size_t b0 = (std::max(sl.lo, n3) - n3), b1 = (std::max(sl.hi, n3) - n3);
    if ((rc = launch_finalize(h, h->compute, dst, 0, raw3, xpl, nrmlz, ntr, sl.jl3_0, sl.jl3_1, b1 > b0, b0, b1))) return rc;
  }
  if ((rc = stream_barrier(h))) return rc;            // every rank has read what it needs from this rank
  if ((rc = mark_main(h))) return rc;
  for (int a = 0; a < 10; ++a) {
    if (dev_out) dev_out[a] = dst[a];
    else if (bounce[a] && (rc = copy_slab(h, sl, a, (a == 3 || a == 8) ? ((a == 8 && raw3) ? 1 : xpl) : nr, user[a], dst[a]))) return rc;
  }
  return drain(h, nullptr);
}

}  // namespace

int gritas_b200_peer_export(gritas_b200_handle h, void *blob) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (!blob) return fail(h, GRITAS_B200_ERR_ARG, "null blob");
  PeerBlob b;
  memset(&b, 0, sizeof(b));
  b.magic = PEER_MAGIC;
  b.pid = (int32_t)getpid();
  b.device = h->device;
  b.staging = h->staging ? 1 : 0;
  b.nr = h->g.nr;
  b.tile_shift = h->staging ? h->sg.tile_shift : 0;
  b.mode = h->mode;
  b.ntiles = h->staging ? h->sg.ntiles : 0;
  b.cap = h->staging ? h->sg.cap : 0;
  b.nrec = h->nrec;
  b.raw_lists = (uint64_t)(uintptr_t)h->sg.lists;
  b.raw_fill = (uint64_t)(uintptr_t)h->sg.fill;
  b.raw_acc = (uint64_t)(uintptr_t)h->g.acc;
  if (h->staging && h->fused_norm && !h->det_tiled) {
    // export buffer of the planes exchange (what K2x writes): its address must be known to the peers now
    const int nv1 = 1 + (h->g.nr == 1 ? 2 : (h->g.nr == 2 ? 5 : 6));
    const size_t T = (size_t)1 << h->sg.tile_shift;
    if ((rc = grow_dev(h, (void **)&h->red, &h->red_cap, sizeof(double) * (size_t)nv1 * T * h->sg.ntiles))) return rc;
  }
  b.raw_red = (uint64_t)(uintptr_t)h->red;
  gethostname(b.host, sizeof(b.host) - 1);
  b.ipc_ok = 1;
  if (cudaIpcGetMemHandle(&b.ipc_acc, h->g.acc) != cudaSuccess) b.ipc_ok = 0;
  if (h->staging && (cudaIpcGetMemHandle(&b.ipc_lists, h->sg.lists) != cudaSuccess ||
                     cudaIpcGetMemHandle(&b.ipc_fill, h->sg.fill) != cudaSuccess))
    b.ipc_ok = 0;
  if (h->red && cudaIpcGetMemHandle(&b.ipc_red, h->red) != cudaSuccess) b.ipc_ok = 0;
  cudaGetLastError();
  memset(blob, 0, GRITAS_B200_PEER_BLOB_BYTES);
  memcpy(blob, &b, sizeof(b));
  return 0;
}

int gritas_b200_peer_attach(gritas_b200_handle h, int nranks, int rank, const void *blobs) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (nranks < 1 || nranks > MAX_PEERS || rank < 0 || rank >= nranks || !blobs)
    return fail(h, GRITAS_B200_ERR_ARG, "peer_attach: 1 <= nranks <= %d, 0 <= rank < nranks", MAX_PEERS);
  if (h->np > 1) return fail(h, GRITAS_B200_ERR_ARG, "peer_attach: already attached");
  if (h->comm && h->nranks != nranks) return fail(h, GRITAS_B200_ERR_ARG, "peer_attach: nranks differs from comm_init");
  PeerBlob b[MAX_PEERS];
  for (int s = 0; s < nranks; ++s) {
    memcpy(&b[s], (const char *)blobs + (size_t)s * GRITAS_B200_PEER_BLOB_BYTES, sizeof(PeerBlob));
    if (b[s].magic != PEER_MAGIC) return fail(h, GRITAS_B200_ERR_ARG, "peer_attach: blob %d is not a peer_export blob", s);
    // what decides the shape of the exchange must be the same everywhere (every rank sees every blob: all fail alike)
    if (b[s].staging != b[0].staging || b[s].nr != b[0].nr || b[s].tile_shift != b[0].tile_shift || b[s].ntiles != b[0].ntiles ||
        b[s].cap != b[0].cap || b[s].nrec != b[0].nrec || (b[s].mode & 0xff) != (b[0].mode & 0xff))
      return fail(h, GRITAS_B200_ERR_ARG,
                  "peer_attach: rank %d differs from rank 0 (tiled %d/%d, nr %d/%d, tile_shift %d/%d, ntiles %u/%u, cap %u/%u, "
                  "nrec %llu/%llu): same grid, mode and GRITAS_B200_* settings on every rank (GRITAS_B200_STAGE_SLOTS pins cap)",
                  s, b[s].staging, b[0].staging, b[s].nr, b[0].nr, b[s].tile_shift, b[0].tile_shift, b[s].ntiles, b[0].ntiles,
                  b[s].cap, b[0].cap, (unsigned long long)b[s].nrec, (unsigned long long)b[0].nrec);
  }
  if ((uint64_t)(uintptr_t)h->g.acc != b[rank].raw_acc || b[rank].pid != (int32_t)getpid())
    return fail(h, GRITAS_B200_ERR_ARG, "peer_attach: blob %d is not this handle's", rank);
  for (int s = 0; s < nranks; ++s) {
    if (s == rank || b[s].pid == (int32_t)getpid()) {
      // same process (several handles driven by one host thread, or this rank itself): plain pointers
      if (b[s].device != h->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(b[s].device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return fail(h, GRITAS_B200_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", b[s].device, cudaGetErrorString(e));
        cudaGetLastError();
      }
      h->peer_lists[s] = (void *)(uintptr_t)b[s].raw_lists;
      h->peer_fill[s] = (void *)(uintptr_t)b[s].raw_fill;
      h->peer_acc[s] = (void *)(uintptr_t)b[s].raw_acc;
      h->peer_red[s] = (void *)(uintptr_t)b[s].raw_red;
      continue;
    }
    if (strncmp(b[s].host, b[rank].host, sizeof(b[s].host)) != 0)
      return fail(h, GRITAS_B200_ERR_ARG, "peer_attach: rank %d runs on host %s, this rank on %s (peer memory is one node)", s, b[s].host,
                  b[rank].host);
    if (!b[s].ipc_ok) return fail(h, GRITAS_B200_ERR_CUDA, "peer_attach: rank %d could not export CUDA IPC handles", s);
    CU(cudaIpcOpenMemHandle(&h->peer_acc[s], b[s].ipc_acc, cudaIpcMemLazyEnablePeerAccess));
    h->peer_ipc[s] = true;
    if (b[s].staging) {
      CU(cudaIpcOpenMemHandle(&h->peer_lists[s], b[s].ipc_lists, cudaIpcMemLazyEnablePeerAccess));
      CU(cudaIpcOpenMemHandle(&h->peer_fill[s], b[s].ipc_fill, cudaIpcMemLazyEnablePeerAccess));
    }
    if (b[s].raw_red) CU(cudaIpcOpenMemHandle(&h->peer_red[s], b[s].ipc_red, cudaIpcMemLazyEnablePeerAccess));
  }
  h->np = nranks;
  h->prank = rank;
  return 0;
}

int gritas_b200_slab(gritas_b200_handle h, int rank, int nranks, int *jl3_0, int *jl3_1, int *jl2_0, int *jl2_1) {
  if (!h || nranks < 1 || rank < 0 || rank >= nranks) return GRITAS_B200_ERR_ARG;
  // after a sharded Norm on the tiled path: the boundaries that Norm used (balanced by parked entries, the same on every rank)
  const Slab s = (h->bounds_valid && nranks == h->np) ? slab_from(h->g, h->h_bounds[rank], h->h_bounds[rank + 1])
                                                      : slab_of(h->g, h->nrec, rank, nranks);
  if (jl3_0) *jl3_0 = s.jl3_0;
  if (jl3_1) *jl3_1 = s.jl3_1;
  if (jl2_0) *jl2_0 = s.jl2_0;
  if (jl2_1) *jl2_1 = s.jl2_1;
  return 0;
}

int gritas_b200_norm_sharded(gritas_b200_handle h, int nrmlz, int ntresh, double *bias_2d, double *rtms_2d, double *stdv_2d,
                             double *xcov_2d, double *nobs_2d, double *bias_3d, double *rtms_3d, double *stdv_3d,
                             double *xcov_3d, double *nobs_3d) {
  int rc = check_handle(h);
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->copy));
  double *const u[10] = {bias_2d, rtms_2d, stdv_2d, xcov_2d, nobs_2d, bias_3d, rtms_3d, stdv_3d, xcov_3d, nobs_3d};
  if (h->np < 2) {
    if ((rc = settle(h, fused_norm_ok(h), nullptr))) return rc;
    return finalize_to(h, 0, nrmlz ? 1 : 0, ntresh, u, nullptr);
  }
  return finalize_sharded(h, nrmlz ? 1 : 0, ntresh, u, nullptr);
}

int gritas_b200_norm_sharded_device(gritas_b200_handle h, int nrmlz, int ntresh, const double **outs10) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (!outs10) return fail(h, GRITAS_B200_ERR_ARG, "null outs");
  CU(cudaStreamSynchronize(h->copy));
  double *const u[10] = {};
  if (h->np < 2) {
    if ((rc = settle(h, fused_norm_ok(h), nullptr))) return rc;
    return finalize_to(h, 0, nrmlz ? 1 : 0, ntresh, u, outs10);
  }
  return finalize_sharded(h, nrmlz ? 1 : 0, ntresh, u, outs10);
}

// Phase 1 of the planes exchange on its own, for a host that drives several handles without a communicator: every handle
// exports its tile sums, then every handle runs gritas_b200_norm_sharded.  (With a communicator norm_sharded does both.)
int gritas_b200_export_planes(gritas_b200_handle h) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (!h->staging || !h->red || h->det_tiled) return 0;      // nothing to export on this path
  if ((rc = settle(h, true, nullptr))) return rc;
  switch (h->g.nr) {
    case 1: launch_tile_export<1>(h, h->compute, 0u, h->sg.ntiles); break;
    case 2: launch_tile_export<2>(h, h->compute, 0u, h->sg.ntiles); break;
    default: return 0;
  }
  h->launches += 1;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->compute));
  h->export_valid = true;
  return 0;
}

int gritas_b200_host_register(void *p, size_t bytes) {
  if (!p || !bytes) return GRITAS_B200_ERR_ARG;
  cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return 0; }
  if (e != cudaSuccess) { cudaGetLastError(); return GRITAS_B200_ERR_CUDA; }
  return 0;
}

int gritas_b200_host_unregister(void *p) {
  if (!p) return GRITAS_B200_ERR_ARG;
  if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return GRITAS_B200_ERR_CUDA; }
  return 0;
}

// ---- host-side table builders -------------------------------------------------------------------
// GrITAS_Pint_ (m_gritas_grids.f:217-296).  Level k collects p1(k) >= lev > p2(k); the interfaces are
// the log-pressure mid-points, the first interval opens at 2000 hPa and the last closes at 0.
int gritas_b200_pint(int nlevs, const double *plevs, double *p1, double *p2) {
  if (nlevs <= 0 || !plevs || !p1 || !p2) return 1;
  if (nlevs == 1) {  // single-level case: a 0.1 hPa slab below the level (:249-256)
    p1[0] = plevs[0];
    p2[0] = plevs[0] - 0.1;
    return 0;
  }
  for (int k = 1; k < nlevs; ++k)
    if (!(plevs[k - 1] > plevs[k])) return 2;
  if (plevs[nlevs - 1] == 0.0) return 3;
  auto mid = [&](int a, int b) { return exp(0.5 * (log(plevs[a]) + log(plevs[b]))); };
  for (int k = 0; k < nlevs; ++k) {
    p1[k] = (k == 0) ? 2000.0 : mid(k - 1, k);
    p2[k] = (k == nlevs - 1) ? 0.0 : mid(k + 1, k);
  }
  return 0;
}

// GrITAS_KxKt (gritas_kxkt.f:10-119): every rc row becomes one output grid; surface rows (kt < 0 or
// kt in ktSurfAll, gritas_kxkt.f:133-164) number the 2-D grids, the others the 3-D grids; a (kx,kt)
// named by several rows belongs to the last one.
int gritas_b200_kxkt(int ktmax, int kxmax, int l2d_max, int l3d_max, int listsz, const int32_t *kt_list,
                     const int32_t *kx_list, const int32_t *kt_surf, int n_surf, int32_t *kxkt_table,
                     int32_t *list_2d, int32_t *list_3d, int *l2d_out, int *l3d_out) {
  int n2 = 0, n3 = 0;
  for (size_t e = 0; e < (size_t)kxmax * (size_t)ktmax; ++e) kxkt_table[e] = 0;
  for (int row = 0; row < listsz; ++row) {
    const int signed_kt = kt_list[row];
    const int kt = signed_kt < 0 ? -signed_kt : signed_kt;
    if (kt > ktmax) return 1;
    if (kt < 1) return 2;
    bool surface = signed_kt < 0;
    for (int s = 0; s < n_surf && !surface; ++s) surface = (kt_surf[s] == signed_kt);
    int value;
    if (surface) {
      if (++n2 > l2d_max) return 3;
      if (list_2d) list_2d[n2 - 1] = row + 1;
      value = -n2;
    } else {
      if (++n3 > l3d_max) return 4;
      if (list_3d) list_3d[n3 - 1] = row + 1;
      value = n3;
    }
    const int32_t *kxs = kx_list + (size_t)kxmax * (size_t)row;
    for (int e = 0; e < kxmax && kxs[e] >= 0; ++e) {
      if (kxs[e] > kxmax) return 5;
      if (kxs[e] == 0) continue;  // kx = 0 would index out of bounds in the reference
      kxkt_table[(size_t)(kxs[e] - 1) + (size_t)kxmax * (size_t)(kt - 1)] = value;
    }
  }
  if (l2d_out) *l2d_out = n2;
  if (l3d_out) *l3d_out = n3;
  return 0;
}

// ---- utilities ------------------------------------------------------------------------------------
void *gritas_b200_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void gritas_b200_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

void *gritas_b200_stream(gritas_b200_handle h) { return h ? (void *)h->compute : nullptr; }

int gritas_b200_set_stream(gritas_b200_handle h, void *stream) {
  int rc = check_handle(h);
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->compute));
  if (h->own_compute) CU(cudaStreamDestroy(h->compute));
  h->compute = (cudaStream_t)stream;
  h->own_compute = false;
  return 0;
}

int gritas_b200_tiled(gritas_b200_handle h) { return (h && h->staging) ? h->tiled_choice : 0; }
int gritas_b200_fused_norm(gritas_b200_handle h) { return (h && fused_norm_ok(h)) ? 1 : 0; }
uint64_t gritas_b200_launch_count(gritas_b200_handle h) { return h ? h->launches : 0; }
uint64_t gritas_b200_h2d_bytes(gritas_b200_handle h) { return h ? h->h2d : 0; }
uint64_t gritas_b200_d2h_bytes(gritas_b200_handle h) { return h ? h->d2h : 0; }

}  // extern "C"
