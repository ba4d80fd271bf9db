// microbench2.cu -- can a ONE-level partition work?  Every accepted observation takes a ticket in its tile's list
// (global atomicAdd with return) and writes ONE full 32-byte sector {key, -, d1, d2} at the ticket position.
//   F: ticket + one 256-bit store (st.global.v4.f64) / two 128-bit stores / three plane stores, 13236 lists
//   G: the same behind a 52 B/obs streaming read (what K1 would be), 81 % accepted
//   H: scattered 32 B stores without tickets (store path alone)
// Not part of the product.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}

struct alignas(32) Entry { unsigned key, pad0; double pad1; double d1, d2; };

__device__ __forceinline__ void st256(Entry *p, unsigned key, double d1, double d2) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(__longlong_as_double((long long)key)), "d"(0.0), "d"(d1), "d"(d2) : "memory");
}
__device__ __forceinline__ void st128x2(Entry *p, unsigned key, double d1, double d2) {
  double2 *q = reinterpret_cast<double2 *>(p);
  q[0] = make_double2(__longlong_as_double((long long)key), 0.0);
  q[1] = make_double2(d1, d2);
}

// MODE 0: ticket only; 1: ticket + st256; 2: ticket + 2 x st128; 3: st256 at a random slot without ticket
template <int MODE, int ILP>
__global__ void k_f(unsigned *cnt, unsigned nc, unsigned cap, Entry *lists, uint32_t nvisit, uint32_t seed) {
  for (uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * ILP; i0 < nvisit; i0 += gridDim.x * blockDim.x * ILP) {
    unsigned t[ILP], pos[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) t[k] = (unsigned)(((uint64_t)hash32((i0 + k) * 2654435761u + seed) * nc) >> 32);
#pragma unroll
    for (int k = 0; k < ILP; ++k) pos[k] = (MODE == 3) ? (hash32(i0 + k + seed * 7u) % cap) : atomicAdd(cnt + t[k], 1u);
    if (MODE == 0) {
      unsigned s = 0;
#pragma unroll
      for (int k = 0; k < ILP; ++k) s += pos[k];
      if (s == 0xdeadbeef) lists[0].key = s;
    } else {
#pragma unroll
      for (int k = 0; k < ILP; ++k) {
        Entry *p = lists + (size_t)t[k] * cap + (pos[k] % cap);
        if (MODE == 2) st128x2(p, i0 + k, 1.0 * k, 2.0 * k);
        else st256(p, i0 + k, 1.0 * k, 2.0 * k);
      }
    }
  }
}

// G: 4 obs per thread, 52 B/obs streamed; "classify" = a hash of the loaded data; 81 % accepted
template <int MODE>
__global__ void __launch_bounds__(256)
k_g(const double2 *a, const double2 *b, const double2 *c, const double2 *d, const double2 *e, const int4 *x, const int4 *y,
    const int4 *z, size_t n4, unsigned *cnt, unsigned nc, unsigned cap, Entry *lists, uint32_t seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    double2 p0 = a[2 * i], p1 = a[2 * i + 1], q0 = b[2 * i], q1 = b[2 * i + 1], r0 = c[2 * i], r1 = c[2 * i + 1];
    int4 u = x[i], v = y[i], w = z[i];
    double2 d0 = d[2 * i], d1 = d[2 * i + 1], e0 = e[2 * i], e1 = e[2 * i + 1];
    const double geo[4] = {p0.x + q0.x + r0.x, p0.y + q0.y + r0.y, p1.x + q1.x + r1.x, p1.y + q1.y + r1.y};
    const int ii[4] = {u.x + v.x + w.x, u.y + v.y + w.y, u.z + v.z + w.z, u.w + v.w + w.w};
    const double dd1[4] = {d0.x, d0.y, d1.x, d1.y}, dd2[4] = {e0.x, e0.y, e1.x, e1.y};
    unsigned t[4], pos[4];
    bool acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t h = hash32((uint32_t)(4 * i + k) * 2654435761u + seed + (uint32_t)ii[k] + (uint32_t)__double2int_rn(geo[k]));
      acc[k] = (h & 0xffu) < 207u;   // 81 %
      t[k] = (unsigned)(((uint64_t)hash32(h) * nc) >> 32);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) pos[k] = acc[k] ? atomicAdd(cnt + t[k], 1u) : 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!acc[k]) continue;
      Entry *p = lists + (size_t)t[k] * cap + (pos[k] % cap);
      if (MODE == 2) st128x2(p, t[k], dd1[k], dd2[k]);
      else st256(p, t[k], dd1[k], dd2[k]);
    }
  }
}

template <typename F>
float timeit(F f, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  const unsigned nc = 13236;
  const size_t slots = (size_t)256 << 20;          // 8 GB of 32-byte entries
  const unsigned cap = (unsigned)(slots / nc);
  unsigned *cnt; Entry *lists;
  CK(cudaMalloc(&cnt, 4u << 20)); CK(cudaMalloc(&lists, slots * sizeof(Entry)));
  CK(cudaMemset(lists, 0, slots * sizeof(Entry)));
  const uint32_t nv = 120u << 20;                  // ~ the accepted observations of one month
  printf("F: %u tickets on %u lists (cap %u), one 32 B entry per ticket\n", nv, nc, cap);
  const char *names[4] = {"ticket only", "ticket + st.v4.f64 (256-bit)", "ticket + 2 x 128-bit", "st.v4.f64 at a random slot, no ticket"};
  for (int mode = 0; mode < 4; ++mode)
    for (int ilp : {1, 4})
      for (int bps : {8, 16}) {
        CK(cudaMemset(cnt, 0, 4u << 20));
        uint32_t seed = 3;
        auto f = [&]() {
          seed += 11;
#define LAUNCH(M, I) k_f<M, I><<<148 * bps, 256>>>(cnt, nc, cap, lists, nv, seed)
          if (mode == 0) { if (ilp == 1) LAUNCH(0, 1); else LAUNCH(0, 4); }
          if (mode == 1) { if (ilp == 1) LAUNCH(1, 1); else LAUNCH(1, 4); }
          if (mode == 2) { if (ilp == 1) LAUNCH(2, 1); else LAUNCH(2, 4); }
          if (mode == 3) { if (ilp == 1) LAUNCH(3, 1); else LAUNCH(3, 4); }
        };
        float ms = timeit(f, 3);
        printf("  %-40s ilp %d blk/SM %2d : %8.3f ms  %7.2f G entries/s  (%.0f GB/s of entries)\n", names[mode], ilp, bps, ms,
               nv / ms / 1e6, nv * 32.0 / ms / 1e6);
      }
  {
    const size_t n = 148u << 20;  // obs of one month
    double *cols[5]; int *ic[3];
    for (auto &c : cols) { CK(cudaMalloc(&c, n * 8)); CK(cudaMemset(c, 0, n * 8)); }
    for (auto &c : ic) { CK(cudaMalloc(&c, n * 4)); CK(cudaMemset(c, 0, n * 4)); }
    printf("G: stream 52 B/obs + ticket + 32 B entry for 81 %% of %zu obs\n", n);
    for (int mode : {1, 2})
      for (int bps : {4, 8}) {
        CK(cudaMemset(cnt, 0, 4u << 20));
        uint32_t seed = 5;
        auto f = [&]() {
          seed += 13;
          if (mode == 1)
            k_g<1><<<148 * bps, 256>>>((double2 *)cols[0], (double2 *)cols[1], (double2 *)cols[2], (double2 *)cols[3], (double2 *)cols[4],
                                       (int4 *)ic[0], (int4 *)ic[1], (int4 *)ic[2], n / 4, cnt, nc, cap, lists, seed);
          else
            k_g<2><<<148 * bps, 256>>>((double2 *)cols[0], (double2 *)cols[1], (double2 *)cols[2], (double2 *)cols[3], (double2 *)cols[4],
                                       (int4 *)ic[0], (int4 *)ic[1], (int4 *)ic[2], n / 4, cnt, nc, cap, lists, seed);
        };
        float ms = timeit(f, 3);
        printf("  %s, %d blk/SM : %8.3f ms  %6.2f G obs/s\n", mode == 1 ? "256-bit store" : "2 x 128-bit", bps, ms, n / ms / 1e6);
      }
    // the same in 124 launches of 1/124 of the data (one launch per file), 3 streams
    cudaStream_t st[3];
    for (auto &s : st) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    const size_t per = (n / 124) & ~(size_t)3;
    for (int ns : {1, 3}) {
      CK(cudaMemset(cnt, 0, 4u << 20));
      auto f = [&]() {
        for (int fi = 0; fi < 124; ++fi) {
          const size_t o = fi * per;
          k_g<1><<<148 * 4, 256, 0, st[fi % ns]>>>((double2 *)(cols[0] + o), (double2 *)(cols[1] + o), (double2 *)(cols[2] + o),
                                                    (double2 *)(cols[3] + o), (double2 *)(cols[4] + o), (int4 *)(ic[0] + o),
                                                    (int4 *)(ic[1] + o), (int4 *)(ic[2] + o), per / 4, cnt, nc, cap, lists, 7 + fi);
        }
      };
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      f(); CK(cudaDeviceSynchronize());
      CK(cudaMemset(cnt, 0, 4u << 20));
      CK(cudaDeviceSynchronize());
      cudaEventRecord(a, st[0]);
      f();
      for (int s = 1; s < ns; ++s) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st[s]); cudaStreamWaitEvent(st[0], e, 0); }
      cudaEventRecord(b, st[0]);
      CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, a, b);
      printf("  124 launches of %zu obs on %d stream(s): %8.3f ms  %6.2f G obs/s\n", per, ns, ms, 124.0 * per / ms / 1e6);
    }
  }
  return 0;
}
