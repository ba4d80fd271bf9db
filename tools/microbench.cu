// microbench.cu -- measurements that decide the accumulate design (not part of the product).
//   A: fp64 RED into random 64 B records over a region of R bytes (6 REDs per record visit)
//   B: shared-memory int32 atomicAdd with return (counting-sort rank), random bins
//   C: shared-memory fp64 atomicAdd, random bins
//   D: plain streaming read of 52 B/obs (upper bound of the classify pass)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}

template <int NRED>
__global__ void k_red(double *acc, uint32_t nrec, uint32_t nvisit, uint32_t seed) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nvisit; i += gridDim.x * blockDim.x) {
    uint32_t r = (uint32_t)(((uint64_t)hash32(i * 2654435761u + seed) * nrec) >> 32);
    double *p = acc + (size_t)r * 8;
    double v = (double)(i & 7);
#pragma unroll
    for (int k = 0; k < NRED; ++k) atomicAdd(p + k, v);
  }
}

__global__ void k_smem_rank(uint32_t *out, uint32_t nvisit, uint32_t nbins, uint32_t seed) {
  extern __shared__ uint32_t bins[];
  for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x) bins[b] = 0;
  __syncthreads();
  uint32_t acc = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nvisit; i += gridDim.x * blockDim.x) {
    uint32_t r = (uint32_t)(((uint64_t)hash32(i * 2654435761u + seed) * nbins) >> 32);
    acc += atomicAdd(&bins[r], 1u);
  }
  __syncthreads();
  if (acc == 0xdeadbeef) out[0] = acc + bins[threadIdx.x % nbins];
}

__global__ void k_smem_f64(double *out, uint32_t nvisit, uint32_t nbins, uint32_t seed) {
  extern __shared__ double dbins[];
  for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x) dbins[b] = 0;
  __syncthreads();
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nvisit; i += gridDim.x * blockDim.x) {
    uint32_t r = (uint32_t)(((uint64_t)hash32(i * 2654435761u + seed) * nbins) >> 32);
    atomicAdd(&dbins[r], 1.5);
  }
  __syncthreads();
  if (nvisit == 0xdeadbeef) out[0] = dbins[threadIdx.x % nbins];
}

__global__ void k_stream(const double2 *a, const double2 *b, const double2 *c, const double2 *d, const double2 *e,
                         const int4 *x, const int4 *y, const int4 *z, size_t n4, double *out) {
  double s = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    double2 p = a[2 * i], q = a[2 * i + 1]; s += p.x + p.y + q.x + q.y;
    p = b[2 * i]; q = b[2 * i + 1]; s += p.x + p.y + q.x + q.y;
    p = c[2 * i]; q = c[2 * i + 1]; s += p.x + p.y + q.x + q.y;
    p = d[2 * i]; q = d[2 * i + 1]; s += p.x + p.y + q.x + q.y;
    p = e[2 * i]; q = e[2 * i + 1]; s += p.x + p.y + q.x + q.y;
    int4 u = x[i]; s += u.x + u.y + u.z + u.w;
    u = y[i]; s += u.x + u.y + u.z + u.w;
    u = z[i]; s += u.x + u.y + u.z + u.w;
  }
  if (s == 1.2345) out[0] = s;
}

// E: ticket atomics (atomicAdd with return) on `nc` counters, optionally followed by dependent scattered stores
template <int ILP, bool STORE>
__global__ void k_ticket(unsigned *cnt, unsigned nc, unsigned cap, unsigned *okey, double *od0, double *od1, uint32_t nvisit,
                         uint32_t seed) {
  for (uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * ILP; i0 < nvisit; i0 += gridDim.x * blockDim.x * ILP) {
    unsigned t[ILP], pos[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) t[k] = (unsigned)(((uint64_t)hash32((i0 + k) * 2654435761u + seed) * nc) >> 32);
#pragma unroll
    for (int k = 0; k < ILP; ++k) pos[k] = atomicAdd(cnt + t[k], 1u);
    if (STORE) {
#pragma unroll
      for (int k = 0; k < ILP; ++k) {
        const size_t at = (size_t)t[k] * cap + (pos[k] % cap);
        okey[at] = i0 + k; od0[at] = 1.0 * k; od1[at] = 2.0 * k;
      }
    } else {
      unsigned s = 0;
#pragma unroll
      for (int k = 0; k < ILP; ++k) s += pos[k];
      if (s == 0xdeadbeef) okey[0] = s;
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
  const size_t maxbytes = (size_t)2 << 30;
  double *acc; CK(cudaMalloc(&acc, maxbytes)); CK(cudaMemset(acc, 0, maxbytes));
  const uint32_t nvisit = 16u << 20;
  printf("A: fp64 RED, 64 B records, %u visits\n", nvisit);
  for (size_t mb : {32, 1728}) {
    uint32_t nrec = (uint32_t)((mb << 20) / 64);
    for (int nred : {1, 3, 6}) {
      float ms = 0;
      uint32_t seed = 1;
      auto f = [&]() {
        seed += 77;
        if (nred == 1) k_red<1><<<148 * 8, 256>>>(acc, nrec, nvisit, seed);
        if (nred == 3) k_red<3><<<148 * 8, 256>>>(acc, nrec, nvisit, seed);
        if (nred == 6) k_red<6><<<148 * 8, 256>>>(acc, nrec, nvisit, seed);
      };
      ms = timeit(f, 5);
      printf("  region %5zu MB  nred %d : %8.3f ms  %7.2f G visits/s  %7.2f G red/s\n", mb, nred, ms, nvisit / ms / 1e6,
             (double)nvisit * nred / ms / 1e6);
    }
  }
  {
    const uint32_t nv = 32u << 20;
    const size_t slots = (size_t)256 << 20;
    unsigned *cnt, *okey; double *od0, *od1;
    CK(cudaMalloc(&cnt, 4u << 20)); CK(cudaMalloc(&okey, slots * 4)); CK(cudaMalloc(&od0, slots * 8)); CK(cudaMalloc(&od1, slots * 8));
    printf("E: ticket atomics with return, %u visits (1.21M-visit launches emulate one file)\n", nv);
    for (unsigned nc : {256u, 1024u, 6618u, 13236u, 52944u, 423552u}) {
      unsigned cap = (unsigned)(slots / nc);
      for (int store = 0; store < 2; ++store) {
        for (int ilp : {1, 4}) {
          CK(cudaMemset(cnt, 0, 4u << 20));
          uint32_t seed = 3;
          auto f = [&]() {
            seed += 11;
            if (store) { if (ilp == 1) k_ticket<1, true><<<148 * 8, 256>>>(cnt, nc, cap, okey, od0, od1, nv, seed);
                         else k_ticket<4, true><<<148 * 8, 256>>>(cnt, nc, cap, okey, od0, od1, nv, seed); }
            else { if (ilp == 1) k_ticket<1, false><<<148 * 8, 256>>>(cnt, nc, cap, okey, od0, od1, nv, seed);
                   else k_ticket<4, false><<<148 * 8, 256>>>(cnt, nc, cap, okey, od0, od1, nv, seed); }
          };
          float ms = timeit(f, 3);
          printf("  counters %7u store %d ilp %d : %8.3f ms  %7.2f G tickets/s\n", nc, store, ilp, ms, nv / ms / 1e6);
        }
      }
    }
    cudaFree(cnt); cudaFree(okey); cudaFree(od0); cudaFree(od1);
  }
  uint32_t *o32; CK(cudaMalloc(&o32, 1024));
  double *o64; CK(cudaMalloc(&o64, 1024));
  const uint32_t nv2 = 64u << 20;
  printf("B: smem int32 atomicAdd with return, %u visits\n", nv2);
  for (uint32_t nbins : {256u, 2048u, 8192u, 32768u}) {
    for (int thr : {256, 1024}) {
      auto f = [&]() { k_smem_rank<<<148 * (2048 / thr), thr, nbins * 4>>>(o32, nv2, nbins, 5); };
      CK(cudaFuncSetAttribute(k_smem_rank, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      float ms = timeit(f, 5);
      printf("  bins %6u thr %4d : %8.3f ms  %7.2f G atom/s\n", nbins, thr, ms, nv2 / ms / 1e6);
    }
  }
  printf("C: smem fp64 atomicAdd, %u visits\n", nv2);
  for (uint32_t nbins : {256u, 2048u, 8192u, 24576u}) {
    auto f = [&]() { k_smem_f64<<<148 * 2, 1024, nbins * 8>>>(o64, nv2, nbins, 5); };
    CK(cudaFuncSetAttribute(k_smem_f64, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    float ms = timeit(f, 5);
    printf("  bins %6u : %8.3f ms  %7.2f G atom/s\n", nbins, ms, nv2 / ms / 1e6);
  }
  {
    const size_t n = 64u << 20;  // obs
    double *cols[5]; int *ic[3];
    for (auto &c : cols) { CK(cudaMalloc(&c, n * 8)); CK(cudaMemset(c, 0, n * 8)); }
    for (auto &c : ic) { CK(cudaMalloc(&c, n * 4)); CK(cudaMemset(c, 0, n * 4)); }
    for (int bps : {4, 8, 16}) {
      auto f = [&]() {
        k_stream<<<148 * bps, 256>>>((double2 *)cols[0], (double2 *)cols[1], (double2 *)cols[2], (double2 *)cols[3],
                                     (double2 *)cols[4], (int4 *)ic[0], (int4 *)ic[1], (int4 *)ic[2], n / 4, o64);
      };
      float ms = timeit(f, 5);
      printf("D: stream 52 B/obs, %zu obs, %d blk/SM: %8.3f ms  %7.1f GB/s  %6.2f G obs/s\n", n, bps, ms, n * 52.0 / ms / 1e6,
             n / ms / 1e6);
    }
  }
  return 0;
}
