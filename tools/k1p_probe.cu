// k1p_probe.cu -- phase timing of the partition kernel on synthetic device-resident columns
// (development tool, not part of the product).  Build: nvcc -DGB_PHASE_TIMING ... -I gritas_b200/csrc
#define GB_PHASE_TIMING
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include "../gritas_b200/csrc/kernels.cuh"
using namespace gb200;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

int main(int argc, char **argv) {
  const int n = 1210000, im = 360, jm = 181, km = 26, l3d = 16, nr = 2;
  const double plevs[26] = {1000, 975, 950, 925, 900, 850, 800, 750, 700, 650, 600, 550, 500, 450, 400, 350, 300, 250, 200, 150, 100, 70, 50, 30, 20, 10};
  std::vector<double> p1(26), p2(26);
  for (int k = 0; k < 26; ++k) {
    p1[k] = k == 0 ? 2000.0 : exp(0.5 * (log(plevs[k - 1]) + log(plevs[k])));
    p2[k] = k == 25 ? 0.0 : exp(0.5 * (log(plevs[k + 1]) + log(plevs[k])));
  }
  const int kxmax = 1000, ktmax = 128;
  std::vector<int> table((size_t)kxmax * ktmax, 0);
  for (int l = 0; l < l3d; ++l) table[(size_t)(100 + l) + (size_t)kxmax * 43] = l + 1;   // kx = 101+l, kt = 44
  std::vector<double> lat(n), lon(n), lev(n), d0(n), d1(n);
  std::vector<int> kx(n), kt(n), qc(n);
  srand(7);
  auto u = []() { return (rand() + 0.5) / (RAND_MAX + 1.0); };
  for (int i = 0; i < n; ++i) {
    lat[i] = asin(2 * u() - 1) * 57.29577951308232; lon[i] = 360 * u() - 180; lev[i] = exp(log(10.0) + u() * (log(1050.0) - log(10.0)));
    d0[i] = u(); d1[i] = u(); kx[i] = 101 + rand() % 16; kt[i] = u() < 0.9 ? 44 : 7; qc[i] = u() < 0.9 ? 0 : 3;
  }
  GridDesc g = {};
  g.im = im; g.jm = jm; g.km = km; g.nlevs = 26; g.l2d = 0; g.l3d = l3d; g.nr = nr; g.rs = 8; g.kxmax = kxmax; g.ktmax = ktmax;
  g.contiguous = 1; g.p2_in_smem = 1; g.dlon = 360.0 / im; g.dlat = 180.0 / (jm - 1); g.rdlon = 1 / g.dlon; g.rdlat = 1 / g.dlat; g.p1_top = 2000.0;
  std::vector<unsigned short> lut(16384);
  g.lut_ncell = build_level_lut(26, p1.data(), p2.data(), lut.data(), 16384, &g.lut_shift, &g.lut_cell0);
  printf("lut cells %d shift %d\n", g.lut_ncell, g.lut_shift);
  g.nbox3 = (unsigned)((size_t)im * jm * km * l3d); g.nbox2 = 0;
  const size_t nrec = g.nbox3;
  auto up = [&](const void *src, size_t bytes) { void *d; CK(cudaMalloc(&d, bytes)); CK(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice)); return d; };
  g.lut = (unsigned short *)up(lut.data(), 16384 * 2);
  g.table = (int *)up(table.data(), table.size() * 4); g.p1 = (double *)up(p1.data(), 208); g.p2 = (double *)up(p2.data(), 208);
  CK(cudaMalloc(&g.acc, nrec * 8 * 8)); CK(cudaMemset(g.acc, 0, nrec * 64));
  ObsCols c = {};
  c.lat = (double *)up(lat.data(), n * 8); c.lon = (double *)up(lon.data(), n * 8); c.lev = (double *)up(lev.data(), n * 8);
  c.d[0] = (double *)up(d0.data(), n * 8); c.d[1] = (double *)up(d1.data(), n * 8);
  c.kx = (int *)up(kx.data(), n * 4); c.kt = (int *)up(kt.data(), n * 4); c.flag = (int *)up(qc.data(), n * 4);
  c.n = n; c.aligned = 1;
  FilterDesc f = {}; f.ods = 1; f.ioptn = 1; f.x_passive = 7; f.x_pre_bad = 1;
  Status *st; CK(cudaMalloc(&st, sizeof(Status)));
  Status hs; hs.first_bad = BAD_NONE; hs.bad_lev = 0; hs.near_pairs = 0; hs.pairs = 0; CK(cudaMemcpy(st, &hs, sizeof(hs), cudaMemcpyHostToDevice));
  StageDesc sg = {};
  sg.tile_shift = 11; const size_t T = 2048; const size_t ntiles = (nrec + T - 1) / T;
  size_t tpb = 1; while ((ntiles + tpb - 1) / tpb > 112) tpb <<= 1; int tsh = 0; while (((size_t)1 << tsh) < tpb) ++tsh;
  sg.bucket_shift = 11 + tsh; sg.ntiles = (unsigned)ntiles; sg.nbuckets = (unsigned)((ntiles + tpb - 1) / tpb);
  sg.cap = 24000; sg.c1 = 32; sg.c1_shift = 5; sg.nbatch_cap = 2048; sg.nbatch = (n + P1_BATCH - 1) / P1_BATCH;
  CK(cudaMalloc(&sg.fill, sg.nbuckets * tpb * 4 + 16)); CK(cudaMemset(sg.fill, 0, sg.nbuckets * tpb * 4 + 16)); sg.dirty = sg.fill + sg.nbuckets * tpb;
  CK(cudaMalloc(&sg.skey, ntiles * sg.cap * 4)); CK(cudaMalloc(&sg.sd[0], ntiles * sg.cap * 8)); CK(cudaMalloc(&sg.sd[1], ntiles * sg.cap * 8));
  const size_t sc = (size_t)sg.nbuckets * sg.nbatch_cap * sg.c1;
  CK(cudaMalloc(&sg.cnt1, (size_t)sg.nbuckets * sg.nbatch_cap * 4)); CK(cudaMalloc(&sg.key1, sc * 4)); CK(cudaMalloc(&sg.d1[0], sc * 8)); CK(cudaMalloc(&sg.d1[1], sc * 8));
  const size_t p1_smem = p1_smem_bytes<2>(sg.nbuckets << sg.c1_shift) + levtab_smem_bytes(g);
  const size_t p2_smem = p2_smem_bytes<2, 4096, 512>();
  CK(cudaFuncSetAttribute(k_part_obs<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p1_smem));
  CK(cudaFuncSetAttribute((k_part_tiles<2, 4096, 512>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p2_smem));
  const int nbatch = sg.nbatch;
  unsigned long long *buf; CK(cudaMalloc(&buf, (size_t)nbatch * 64)); CK(cudaMemset(buf, 0, (size_t)nbatch * 64));
  CK(cudaMemcpyToSymbol(g_phase_buf, &buf, sizeof(buf)));
  const size_t per_bucket = (size_t)nbatch * P1_BATCH / sg.nbuckets; const int parts = (int)std::max((size_t)1, (per_bucket + 3799) / 3800);
  const int g2 = std::min((int)sg.nbuckets * parts, 148 * 2);
  cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  float t1 = 0, t2 = 0; const int reps = 20;
  for (int r = 0; r < reps + 3; ++r) {
    CK(cudaMemsetAsync(sg.fill, 0, sg.nbuckets * tpb * 4));
    cudaEventRecord(e0);
    k_part_obs<2, 1><<<nbatch, P1_THREADS, p1_smem>>>(g, c, f, st, 1u, sg);
    cudaEventRecord(e1);
    k_part_tiles<2, 4096, 512><<<g2, 512, p2_smem>>>(g, sg, parts);
    cudaEventRecord(e2);
    CK(cudaDeviceSynchronize());
    float a, b; cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e1, e2);
    if (r >= 3) { t1 += a; t2 += b; }
  }
  printf("K1p %.2f us  K1q %.2f us  (n=%d, %d batches of %d, %d threads, buckets %u, tiles/bucket %zu, parts %d)\n", 1e3 * t1 / reps, 1e3 * t2 / reps, n,
         nbatch, P1_BATCH, P1_THREADS, sg.nbuckets, tpb, parts);
  std::vector<unsigned long long> hb((size_t)nbatch * 8);
  CK(cudaMemcpy(hb.data(), buf, hb.size() * 8, cudaMemcpyDeviceToHost));
  unsigned long long t0 = ~0ull, tend = 0;
  for (int b = 0; b < nbatch; ++b) { t0 = std::min(t0, hb[b * 8]); tend = std::max(tend, hb[b * 8 + 6]); }
  double ph[7] = {0}; double start_avg = 0, start_max = 0;
  for (int b = 0; b < nbatch; ++b) {
    for (int k = 1; k < 7; ++k) ph[k] += (double)(hb[b * 8 + k] - hb[b * 8 + k - 1]);
    const double s = (double)(hb[b * 8] - t0); start_avg += s; start_max = std::max(start_max, s);
  }
  printf("kernel span %.2f us; CTA start offset avg %.2f max %.2f us\n", (tend - t0) * 1e-3, start_avg / nbatch * 1e-3, start_max * 1e-3);
  const char *nm[7] = {"", "load_p2", "classify+ticket+bin", "-", "-", "writeout(2->5)", "end"};
  for (int k = 1; k < 7; ++k) printf("  %-28s %8.2f us avg per CTA\n", nm[k], ph[k] / nbatch * 1e-3);
  // histogram of CTA lifetimes
  std::vector<double> life(nbatch); for (int b = 0; b < nbatch; ++b) life[b] = (hb[b * 8 + 6] - hb[b * 8]) * 1e-3;
  std::sort(life.begin(), life.end());
  printf("CTA lifetime us: min %.2f p50 %.2f p90 %.2f max %.2f\n", life[0], life[nbatch / 2], life[nbatch * 9 / 10], life[nbatch - 1]);
  // ---- K1q phases (stamps of the last round of each CTA) ----
  CK(cudaMemset(buf, 0, (size_t)nbatch * 64));
  CK(cudaMemsetAsync(sg.fill, 0, sg.nbuckets * tpb * 4));
  k_part_obs<2, 1><<<nbatch, P1_THREADS, p1_smem>>>(g, c, f, st, 1u, sg);
  CK(cudaDeviceSynchronize());
  CK(cudaMemset(buf, 0, (size_t)nbatch * 64));
  k_part_tiles<2, 4096, 512><<<g2, 512, p2_smem>>>(g, sg, parts);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(hb.data(), buf, (size_t)g2 * 64, cudaMemcpyDeviceToHost));
  double q[6] = {0}; int cnt = 0;
  for (int b = 0; b < g2; ++b) { if (!hb[b * 8 + 5]) continue; ++cnt; for (int k = 1; k < 6; ++k) q[k] += (double)(hb[b * 8 + k] - hb[b * 8 + k - 1]); }
  const char *qn[6] = {"", "scan run counts", "pass1 count", "layout+tickets", "pass2 scatter", "writeout"};
  printf("K1q: %d CTAs of %d, parts %d\n", cnt, g2, parts);
  for (int k = 1; k < 6; ++k) printf("  %-20s %8.2f us avg (last round of a CTA)\n", qn[k], q[k] / std::max(cnt, 1) * 1e-3);
  return 0;
}
