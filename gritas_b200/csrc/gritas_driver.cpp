// gritas_driver.cpp -- a small C++ stand-in for the caller of the hot path, Program GrITAS
// (src/Applications/GrITAS_App/gritas.f:389-808 file loop, :885-887 Norm), written against
// m_gritas_grids.hpp so that the boundary is exercised by compiled host code the way the Fortran driver
// would.  It is NOT the reference's CLI: no rc parsing, no ODS/NetCDF, no GrADS writer.
//
// usage: gritas_b200_driver <setup.bin> <out.bin> [-det] [-oma | -omf] [-nopassive] file1.sods file2.sods ...
//   setup.bin : int32 im jm nlevs KXMAX KTMAX listsz, then plevs[nlevs] f64, kt_list[listsz] i32,
//               kx_list[KXMAX*listsz] i32 (column-major, negative terminator), nsurf i32, ktSurfAll[nsurf] i32
//   *.sods    : one synoptic time of ODS columns, flat: int32 n, then lat lon lev f64[n], time kt kx ks qcexcl
//               i32[n], obs omf oma xvec f64[n]   (gritas_b200/synth.py write_flat)
//   out.bin   : int32 l2d l3d nr, then bias rtms stdv xcov(ir=1) nobs for the 3-D grids (f64, reference layout)
#include <cstdint>
#include <cstring>
#include <string>

#include "m_gritas_grids.hpp"

using namespace m_gritas_grids;

template <typename T>
static std::vector<T> rd(FILE *f, size_t n) {
  std::vector<T> v(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) gritas_perror("driver: short read");
  return v;
}

int main(int argc, char **argv) {
  if (argc < 4) { std::printf("usage: %s setup.bin out.bin [-det] [-omf|-oma] [-nopassive] files...\n", argv[0]); return 2; }
  int iopt = 1, ioptn = 1;                        // -omf (gritas.f:1408-1481: odd iopt -> QC filter runs)
  std::vector<std::string> files;
  for (int a = 3; a < argc; ++a) {
    if (!strcmp(argv[a], "-det")) b200_mode = GRITAS_B200_MODE_DETERMINISTIC;
    else if (!strcmp(argv[a], "-oma")) iopt = 3;
    else if (!strcmp(argv[a], "-omf")) iopt = 1;
    else if (!strcmp(argv[a], "-nopassive")) ioptn = -1;
    else files.push_back(argv[a]);
  }
  ioptn *= iopt;
  FILE *fs = fopen(argv[1], "rb");
  if (!fs) gritas_perror("driver: cannot open setup");
  const auto hd = rd<int32_t>(fs, 6);
  const int KXMAX = hd[3], KTMAX = hd[4], listsz = hd[5];
  nlevs = hd[2];
  plevs = rd<double>(fs, (size_t)nlevs);
  const auto kt_list = rd<int32_t>(fs, (size_t)listsz);
  const auto kx_list = rd<int32_t>(fs, (size_t)KXMAX * listsz);
  const int nsurf = rd<int32_t>(fs, 1)[0];
  const auto ktSurfAll = rd<int32_t>(fs, (size_t)nsurf);
  fclose(fs);
  grids_init(hd[0], hd[1], nlevs);                                  // gritas.f:1885-1889 (km = nlevs)

  std::vector<int> kxkt_table((size_t)KXMAX * KTMAX), list_2d(64), list_3d(4096);
  int l2d = 0, l3d = 0;
  GrITAS_KxKt(KTMAX, KXMAX, 64, 4096, listsz, kt_list.data(), kx_list.data(), ktSurfAll.data(), nsurf, kxkt_table.data(),
              list_2d.data(), l2d, list_3d.data(), l3d);               // gritas.f:310
  std::vector<double> p1((size_t)nlevs), p2((size_t)nlevs);
  GrITAS_Pint(p1.data(), p2.data());                                  // gritas.f:378
  const int nr = 1;
  const size_t n3 = (size_t)im * jm * km * (l3d > 0 ? l3d : 1), n2 = (size_t)im * jm * (l2d > 0 ? l2d : 1);
  std::vector<double> bias_3d(n3 * nr), rtms_3d(n3 * nr), stdv_3d(n3 * nr), xcov_3d(n3 * nr), nobs_3d(n3);   // gritas.f:336-363
  std::vector<double> bias_2d(n2 * nr), rtms_2d(n2 * nr), stdv_2d(n2 * nr), xcov_2d(n2 * nr), nobs_2d(n2);

  const int X_PASSIVE = 7, X_PRE_BAD = 1;                            // m_odsmeta (not in the reference tree)
  for (const std::string &fn : files) {                              // gritas.f:389 file loop (one synoptic time each)
    FILE *f = fopen(fn.c_str(), "rb");
    if (!f) gritas_perror("driver: cannot open ODS columns");
    const int nobs = rd<int32_t>(f, 1)[0];
    const auto lat = rd<double>(f, nobs), lon = rd<double>(f, nobs), lev = rd<double>(f, nobs);
    const auto time = rd<int32_t>(f, nobs), kt = rd<int32_t>(f, nobs), kx = rd<int32_t>(f, nobs), ks = rd<int32_t>(f, nobs),
               qcexcl = rd<int32_t>(f, nobs);
    const auto obs = rd<double>(f, nobs), omf = rd<double>(f, nobs), oma = rd<double>(f, nobs), xvec = rd<double>(f, nobs);
    fclose(f);
    (void)time; (void)ks; (void)obs; (void)xvec;
    if (nobs == 0) continue;                                         // gritas.f:464-471
    std::vector<double> del(iopt == 1 ? omf : oma);                  // gritas.f:562-587
    std::vector<int> ob_flag((size_t)nobs, 0);                       // gritas.f:551
    for (int i = 0; i < nobs; ++i) {                                 // gritas.f:710-739 (no -time / -lon window)
      const bool iampassive = qcexcl[i] == X_PASSIVE || qcexcl[i] == X_PRE_BAD;
      const bool iwantpassive = ioptn > 0 && iampassive;
      ob_flag[i] = (qcexcl[i] == 0 || iwantpassive) ? 0 : 1;
    }
    GrITAS_Accum(lat.data(), lon.data(), lev.data(), kx.data(), kt.data(), del.data(), nobs, nr, kxkt_table.data(), KXMAX,
                 KTMAX, p1.data(), p2.data(), ob_flag.data(), l2d, bias_2d.data(), rtms_2d.data(), xcov_2d.data(),
                 nobs_2d.data(), l3d, bias_3d.data(), rtms_3d.data(), xcov_3d.data(), nobs_3d.data());   // gritas.f:777-781
  }
  GrITAS_Norm(nr, true, l2d, l2d ? bias_2d.data() : nullptr, l2d ? rtms_2d.data() : nullptr, l2d ? stdv_2d.data() : nullptr,
              l2d ? xcov_2d.data() : nullptr, l2d ? nobs_2d.data() : nullptr, l3d, l3d ? bias_3d.data() : nullptr,
              l3d ? rtms_3d.data() : nullptr, l3d ? stdv_3d.data() : nullptr, l3d ? xcov_3d.data() : nullptr,
              l3d ? nobs_3d.data() : nullptr);                        // gritas.f:885-887
  FILE *fo = fopen(argv[2], "wb");
  if (!fo) gritas_perror("driver: cannot open output");
  const int32_t oh[3] = {l2d, l3d, nr};
  fwrite(oh, 4, 3, fo);
  for (const auto *v : {&bias_3d, &rtms_3d, &stdv_3d, &xcov_3d, &nobs_3d}) fwrite(v->data(), 8, v->size(), fo);
  fclose(fo);
  GrITAS_Clean();
  std::printf("gritas_b200_driver: %zu files, %d upper-air grids, %d surface grids\n", files.size(), l3d, l2d);
  return 0;
}
