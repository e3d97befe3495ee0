// Host build of the 3D clip core (non_matching_test_suite_b200/csrc/nm_clip_core.cuh): what one candidate of
// clip_moments3_kernel computes -- gate, polygon per triangle of the split, closed-form moments, Q1 x P0 local vector --
// run on the CPU so that tests/test_clip_core_cpu.py can compare it pair by pair with the oracle without a GPU.
// Input (binary, doubles): n, then per candidate lo[3] hi[3] quad[12] code.  Output: per candidate sumw local[8] nfan dropped.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../non_matching_test_suite_b200/csrc/nm_clip_core.cuh"

using namespace nmclip;

#if HOST_POOL9
// the nine-slot pool of the GPU kernel (shared memory there), with a bound check: item_clip_polygon_v2 must never ask for more
struct Pool9 {
  double c[3][9];
  double get(int k, int i) const { if (i < 0 || i >= 9) { std::fprintf(stderr, "pool slot %d read\n", i); std::abort(); } return c[k][i]; }
  void set(int k, int i, double v) { if (i < 0 || i >= 9) { std::fprintf(stderr, "pool slot %d written\n", i); std::abort(); } c[k][i] = v; }
};
#endif

int main(int argc, char **argv)
{
  if (argc < 3) { std::fprintf(stderr, "usage: clip_core_host <in.bin> <out.bin>\n"); return 2; }
  std::FILE *f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  double nd = 0;
  if (std::fread(&nd, 8, 1, f) != 1) return 2;
  const size_t n = (size_t)nd;
  std::vector<double> in(n * 19), out(n * 11, 0.0);
  if (std::fread(in.data(), 8, in.size(), f) != in.size()) return 2;
  std::fclose(f);
  uint32_t lut[256];
  section_lut_build(lut);
  const double tol = argc > 3 ? std::atof(argv[3]) : 1e-15;
  for (size_t i = 0; i < n; ++i) {
    const double *r = in.data() + i * 19;
    const double box[8] = {r[0], r[1], r[2], 0.0, r[3], r[4], r[5], 0.0};
    const double *quad = r + 6;
    const int code = (int)r[18];
    const double H[3] = {box[4] - box[0], box[5] - box[1], box[6] - box[2]};
    double tot[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned nfan = 0, dropped = 0;
    for (int t = 0; t < 4; ++t) {
      if (!((code >> t) & 1)) continue;
      const int iv[3] = {t == 0 ? 1 : 0, t <= 1 ? 2 : 1, t <= 2 ? 3 : 2};
      double tri[3][3];
      for (int v = 0; v < 3; ++v)
        for (int k = 0; k < 3; ++k) tri[v][k] = quad[iv[v] * 3 + k] - box[k];
      if (!tri_gate(tri, H)) continue;
      ItemOut o;
      o.area2 = 0; o.nfan = 0; o.dropped = 0;
      for (int k = 0; k < 7; ++k) o.m[k] = 0;
#if HOST_POOL9
      ClippedPolyT<Pool9> C9;
      item_moments_pool(box, quad, t, tol, C9, o);
#else
      item_moments(box, quad, t, tol, lut, o);
#endif
      tot[0] += o.area2;
      for (int k = 0; k < 7; ++k) tot[k + 1] += o.m[k];
      nfan += o.nfan; dropped += o.dropped;
    }
    double *w = out.data() + i * 11;
    double sumw, l[8];
    moments_to_local(tot, sumw, l);
    w[0] = sumw;
    for (int k = 0; k < 8; ++k) w[1 + k] = sumw > tol ? l[k] : 0.0;
    w[9] = nfan; w[10] = dropped;
  }
  f = std::fopen(argv[2], "wb");
  std::fwrite(out.data(), 8, out.size(), f);
  std::fclose(f);
  return 0;
}
