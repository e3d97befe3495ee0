// Per-item statistics of the 3D clip core for profiles/clip_lane_sim.py (host build of csrc/nm_clip_core.cuh).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../non_matching_test_suite_b200/csrc/nm_clip_core.cuh"
using namespace nmclip;
int main(int argc, char **argv)
{
  std::FILE *f = std::fopen(argv[1], "rb");
  double nd = 0; if (std::fread(&nd, 8, 1, f) != 1) return 2;
  const size_t n = (size_t)nd;
  std::vector<double> in(n * 19);
  if (std::fread(in.data(), 8, in.size(), f) != in.size()) return 2;
  std::fclose(f);
  // per candidate: up to 2 items; output int8 triples: gate(0/1), planes, n
  std::vector<signed char> out(n * 8, 0);
  for (size_t i = 0; i < n; ++i) {
    const double *r = in.data() + i * 19;
    const double box[8] = {r[0], r[1], r[2], 0.0, r[3], r[4], r[5], 0.0};
    const double *quad = r + 6;
    const int code = (int)r[18];
    const double H[3] = {box[4] - box[0], box[5] - box[1], box[6] - box[2]};
    int k = 0;
    for (int t = 0; t < 4 && k < 2; ++t) {
      if (!((code >> t) & 1)) continue;
      const int iv[3] = {t == 0 ? 1 : 0, t <= 1 ? 2 : 1, t <= 2 ? 3 : 2};
      double tri[3][3];
      for (int v = 0; v < 3; ++v) for (int c = 0; c < 3; ++c) tri[v][c] = quad[iv[v] * 3 + c] - box[c];
      signed char *o = out.data() + i * 8 + k * 4;
      ++k;
      o[0] = tri_gate(tri, H) ? 1 : 0;
      o[3] = tri_box_overlap<false>(tri, H) ? 1 : 0;
      unsigned all = 63;
      for (int v = 0; v < 3; ++v) all &= outcode(tri[v][0], tri[v][1], tri[v][2], H);
      o[1] = (signed char)__builtin_popcount(~all & 63u);
      if (o[0]) { ClippedPoly C; o[2] = item_clip_polygon_v2(box, quad, t, C) ? (signed char)C.n : 0; }
    }
  }
  f = std::fopen(argv[2], "wb"); std::fwrite(out.data(), 1, out.size(), f); std::fclose(f);
  return 0;
}
