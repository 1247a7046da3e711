// Helpers shared by the generator's non-GEMM kernels (two-pass row mapping, nearest-upsample fan-in).
#pragma once
#include "common.cuh"

namespace es {

// row r of a generator batch -> (group, pass, half-batch row j)
struct RowMap { int g, pass, j; };
__device__ __forceinline__ RowMap map_row(const es_group* grp, int E, int r, int two_pass) {
  RowMap m{-1, 0, 0};
  m.g = find_group(grp, E, r);
  if (m.g < 0) return m;
  const es_group G = grp[m.g];
  const int local = r - G.row_start;
  if (two_pass) {
    m.pass = local >= G.pass_rows ? 1 : 0;
    m.j = G.row_start / 2 + local - m.pass * G.pass_rows;
  } else {
    m.j = r;
  }
  return m;
}

// nearest-upsample fan-in tables: source index s receives upsampled indices [lo[s], hi[s])
__device__ __forceinline__ void build_fanin(int Ns, int Nu, int* lo, int* hi) {
  const float sc = (float)Ns / (float)Nu;
  for (int s = 0; s < Ns; ++s) { lo[s] = Nu; hi[s] = 0; }
  for (int u = 0; u < Nu; ++u) {
    int s = (int)floorf((float)u * sc);
    s = s < Ns - 1 ? s : Ns - 1;
    if (u < lo[s]) lo[s] = u;
    if (u + 1 > hi[s]) hi[s] = u + 1;
  }
}

// gradient arriving at source pixel (sy,sx), channels [c8, c8+8): sum over its upsample fan-out
__device__ __forceinline__ void load_da8(const __nv_bfloat16* __restrict__ dy_row, int Wu, int C, int c8, const int* ylo,
                                         const int* yhi, const int* xlo, const int* xhi, int sy, int sx, float* out) {
#pragma unroll
  for (int k = 0; k < 8; ++k) out[k] = 0.f;
  float f[8];
  for (int uy = ylo[sy]; uy < yhi[sy]; ++uy)
    for (int ux = xlo[sx]; ux < xhi[sx]; ++ux) {
      unpack8(__ldg(reinterpret_cast<const uint4*>(dy_row + ((size_t)uy * Wu + ux) * C + c8)), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) out[k] += f[k];
    }
}

}  // namespace es
