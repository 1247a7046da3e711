// Offline preprocessing of the training set — the producers of two inputs of the hot path:
//   * `positions`: arg-max pixel (row, col) of every shower — get_max_value_image_coordinates
//     (expertsim/train/utils.py:81-82, looped over the data set in notebooks/calculate_and_analysis_of_max_coordinates.ipynb
//     cell 6): np.unravel_index(np.argmax(img), img.shape), i.e. the FIRST maximum in row-major order;
//   * `std`: per-condition-group pixel standard deviation (notebooks/calculating_diversity_for_data.ipynb cells 16-23):
//     samples with identical conditioning vectors form a group; per group and pixel the population standard deviation
//     (np.std, ddof = 0) over the group's showers; summed over the pixels; divided by the largest such sum.
// Both are HBM-bound single passes over the image set (6.7-7.7 KB per shower): one warp per image for the arg-max, one
// thread per (group, pixel) walking the group's showers for the moments (lanes along pixels: coalesced rows).
#include "common.cuh"

namespace es {
namespace {

__global__ void __launch_bounds__(256)
argmax_coords_kernel(const float* __restrict__ img, int rows, int HW, int W, int32_t* __restrict__ out_i,
                     float* __restrict__ out_f) {
  const int lane = threadIdx.x & 31, r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* p = img + (size_t)r * HW;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = lane; i < HW; i += 32) {
    const float v = p[i];
    if (v > best || (v != v && best == best)) { best = v; bi = i; }   // first NaN wins like numpy; indices ascend per lane
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    const bool onan = ob != ob, bnan = best != best;
    const bool take = (onan && !bnan) || (onan == bnan && (ob > best || ((ob == best || (onan && bnan)) && oi < bi)));
    if (take) { best = ob; bi = oi; }
  }
  if (lane == 0) {
    if (bi == 0x7fffffff) bi = 0;
    const int y = bi / W, x = bi - y * W;
    if (out_i) { out_i[2 * r] = y; out_i[2 * r + 1] = x; }
    if (out_f) { out_f[2 * r] = (float)y; out_f[2 * r + 1] = (float)x; }
  }
}

// order[seg[g] .. seg[g+1]) = indices of the showers of group g.  sums[g] += sum over this CTA's pixels of std_g(pixel).
__global__ void __launch_bounds__(256)
group_pixel_std_kernel(const float* __restrict__ img, int HW, const int32_t* __restrict__ order,
                       const int32_t* __restrict__ seg, double* __restrict__ sums) {
  __shared__ double red[8];
  const int g = blockIdx.x, px = blockIdx.y * blockDim.x + threadIdx.x;
  const int lo = seg[g], hi = seg[g + 1], n = hi - lo;
  double sd = 0.0;
  if (px < HW && n > 0) {
    double s = 0.0;
    for (int k = lo; k < hi; ++k) s += (double)img[(size_t)order[k] * HW + px];
    const double mean = s / n;
    double q = 0.0;
    for (int k = lo; k < hi; ++k) { const double d = (double)img[(size_t)order[k] * HW + px] - mean; q += d * d; }
    sd = sqrt(q / n);
  }
  sd = warp_sum_d(sd);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sd;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    atomicAdd(&sums[g], t);
  }
}

__global__ void __launch_bounds__(256)
group_std_finalize_kernel(const double* __restrict__ sums, int G, const int32_t* __restrict__ gid, int rows,
                          float* __restrict__ out) {
  __shared__ double red[8];
  double m = 0.0;
  for (int i = threadIdx.x; i < G; i += blockDim.x) m = fmax(m, sums[i]);   // every CTA recomputes the maximum (G is small)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, red[w]);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) out[i] = (float)(sums[gid[i]] / m);
}

}  // namespace
}  // namespace es

using namespace es;

extern "C" int es_argmax_coords(const float* img, int rows, int H, int W, int32_t* out_rowcol, float* out_rowcol_f32,
                                void* stream) {
  ES_REQUIRE(img && (out_rowcol || out_rowcol_f32) && rows > 0 && H > 0 && W > 0, "bad arguments");
  argmax_coords_kernel<<<ceil_div(rows, 8), 256, 0, as_stream(stream)>>>(img, rows, H * W, W, out_rowcol, out_rowcol_f32);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_group_pixel_std(const float* img, int rows, int HW, const int32_t* order, const int32_t* seg, int n_groups,
                                  const int32_t* group_of_row, double* group_sums, float* out_std, void* stream) {
  ES_REQUIRE(img && order && seg && group_of_row && group_sums && out_std && rows > 0 && HW > 0 && n_groups > 0 &&
                 n_groups <= 2147483647 / 2, "bad arguments");
  cudaStream_t st = as_stream(stream);
  ES_CUDA(cudaMemsetAsync(group_sums, 0, (size_t)n_groups * sizeof(double), st));
  ES_REQUIRE(ceil_div(HW, 256) <= 65535, "image too large");
  group_pixel_std_kernel<<<dim3(n_groups, ceil_div(HW, 256)), 256, 0, st>>>(img, HW, order, seg, group_sums);
  ES_LAUNCH_CHECK();
  group_std_finalize_kernel<<<ceil_div(rows, 256), 256, 0, st>>>(group_sums, n_groups, group_of_row, rows, out_std);
  ES_LAUNCH_CHECK();
  return ES_OK;
}
