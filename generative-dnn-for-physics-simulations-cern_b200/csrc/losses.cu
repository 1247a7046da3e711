// K4 — fused loss tails of the MoE-GAN step: discriminator hinge loss, SDI-GAN diversity loss, photon-sum intensity
// loss and max-coordinate regression loss, with all gradients that enter the backward pass.
// Reference: MoEWrapper.discriminator_train_step / generator_train_step / sdi_gan_regularization /
// intensity_regularization (expertsim/models/moe.py:506-642) and AuxReg.regressor_loss (proton/aux_reg.py:42-45).
//
// One warp per sample row, float4 (128-bit) coalesced image reads, warp-shuffle reductions, one fp64 atomic per
// quantity per CTA.  Because  mean_over_B_e(x) * (B_e/B)  ==  sum(x)/B, only the SDI term needs the per-expert count,
// and per-expert partial sums are the only thing that has to cross ranks under data parallelism.
#include "common.cuh"

namespace es {

constexpr int kLat = 64, kZ = 10;
// sums[e][k]
enum { S_STD = 0, S_INVDIV = 1, S_SUM = 2, S_SUMSQ = 3, S_ABSERR = 4, S_COORD = 5, S_SCORE = 6, S_ROWS = 7 };

__global__ void __launch_bounds__(256)
hinge_d_kernel(const float* __restrict__ real, const float* __restrict__ fake, const es_group* __restrict__ grp,
               int B_global, float* __restrict__ d_real, float* __restrict__ d_fake, float* __restrict__ loss) {
  __shared__ float red[32];
  const es_group g = grp[blockIdx.x];
  const float invB = 1.f / (float)B_global;
  float acc = 0.f;
  for (int i = threadIdx.x; i < g.rows; i += blockDim.x) {
    const int r = g.row_start + i;
    const float a = 1.f - real[r], b = 1.f + fake[r];
    acc += fmaxf(a, 0.f) + fmaxf(b, 0.f);
    d_real[r] = a > 0.f ? -invB : 0.f;
    d_fake[r] = b > 0.f ? invB : 0.f;
  }
  const float t = block_sum(acc, red);
  if (threadIdx.x == 0) loss[blockIdx.x] = t * invB;
}

__device__ __forceinline__ float softplus_f(float x) {  // torch softplus, beta=1, threshold=20
  return x > 20.f ? x : log1pf(expf(x));
}

__global__ void __launch_bounds__(256)
gen_loss_reduce_kernel(const float* __restrict__ img, int HW, const float* __restrict__ lat1,
                       const float* __restrict__ lat2, const float* __restrict__ z1, const float* __restrict__ z2,
                       const float* __restrict__ stdv, const float* __restrict__ intensity,
                       const float* __restrict__ coords, const float* __restrict__ pos,
                       const float* __restrict__ score1, const es_group* __restrict__ grp, int E, int total_rows,
                       float* __restrict__ s_out, float* __restrict__ div_out, double* __restrict__ sums) {
  __shared__ float s_part[8][8];
  __shared__ int s_grp[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = blockIdx.x * 8 + warp;
  int g = -1;
  float part[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (r < total_rows) g = find_group(grp, E, r);
  if (g >= 0) {
    // photon sum: sum_hw (exp(img) - 1)
    const float4* p4 = reinterpret_cast<const float4*>(img + (size_t)r * HW);
    float s = 0.f;
    for (int i = lane; i < HW / 4; i += 32) {
      const float4 v = __ldg(p4 + i);
      s += (expf(v.x) - 1.f) + (expf(v.y) - 1.f) + (expf(v.z) - 1.f) + (expf(v.w) - 1.f);
    }
    s = warp_sum(s);
    // SDI ratio
    float a = fabsf(lat1[(size_t)r * kLat + lane] - lat2[(size_t)r * kLat + lane]) +
              fabsf(lat1[(size_t)r * kLat + 32 + lane] - lat2[(size_t)r * kLat + 32 + lane]);
    a = warp_sum(a) * (1.f / kLat);
    float n = lane < kZ ? fabsf(z1[(size_t)r * kZ + lane] - z2[(size_t)r * kZ + lane]) : 0.f;
    n = warp_sum(n) * (1.f / kZ);
    const float dv = a / (n + 1e-5f);
    // log-cosh coordinate loss terms: d + softplus(-2d) - ln 2
    float c = 0.f;
    if (lane < 2) {
      const float d = coords[(size_t)r * 2 + lane] - pos[(size_t)r * 2 + lane];
      c = d + softplus_f(-2.f * d) - 0.69314718055994531f;
    }
    c = warp_sum(c);
    if (lane == 0) {
      s_out[r] = s;
      div_out[r] = dv;
      part[S_STD] = stdv[r];
      part[S_INVDIV] = 1.f / (dv + 1e-5f);
      part[S_SUM] = s;
      part[S_SUMSQ] = s * s;
      part[S_ABSERR] = fabsf(s - intensity[r]);
      part[S_COORD] = c;
      part[S_SCORE] = score1[r];
      part[S_ROWS] = 1.f;
    }
  }
  if (lane == 0) {
    s_grp[warp] = g;
#pragma unroll
    for (int k = 0; k < 8; ++k) s_part[warp][k] = part[k];
  }
  __syncthreads();
  if (threadIdx.x < 8) {  // thread k merges runs of equal group, one fp64 atomic per run
    const int k = threadIdx.x;
    int cur = -1;
    double acc = 0.0;
    for (int w = 0; w < 8; ++w) {
      const int gw = s_grp[w];
      if (gw != cur) {
        if (cur >= 0) atomicAdd(&sums[cur * 8 + k], acc);
        cur = gw;
        acc = 0.0;
      }
      if (gw >= 0) acc += (k == S_SUMSQ) ? (double)s_part[w][S_SUM] * (double)s_part[w][S_SUM] : (double)s_part[w][k];
    }
    if (cur >= 0) atomicAdd(&sums[cur * 8 + k], acc);
  }
}

__global__ void __launch_bounds__(256)
gen_loss_grads_kernel(const float* __restrict__ img, int HW, const float* __restrict__ lat1,
                      const float* __restrict__ lat2, const float* __restrict__ z1, const float* __restrict__ z2,
                      const float* __restrict__ intensity, const float* __restrict__ coords,
                      const float* __restrict__ pos, const float* __restrict__ s_in, const float* __restrict__ div_in,
                      const es_group* __restrict__ grp, int E, int total_rows, const double* __restrict__ sums,
                      int B_global, float di_strength, float in_strength, float aux_strength,
                      float* __restrict__ d_score1, float* __restrict__ d_lat1, float* __restrict__ d_lat2,
                      float* __restrict__ d_coords, float* __restrict__ d_img, float* __restrict__ losses) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float invB = 1.f / (float)B_global;
  if (blockIdx.x == 0 && threadIdx.x < E && losses) {
    const int e = threadIdx.x;
    const double* S = sums + e * 8;
    const double n = S[S_ROWS];
    float out[6] = {0, 0, 0, 0, 0, 0};
    // n is the expert's (global, under data parallelism all-reduced) row count: 0 for skipped experts.  Not gated on this
    // rank's rows — a rank that holds none of a live expert's rows must report the same metrics as the others.
    if (n > 0.0) {
      const double mstd = S[S_STD] / n;
      const double div_l = mstd * mstd * (S[S_INVDIV] / n) * di_strength;
      const double int_l = S[S_ABSERR] / n * in_strength;
      const double aux_l = S[S_COORD] / (2.0 * n) * aux_strength;
      const double w = n / (double)B_global;
      out[0] = (float)((-S[S_SCORE] / n + div_l + int_l + aux_l) * w);
      out[1] = (float)div_l;
      out[2] = (float)int_l;
      out[3] = (float)aux_l;
      const double mean = S[S_SUM] / n;
      const double var = n > 1.0 ? (S[S_SUMSQ] - S[S_SUM] * mean) / (n - 1.0) : 0.0;
      out[4] = (float)sqrt(var > 0.0 ? var : 0.0);
      out[5] = (float)mean;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) losses[e * 6 + k] = out[k];
  }
  const int r = blockIdx.x * 8 + warp;
  if (r >= total_rows) return;
  const int g = find_group(grp, E, r);
  if (g < 0) return;
  const double n = sums[g * 8 + S_ROWS];
  const float mstd = (float)(sums[g * 8 + S_STD] / n);
  if (lane == 0) d_score1[r] = -invB;
  // SDI: L*w = mstd^2 * k / B * sum_b 1/(div_b + eps)
  const float dv = div_in[r];
  float nz = lane < kZ ? fabsf(z1[(size_t)r * kZ + lane] - z2[(size_t)r * kZ + lane]) : 0.f;
  nz = warp_sum(nz) * (1.f / kZ);
  const float dL_ddiv = -mstd * mstd * di_strength * invB / ((dv + 1e-5f) * (dv + 1e-5f));
  const float dL_da = dL_ddiv / (nz + 1e-5f);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = lane + 32 * h;
    const float d = lat1[(size_t)r * kLat + j] - lat2[(size_t)r * kLat + j];
    const float sg = (d > 0.f) - (d < 0.f);
    const float v = dL_da * sg * (1.f / kLat);
    d_lat1[(size_t)r * kLat + j] = v;
    d_lat2[(size_t)r * kLat + j] = -v;
  }
  if (lane < 2) {
    const float d = coords[(size_t)r * 2 + lane] - pos[(size_t)r * 2 + lane];
    d_coords[(size_t)r * 2 + lane] = aux_strength * tanhf(d) * 0.5f * invB;
  }
  // intensity: d/d img = in_strength * sign(s - I) / B * exp(img)
  const float diff = s_in[r] - intensity[r];
  const float coef = in_strength * ((diff > 0.f) - (diff < 0.f)) * invB;
  const float4* p4 = reinterpret_cast<const float4*>(img + (size_t)r * HW);
  float4* g4 = reinterpret_cast<float4*>(d_img + (size_t)r * HW);
  for (int i = lane; i < HW / 4; i += 32) {
    const float4 v = __ldg(p4 + i);
    float4 o = g4[i];
    o.x += coef * expf(v.x);
    o.y += coef * expf(v.y);
    o.z += coef * expf(v.z);
    o.w += coef * expf(v.w);
    g4[i] = o;
  }
}

// gradient of the coordinate-regression loss alone (it needs nothing but the regressor's own output and the global batch
// size), so the auxiliary regressor's backward can start before the discriminator passes of the generator step finish
__global__ void aux_loss_grad_kernel(const float* __restrict__ coords, const float* __restrict__ pos,
                                     const es_group* __restrict__ grp, int E, int total_rows, float scale,
                                     float* __restrict__ d_coords) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_rows * 2) return;
  const int g = find_group(grp, E, i >> 1);
  d_coords[i] = g < 0 ? 0.f : scale * tanhf(coords[i] - pos[i]);
}

__global__ void expm1_scatter_kernel(const float* __restrict__ img, const int32_t* __restrict__ perm, int HW,
                                     double* __restrict__ out64, float* __restrict__ out32) {
  const int r = blockIdx.x;
  const int dst = perm ? perm[r] : r;
  const float* s = img + (size_t)r * HW;
  if ((HW & 3) == 0) {      // 16-byte accesses (rows are 16-byte aligned then): the scalar loop ran at 0.45 of the HBM rate
    const float4* s4 = reinterpret_cast<const float4*>(s);
    for (int i = threadIdx.x; i < HW / 4; i += blockDim.x) {
      const float4 q = __ldg(s4 + i);
      const float4 v = make_float4(expm1f(q.x), expm1f(q.y), expm1f(q.z), expm1f(q.w));
      if (out64) {
        double2* o = reinterpret_cast<double2*>(out64 + (size_t)dst * HW) + 2 * i;
        o[0] = make_double2((double)v.x, (double)v.y);
        o[1] = make_double2((double)v.z, (double)v.w);
      }
      if (out32) reinterpret_cast<float4*>(out32 + (size_t)dst * HW)[i] = v;
    }
    return;
  }
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const float v = expm1f(s[i]);
    if (out64) out64[(size_t)dst * HW + i] = (double)v;
    if (out32) out32[(size_t)dst * HW + i] = v;
  }
}

// Evaluation metric front end (reference train/utils.py:18-78): the five "channel" sums of a shower —
// the checkerboard cells (i%2 != j%2) of the bottom-left, bottom-right, top-left and top-right quadrants, and the
// complementary checkerboard of the whole image — with the expm1 of the batch-inference tail fused in.  One warp per image,
// fp64 accumulation (the reference sums float64 copies on the host).
__global__ void __launch_bounds__(256)
channel_sums_kernel(const float* __restrict__ img, int H, int W, int rows, int apply_expm1, double* __restrict__ out) {
  const int lane = threadIdx.x & 31, r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* p = img + (size_t)r * H * W;
  const int mr = H / 2, mc = W / 2;
  double s[5] = {0, 0, 0, 0, 0};
  for (int i = lane; i < H * W; i += 32) {
    const int y = i / W, x = i - y * W;
    const float v = apply_expm1 ? expm1f(p[i]) : p[i];
    if ((y & 1) != (x & 1)) s[(y >= mr ? 0 : 2) + (x >= mc ? 1 : 0)] += (double)v;
    else s[4] += (double)v;
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const double t = warp_sum_d(s[k]);
    if (lane == 0) out[(size_t)r * 5 + k] = t;
  }
}

// 1-D Wasserstein distance between two equally sized, already sorted samples: mean |a_i - b_i|
__global__ void __launch_bounds__(256)
w1_sorted_kernel(const double* __restrict__ a, const double* __restrict__ b, int n, int stride, double* __restrict__ out) {
  __shared__ double red[8];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += fabs(a[(size_t)i * stride] - b[(size_t)i * stride]);
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    out[blockIdx.x] = t / (double)n;
  }
}

}  // namespace es

using namespace es;

extern "C" int es_channel_sums(const float* img, int H, int W, int rows, int apply_expm1, double* out5, void* stream) {
  ES_REQUIRE(img && out5 && H > 1 && W > 1 && rows > 0, "bad arguments");
  channel_sums_kernel<<<ceil_div(rows, 8), 256, 0, as_stream(stream)>>>(img, H, W, rows, apply_expm1, out5);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_w1_sorted(const double* a, const double* b, int n, int n_cols, double* out, void* stream) {
  ES_REQUIRE(a && b && out && n > 0 && n_cols >= 1, "bad arguments");
  for (int c = 0; c < n_cols; ++c)   // column c of row-major [n, n_cols] arrays
    w1_sorted_kernel<<<1, 256, 0, as_stream(stream)>>>(a + c, b + c, n, n_cols, out + c);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_hinge_d(const float* real_score, const float* fake_score, const es_group* grp, int E,
                          const float* counts_global, int B_global, float* d_real, float* d_fake, float* loss,
                          void* stream) {
  (void)counts_global;  // mean_{B_e}(x) * B_e/B == sum(x)/B: the count cancels
  ES_REQUIRE(real_score && fake_score && grp && d_real && d_fake && loss, "null pointer");
  ES_REQUIRE(E >= 1 && E <= kMaxGroups && B_global > 0, "bad sizes");
  hinge_d_kernel<<<E, 256, 0, as_stream(stream)>>>(real_score, fake_score, grp, B_global, d_real, d_fake, loss);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_gen_loss_reduce(const float* img, int HW, const float* lat1, const float* lat2, const float* z1,
                                  const float* z2, const float* stdv, const float* intensity, const float* coords,
                                  const float* pos, const float* score1, const es_group* grp, int E, int total_rows,
                                  float* s_out, float* div_out, double* sums, void* stream) {
  ES_REQUIRE(img && lat1 && lat2 && z1 && z2 && stdv && intensity && coords && pos && score1 && grp, "null input");
  ES_REQUIRE(s_out && div_out && sums, "null output");
  ES_REQUIRE(HW > 0 && HW % 4 == 0 && E >= 1 && E <= kMaxGroups && total_rows > 0, "bad sizes");
  ES_CUDA(cudaMemsetAsync(sums, 0, (size_t)E * 8 * sizeof(double), as_stream(stream)));
  gen_loss_reduce_kernel<<<ceil_div(total_rows, 8), 256, 0, as_stream(stream)>>>(
      img, HW, lat1, lat2, z1, z2, stdv, intensity, coords, pos, score1, grp, E, total_rows, s_out, div_out, sums);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_gen_loss_grads(const float* img, int HW, const float* lat1, const float* lat2, const float* z1,
                                 const float* z2, const float* stdv, const float* intensity, const float* coords,
                                 const float* pos, const float* s, const float* divv, const es_group* grp, int E,
                                 int total_rows, const double* sums, int B_global, float di_strength,
                                 float in_strength, float aux_strength, float* d_score1, float* d_lat1, float* d_lat2,
                                 float* d_coords, float* d_img, float* losses, void* stream) {
  (void)stdv;
  ES_REQUIRE(img && lat1 && lat2 && z1 && z2 && intensity && coords && pos && s && divv && grp && sums, "null input");
  ES_REQUIRE(d_score1 && d_lat1 && d_lat2 && d_coords && d_img, "null output");
  ES_REQUIRE(HW > 0 && HW % 4 == 0 && E >= 1 && E <= 32 && total_rows > 0 && B_global > 0, "bad sizes");
  gen_loss_grads_kernel<<<ceil_div(total_rows, 8), 256, 0, as_stream(stream)>>>(
      img, HW, lat1, lat2, z1, z2, intensity, coords, pos, s, divv, grp, E, total_rows, sums, B_global, di_strength,
      in_strength, aux_strength, d_score1, d_lat1, d_lat2, d_coords, d_img, losses);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_aux_loss_grad(const float* coords, const float* pos, const es_group* grp, int E, int total_rows,
                                int B_global, float aux_strength, float* d_coords, void* stream) {
  ES_REQUIRE(coords && pos && grp && d_coords && E >= 1 && E <= kMaxGroups && total_rows > 0 && B_global > 0, "bad arguments");
  aux_loss_grad_kernel<<<ceil_div(total_rows * 2, 256), 256, 0, as_stream(stream)>>>(
      coords, pos, grp, E, total_rows, aux_strength * 0.5f / (float)B_global, d_coords);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_expm1_scatter(const float* img, const int32_t* perm, int rows, int HW, double* out_f64,
                                float* out_f32, void* stream) {
  ES_REQUIRE(img && rows > 0 && HW > 0 && (out_f64 || out_f32), "bad arguments");
  expm1_scatter_kernel<<<rows, 256, 0, as_stream(stream)>>>(img, perm, HW, out_f64, out_f32);
  ES_LAUNCH_CHECK();
  return ES_OK;
}
