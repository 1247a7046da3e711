// K3f — the discriminator's convolutional trunk as four fused kernels (fp32, grouped by expert).
//
// Reference: Discriminator.forward (expertsim/models/proton/discriminator.py:121-155) and DiscriminatorNeutron
// (expertsim/models/neutron/discriminator.py:11-48):
//     SN-Conv(1->32,k3) -> GroupNorm(8) -> LeakyReLU -> MaxPool2            "stem"
//     SN-Conv(32->16,k3) -> GroupNorm(8) -> LeakyReLU -> MaxPool (2,1)/(2,2) -> flatten || cond     "stage 2"
//
// The layer-by-layer build wrote the stem's 32 x 54 x 28 activation (198 MB at 1024 samples) four times per forward and
// read it as often per backward; the discriminator runs 4x forward and 4x backward per step, so 0.3 % of the step's FLOPs
// took a quarter of its time.  Here a GroupNorm group never leaves shared memory:
//   stem forward   CTA = (sample, GN group of 4 channels): conv -> statistics -> norm + LReLU + 2x2 max -> pooled map.
//                  Only the pooled map (48 KB / sample) and the statistics are written.
//   stem backward  same CTA shape; RE-COMPUTES the conv from the 6.7 KB image (9 MACs per value — cheaper than
//                  storing it), routes the pooled gradient to the arg-max (first maximum in window scan order, as
//                  torch), GroupNorm backward in shared memory, then weight / bias / affine gradients (one atomic per
//                  value per CTA) and the image gradient (atomics: the 8 groups of a sample meet in the image).
//   stage 2 fwd    CTA = sample: the 32-channel pooled map and the 16 x 288 weights live in shared memory, one thread
//                  per output pixel holds the 16 channel accumulators (weights are warp-broadcast float4 reads).
//   stage 2 bwd    CTA = `per` samples of one expert: GN backward, weight gradient (thread = (ci, ky) x 4 channels,
//                  3 kx taps through a sliding register window), data gradient (thread = input pixel x 32 channels).
#include "common.cuh"

namespace es {
namespace {

constexpr int kS2Threads = 384;

// ------------------------------------------------------------------------------------------------------ stem
// conv 1 -> 4 channels of one GroupNorm group, valid 3x3; y[c][p] = bias_c + sum_t w[c][t] img[window t]
__device__ __forceinline__ void stem_load_w(const float* __restrict__ w, const float* __restrict__ bias, int ch0,
                                            float (&wr)[4][9], float (&br)[4]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[c][t] = w[(ch0 + c) * 9 + t];
    br[c] = bias[ch0 + c];
  }
}

__device__ __forceinline__ float stem_conv(const float* __restrict__ s_img, int W, int Wo, int P,
                                           const float (&wr)[4][9], const float (&br)[4], float* __restrict__ s_y) {
  float s = 0.f;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const int oy = p / Wo, ox = p - oy * Wo;
    const float* ip = s_img + oy * W + ox;
    float v[9];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) v[ky * 3 + kx] = ip[ky * W + kx];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float a = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) a = fmaf(v[t], wr[c][t], a);
      a += br[c];
      s_y[c * P + p] = a;
      s += a;
    }
  }
  return s;
}

__global__ void __launch_bounds__(256)
disc_stem_fwd_kernel(const float* __restrict__ img, const float* __restrict__ w, long sw, const float* __restrict__ bias,
                     long sb, const float* __restrict__ gamma, const float* __restrict__ beta, long sn, int H, int W,
                     const es_group* __restrict__ grp, int n_groups, float* __restrict__ p1, float* __restrict__ stats) {
  extern __shared__ float sm[];
  __shared__ float red[32];
  const int r = blockIdx.x >> 3, gq = blockIdx.x & 7;
  const int gi = find_group(grp, n_groups, r);
  if (gi < 0) return;
  const int slot = grp[gi].slot;
  const int Ho = H - 2, Wo = W - 2, P = Ho * Wo, Hp = Ho / 2, Wp = Wo / 2, HW = H * W;
  float* s_img = sm;
  float* s_y = sm + ((HW + 3) & ~3);
  for (int i = threadIdx.x; i < HW; i += blockDim.x) s_img[i] = img[(size_t)r * HW + i];
  float wr[4][9], br[4];
  stem_load_w(w + slot * sw, bias + slot * sb, gq * 4, wr, br);
  __syncthreads();
  const float n = (float)(4 * P);
  const float mean = block_sum(stem_conv(s_img, W, Wo, P, wr, br, s_y), red) / n;
  float q = 0.f;
  for (int i = threadIdx.x; i < 4 * P; i += blockDim.x) { const float d = s_y[i] - mean; q += d * d; }
  const float rstd = rsqrtf(block_sum(q, red) / n + kNormEps);
  if (threadIdx.x == 0) { stats[(size_t)blockIdx.x * 2] = mean; stats[(size_t)blockIdx.x * 2 + 1] = rstd; }
  const int PP = Hp * Wp;
  for (int c = 0; c < 4; ++c) {
    const float ga = gamma[slot * sn + gq * 4 + c], be = beta[slot * sn + gq * 4 + c];
    float* dst = p1 + ((size_t)r * 32 + gq * 4 + c) * PP;
    const float* yc = s_y + c * P;
    for (int pp = threadIdx.x; pp < PP; pp += blockDim.x) {
      const int py = pp / Wp, px = pp - py * Wp;
      float best = -INFINITY;
#pragma unroll
      for (int ky = 0; ky < 2; ++ky)
#pragma unroll
        for (int kx = 0; kx < 2; ++kx) {
          const float a = lrelu((yc[(2 * py + ky) * Wo + 2 * px + kx] - mean) * rstd * ga + be);
          if (a > best) best = a;
        }
      dst[pp] = best;
    }
  }
}

template <bool WANT_W, bool WANT_DIMG>
__global__ void __launch_bounds__(256)
disc_stem_bwd_kernel(const float* __restrict__ dp1, const float* __restrict__ img, const float* __restrict__ w, long sw,
                     const float* __restrict__ bias, long sb, const float* __restrict__ gamma,
                     const float* __restrict__ beta, long sn, const float* __restrict__ stats, int H, int W,
                     const es_group* __restrict__ grp, int n_groups, float* __restrict__ d_img, float* __restrict__ dw,
                     long sdw, float* __restrict__ dbias, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  extern __shared__ float sm[];
  __shared__ float red[32];
  __shared__ float s_red[8][40];
  const int r = blockIdx.x >> 3, gq = blockIdx.x & 7;
  const int gi = find_group(grp, n_groups, r);
  if (gi < 0) return;
  const int slot = grp[gi].slot;
  const int Ho = H - 2, Wo = W - 2, P = Ho * Wo, Hp = Ho / 2, Wp = Wo / 2, HW = H * W, PP = Hp * Wp;
  float* s_img = sm;
  float* s_y = sm + ((HW + 3) & ~3);
  float* s_d = s_y + 4 * P;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) s_img[i] = img[(size_t)r * HW + i];
  for (int i = threadIdx.x; i < 4 * P; i += blockDim.x) s_d[i] = 0.f;
  float wr[4][9], br[4];
  stem_load_w(w + slot * sw, bias + slot * sb, gq * 4, wr, br);
  __syncthreads();
  stem_conv(s_img, W, Wo, P, wr, br, s_y);
  __syncthreads();
  const float mean = stats[(size_t)blockIdx.x * 2], rstd = stats[(size_t)blockIdx.x * 2 + 1];
  const float n = (float)(4 * P);
  // route the pooled gradient to the arg-max pixel (first maximum in (ky, kx) scan order) and through the LeakyReLU
  float gam[4], s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float ga = gamma[slot * sn + gq * 4 + c], be = beta[slot * sn + gq * 4 + c];
    gam[c] = ga;
    const float* dsrc = dp1 + ((size_t)r * 32 + gq * 4 + c) * PP;
    const float* yc = s_y + c * P;
    float* dc = s_d + c * P;
    float a1 = 0.f, b1 = 0.f;
    for (int pp = threadIdx.x; pp < PP; pp += blockDim.x) {
      const int py = pp / Wp, px = pp - py * Wp;
      float best = -INFINITY, bxh = 0.f, bpre = 0.f;
      int bpos = 0;
#pragma unroll
      for (int ky = 0; ky < 2; ++ky)
#pragma unroll
        for (int kx = 0; kx < 2; ++kx) {
          const int pos = (2 * py + ky) * Wo + 2 * px + kx;
          const float xh = (yc[pos] - mean) * rstd;
          const float pre = xh * ga + be;
          const float a = lrelu(pre);
          if (a > best) { best = a; bpos = pos; bxh = xh; bpre = pre; }
        }
      const float d = dsrc[pp] * (bpre > 0.f ? 1.f : kLReLU);
      dc[bpos] = d;
      a1 += d * bxh;
      b1 += d;
    }
    a1 = block_sum(a1, red);
    b1 = block_sum(b1, red);
    if (WANT_W && threadIdx.x == 0) {
      atomicAdd(&dgamma[slot * sn + gq * 4 + c], a1);
      atomicAdd(&dbeta[slot * sn + gq * 4 + c], b1);
    }
    s1 += b1 * ga;
    s2 += a1 * ga;
  }
  s1 /= n;
  s2 /= n;
  __syncthreads();
  // GroupNorm backward for every pixel; the conv weight / bias gradients ride along in registers
  float accw[4][9], accb[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    accb[c] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) accw[c][t] = 0.f;
  }
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    float v[9];
    if (WANT_W) {
      const int oy = p / Wo, ox = p - oy * Wo;
      const float* ip = s_img + oy * W + ox;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) v[ky * 3 + kx] = ip[ky * W + kx];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float xh = (s_y[c * P + p] - mean) * rstd;
      const float dx = rstd * (s_d[c * P + p] * gam[c] - s1 - xh * s2);
      s_d[c * P + p] = dx;
      if (WANT_W) {
        accb[c] += dx;
#pragma unroll
        for (int t = 0; t < 9; ++t) accw[c][t] = fmaf(dx, v[t], accw[c][t]);
      }
    }
  }
  if (WANT_W) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float x = warp_sum(accw[c][t]);
        if (lane == 0) s_red[warp][c * 9 + t] = x;
      }
      const float x = warp_sum(accb[c]);
      if (lane == 0) s_red[warp][36 + c] = x;
    }
  }
  __syncthreads();
  if (WANT_W && threadIdx.x < 40) {
    float x = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) x += s_red[k][threadIdx.x];
    if (threadIdx.x < 36) atomicAdd(&dw[slot * sdw + gq * 36 + threadIdx.x], x);
    else atomicAdd(&dbias[slot * sb + gq * 4 + threadIdx.x - 36], x);
  }
  if (WANT_DIMG) {
    for (int qd = threadIdx.x; qd < HW; qd += blockDim.x) {
      const int qy = qd / W, qx = qd - qy * W;
      float acc = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int oy = qy - ky;
        if (oy < 0 || oy >= Ho) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ox = qx - kx;
          if (ox < 0 || ox >= Wo) continue;
#pragma unroll
          for (int c = 0; c < 4; ++c) acc = fmaf(s_d[c * P + oy * Wo + ox], wr[c][ky * 3 + kx], acc);
        }
      }
      atomicAdd(&d_img[(size_t)r * HW + qd], acc);
    }
  }
}

// ------------------------------------------------------------------------------------------------------ stage 2
// smem layout shared by forward and backward: s_x [32][H1*W1] | s_w | ...
template <int PKW>
__global__ void __launch_bounds__(kS2Threads)
disc_stage2_fwd_kernel(const float* __restrict__ p1, const float* __restrict__ w, long sw, const float* __restrict__ bias,
                       long sb, const float* __restrict__ gamma, const float* __restrict__ beta, long sn,
                       const float* __restrict__ cond, int H1, int W1, const es_group* __restrict__ grp, int n_groups,
                       float* __restrict__ y2, float* __restrict__ stats, float* __restrict__ fcin, int ldf) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float s_st[8][2];
  const int r = blockIdx.x;
  const int gi = find_group(grp, n_groups, r);
  if (gi < 0) return;
  const int slot = grp[gi].slot;
  const int PX1 = H1 * W1, Ho = H1 - 2, Wo = W1 - 2, P2 = Ho * Wo;
  const int Hp = Ho / 2, Wp = Wo / PKW, flat = 16 * Hp * Wp;
  float* s_x = sm;                                   // [32][PX1]
  float* s_w = s_x + ((32 * PX1 + 3) & ~3);          // [288][16]
  float* s_y = s_w + 288 * 16;                       // [16][P2]
  const int tid = threadIdx.x;
  {
    const float4* src = reinterpret_cast<const float4*>(p1 + (size_t)r * 32 * PX1);   // 32*PX1 is a multiple of 4
    float4* dst = reinterpret_cast<float4*>(s_x);
    for (int i = tid; i < 32 * PX1 / 4; i += kS2Threads) dst[i] = src[i];
    const float* ws = w + slot * sw;
    for (int i = tid; i < 16 * 288; i += kS2Threads) {
      const int co = i / 288, k = i - co * 288;
      s_w[k * 16 + co] = ws[i];
    }
  }
  __syncthreads();
  for (int p = tid; p < P2; p += kS2Threads) {
    const int oy = p / Wo, ox = p - oy * Wo;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    const float* xb = s_x + oy * W1 + ox;
    for (int ci = 0; ci < 32; ++ci) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float v = xb[ci * PX1 + ky * W1 + kx];
          const float4* wv = reinterpret_cast<const float4*>(s_w + (ci * 9 + ky * 3 + kx) * 16);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 q = wv[j4];
            acc[j4 * 4 + 0] = fmaf(v, q.x, acc[j4 * 4 + 0]);
            acc[j4 * 4 + 1] = fmaf(v, q.y, acc[j4 * 4 + 1]);
            acc[j4 * 4 + 2] = fmaf(v, q.z, acc[j4 * 4 + 2]);
            acc[j4 * 4 + 3] = fmaf(v, q.w, acc[j4 * 4 + 3]);
          }
        }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float a = acc[j] + bias[slot * sb + j];
      s_y[j * P2 + p] = a;
      y2[((size_t)r * 16 + j) * P2 + p] = a;
    }
  }
  __syncthreads();
  // GroupNorm statistics: warp g < 8 owns group g = channels 2g, 2g+1 (contiguous 2*P2 values)
  const int lane = tid & 31, warp = tid >> 5;
  if (warp < 8) {
    const float* ys = s_y + warp * 2 * P2;
    const int n = 2 * P2;
    float s = 0.f;
    for (int i = lane; i < n; i += 32) s += ys[i];
    const float mean = warp_sum(s) / n;
    float q = 0.f;
    for (int i = lane; i < n; i += 32) { const float d = ys[i] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / n + kNormEps);
    if (lane == 0) {
      s_st[warp][0] = mean; s_st[warp][1] = rstd;
      stats[((size_t)r * 8 + warp) * 2] = mean; stats[((size_t)r * 8 + warp) * 2 + 1] = rstd;
    }
  }
  __syncthreads();
  const int PP = Hp * Wp;
  for (int i = tid; i < flat; i += kS2Threads) {
    const int c = i / PP, pp = i - c * PP, py = pp / Wp, px = pp - py * Wp;
    const float mean = s_st[c >> 1][0], rstd = s_st[c >> 1][1];
    const float ga = gamma[slot * sn + c], be = beta[slot * sn + c];
    float best = -INFINITY;
#pragma unroll
    for (int ky = 0; ky < 2; ++ky)
#pragma unroll
      for (int kx = 0; kx < PKW; ++kx) {
        const float a = lrelu((s_y[c * P2 + (2 * py + ky) * Wo + PKW * px + kx] - mean) * rstd * ga + be);
        if (a > best) best = a;
      }
    fcin[(size_t)r * ldf + i] = best;
  }
  if (tid < 9) fcin[(size_t)r * ldf + flat + tid] = cond[(size_t)r * 9 + tid];
}

template <int PKW, bool WANT_W>
__global__ void __launch_bounds__(kS2Threads)
disc_stage2_bwd_kernel(const float* __restrict__ dfc, int ldf, const float* __restrict__ y2, const float* __restrict__ stats,
                       const float* __restrict__ p1, const float* __restrict__ w, long sw, const float* __restrict__ gamma,
                       const float* __restrict__ beta, long sn, int H1, int W1, const es_group* __restrict__ grp,
                       int n_groups, int per, float* __restrict__ dp1, float* __restrict__ dw, long sdw,
                       float* __restrict__ dbias, long sb, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  extern __shared__ __align__(16) float sm[];
  int gi = -1, row0 = 0, ns = 0;
  {
    int cta = blockIdx.x;
    for (int i = 0; i < n_groups; ++i) {
      const int ch = ceil_div(grp[i].rows, per);
      if (cta < ch) { gi = i; row0 = grp[i].row_start + cta * per; ns = min(per, grp[i].rows - cta * per); break; }
      cta -= ch;
    }
  }
  if (gi < 0) return;
  const int slot = grp[gi].slot;
  const int PX1 = H1 * W1, Ho = H1 - 2, Wo = W1 - 2, P2 = Ho * Wo;
  const int Hp = Ho / 2, Wp = Wo / PKW, PP = Hp * Wp, flat = 16 * PP;
  float* s_x = sm;                                   // [32][PX1]         (weight gradient only)
  float* s_wd = s_x + ((32 * PX1 + 3) & ~3);         // [16*9][32]        w[co][ci][t] -> [(co*9+t)*32 + ci]
  float* s_dcp = s_wd + 144 * 32;                    // [16][P2]          gradient w.r.t. the conv output, channel-major
  float* s_dpc = s_dcp + ((16 * P2 + 3) & ~3);       // [P2][16]          same, pixel-major (weight gradient)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    const float* ws = w + slot * sw;
    for (int i = tid; i < 16 * 288; i += kS2Threads) {
      const int co = i / 288, k = i - co * 288, ci = k / 9, t = k - ci * 9;
      s_wd[(co * 9 + t) * 32 + ci] = ws[i];
    }
  }
  // weight-gradient ownership: thread = (ci, ky) x channel block of 4; 3 kx taps
  const int cb = tid & 3, cik = tid >> 2, wci = cik / 3, wky = cik - wci * 3;
  float accw[4][3];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int t = 0; t < 3; ++t) accw[j][t] = 0.f;

  for (int s = 0; s < ns; ++s) {
    const size_t r = row0 + s;
    __syncthreads();
    if (WANT_W) {
      const float4* src = reinterpret_cast<const float4*>(p1 + r * 32 * PX1);
      float4* dst = reinterpret_cast<float4*>(s_x);
      for (int i = tid; i < 32 * PX1 / 4; i += kS2Threads) dst[i] = src[i];
    }
    for (int i = tid; i < 16 * P2; i += kS2Threads) s_dcp[i] = 0.f;
    __syncthreads();
    // max-pool + LeakyReLU backward: gradient lands on the arg-max pixel of its window
    const float* yr = y2 + r * 16 * P2;
    for (int i = tid; i < flat; i += kS2Threads) {
      const int c = i / PP, pp = i - c * PP, py = pp / Wp, px = pp - py * Wp;
      const float mean = stats[(r * 8 + (c >> 1)) * 2], rstd = stats[(r * 8 + (c >> 1)) * 2 + 1];
      const float ga = gamma[slot * sn + c], be = beta[slot * sn + c];
      float best = -INFINITY, bpre = 0.f;
      int bpos = 0;
#pragma unroll
      for (int ky = 0; ky < 2; ++ky)
#pragma unroll
        for (int kx = 0; kx < PKW; ++kx) {
          const int pos = (2 * py + ky) * Wo + PKW * px + kx;
          const float pre = (yr[c * P2 + pos] - mean) * rstd * ga + be;
          const float a = lrelu(pre);
          if (a > best) { best = a; bpos = pos; bpre = pre; }
        }
      s_dcp[c * P2 + bpos] = dfc[r * ldf + i] * (bpre > 0.f ? 1.f : kLReLU);
    }
    __syncthreads();
    // GroupNorm backward, warp g < 8 owns group g
    if (warp < 8) {
      const float mean = stats[(r * 8 + warp) * 2], rstd = stats[(r * 8 + warp) * 2 + 1];
      float s1 = 0.f, s2 = 0.f;
      for (int cc = 0; cc < 2; ++cc) {
        const int c = warp * 2 + cc;
        const float ga = gamma[slot * sn + c];
        float a = 0.f, b = 0.f;
        for (int i = lane; i < P2; i += 32) {
          const float d = s_dcp[c * P2 + i];
          if (d != 0.f) { a += d * ((yr[c * P2 + i] - mean) * rstd); b += d; }
        }
        a = warp_sum(a);
        b = warp_sum(b);
        if (WANT_W && lane == 0) { atomicAdd(&dgamma[slot * sn + c], a); atomicAdd(&dbeta[slot * sn + c], b); }
        s1 += b * ga;
        s2 += a * ga;
      }
      s1 /= (float)(2 * P2);
      s2 /= (float)(2 * P2);
      for (int cc = 0; cc < 2; ++cc) {
        const int c = warp * 2 + cc;
        const float ga = gamma[slot * sn + c];
        float bs = 0.f;
        for (int i = lane; i < P2; i += 32) {
          const float xh = (yr[c * P2 + i] - mean) * rstd;
          const float dx = rstd * (s_dcp[c * P2 + i] * ga - s1 - xh * s2);
          s_dcp[c * P2 + i] = dx;
          s_dpc[i * 16 + c] = dx;
          bs += dx;
        }
        if (WANT_W) {
          bs = warp_sum(bs);
          if (lane == 0) atomicAdd(&dbias[slot * sb + c], bs);
        }
      }
    }
    __syncthreads();
    if (WANT_W) {
      // dW[co][ci][ky][kx] += sum_p dy[co][p] x[ci][oy+ky][ox+kx]
      const float* xr0 = s_x + wci * PX1 + wky * W1;
      for (int oy = 0; oy < Ho; ++oy) {
        const float* xr = xr0 + oy * W1;
        float x0 = xr[0], x1 = xr[1];
        const float4* dv = reinterpret_cast<const float4*>(s_dpc + oy * Wo * 16 + cb * 4);
        for (int ox = 0; ox < Wo; ++ox) {
          const float x2 = xr[ox + 2];
          const float4 d = dv[ox * 4];
          accw[0][0] = fmaf(d.x, x0, accw[0][0]); accw[0][1] = fmaf(d.x, x1, accw[0][1]); accw[0][2] = fmaf(d.x, x2, accw[0][2]);
          accw[1][0] = fmaf(d.y, x0, accw[1][0]); accw[1][1] = fmaf(d.y, x1, accw[1][1]); accw[1][2] = fmaf(d.y, x2, accw[1][2]);
          accw[2][0] = fmaf(d.z, x0, accw[2][0]); accw[2][1] = fmaf(d.z, x1, accw[2][1]); accw[2][2] = fmaf(d.z, x2, accw[2][2]);
          accw[3][0] = fmaf(d.w, x0, accw[3][0]); accw[3][1] = fmaf(d.w, x1, accw[3][1]); accw[3][2] = fmaf(d.w, x2, accw[3][2]);
          x0 = x1;
          x1 = x2;
        }
      }
    }
    // dp1[ci][q] = sum_{co,ky,kx} dy[co][qy-ky][qx-kx] w[co][ci][ky][kx]
    for (int qd = tid; qd < PX1; qd += kS2Threads) {
      const int qy = qd / W1, qx = qd - qy * W1;
      float acc[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = 0.f;
      for (int co = 0; co < 16; ++co) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int oy = qy - ky;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int ox = qx - kx;
            const bool ok = oy >= 0 && oy < Ho && ox >= 0 && ox < Wo;
            const float d = ok ? s_dcp[co * P2 + oy * Wo + ox] : 0.f;
            const float4* wv = reinterpret_cast<const float4*>(s_wd + (co * 9 + ky * 3 + kx) * 32);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 q = wv[j4];
              acc[j4 * 4 + 0] = fmaf(d, q.x, acc[j4 * 4 + 0]);
              acc[j4 * 4 + 1] = fmaf(d, q.y, acc[j4 * 4 + 1]);
              acc[j4 * 4 + 2] = fmaf(d, q.z, acc[j4 * 4 + 2]);
              acc[j4 * 4 + 3] = fmaf(d, q.w, acc[j4 * 4 + 3]);
            }
          }
        }
      }
      float* o = dp1 + r * 32 * PX1 + qd;
#pragma unroll
      for (int j = 0; j < 32; ++j) o[(size_t)j * PX1] = acc[j];
    }
  }
  if (WANT_W) {
    float* dws = dw + slot * sdw;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int t = 0; t < 3; ++t) atomicAdd(&dws[((cb * 4 + j) * 32 + wci) * 9 + wky * 3 + t], accw[j][t]);
  }
}

size_t stage2_fwd_smem(int H1, int W1) {
  const int PX1 = H1 * W1, P2 = (H1 - 2) * (W1 - 2);
  return ((size_t)((32 * PX1 + 3) & ~3) + 288 * 16 + 16 * P2) * sizeof(float);
}
size_t stage2_bwd_smem(int H1, int W1) {
  const int PX1 = H1 * W1, P2 = (H1 - 2) * (W1 - 2);
  return ((size_t)((32 * PX1 + 3) & ~3) + 144 * 32 + ((16 * P2 + 3) & ~3) + 16 * P2) * sizeof(float);
}

}  // namespace
}  // namespace es

using namespace es;

extern "C" int es_disc_stem_fwd(const float* img, const float* w, long slot_stride_w, const float* bias,
                                long slot_stride_b, const float* gamma, const float* beta, long slot_stride_n, int H,
                                int W, const es_group* grp, int n_groups, int total_rows, float* p1, float* stats,
                                void* stream) {
  ES_REQUIRE(img && w && bias && gamma && beta && grp && p1 && stats, "null pointer");
  ES_REQUIRE(H >= 4 && W >= 4 && n_groups >= 1 && n_groups <= kMaxGroups && total_rows > 0, "bad sizes");
  const size_t smem = ((size_t)((H * W + 3) & ~3) + 4 * (H - 2) * (W - 2)) * sizeof(float);
  ES_REQUIRE(smem <= 200 * 1024, "image too large for the fused stem");
  if (smem > 48 * 1024) ES_CUDA(cudaFuncSetAttribute(disc_stem_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  disc_stem_fwd_kernel<<<total_rows * 8, 256, smem, as_stream(stream)>>>(img, w, slot_stride_w, bias, slot_stride_b, gamma,
                                                                         beta, slot_stride_n, H, W, grp, n_groups, p1, stats);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_disc_stem_bwd(const float* dp1, const float* img, const float* w, long slot_stride_w, const float* bias,
                                long slot_stride_b, const float* gamma, const float* beta, long slot_stride_n,
                                const float* stats, int H, int W, const es_group* grp, int n_groups, int total_rows,
                                float* d_img, float* dw, long slot_stride_dw, float* dbias, float* dgamma, float* dbeta,
                                void* stream) {
  ES_REQUIRE(dp1 && img && w && bias && gamma && beta && stats && grp, "null pointer");
  ES_REQUIRE(H >= 4 && W >= 4 && n_groups >= 1 && n_groups <= kMaxGroups && total_rows > 0, "bad sizes");
  const bool want_w = dw != nullptr;
  ES_REQUIRE(!want_w || (dbias && dgamma && dbeta), "weight gradients need dbias, dgamma and dbeta too");
  ES_REQUIRE(want_w || d_img, "nothing to compute");
  const size_t smem = ((size_t)((H * W + 3) & ~3) + 8 * (H - 2) * (W - 2)) * sizeof(float);
  ES_REQUIRE(smem <= 200 * 1024, "image too large for the fused stem");
  cudaStream_t st = as_stream(stream);
#define ES_STEM_BWD(WW, DI)                                                                                            \
  {                                                                                                                    \
    if (smem > 48 * 1024) ES_CUDA(cudaFuncSetAttribute(disc_stem_bwd_kernel<WW, DI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    disc_stem_bwd_kernel<WW, DI><<<total_rows * 8, 256, smem, st>>>(dp1, img, w, slot_stride_w, bias, slot_stride_b, gamma, beta, \
        slot_stride_n, stats, H, W, grp, n_groups, d_img, dw, slot_stride_dw, dbias, dgamma, dbeta);                   \
  }
  if (want_w && d_img) ES_STEM_BWD(true, true)
  else if (want_w) ES_STEM_BWD(true, false)
  else ES_STEM_BWD(false, true)
#undef ES_STEM_BWD
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_disc_stage2_fwd(const float* p1, const float* w, long slot_stride_w, const float* bias,
                                  long slot_stride_b, const float* gamma, const float* beta, long slot_stride_n,
                                  const float* cond, int H1, int W1, int pool_kw, const es_group* grp, int n_groups,
                                  int total_rows, float* y2, float* stats, float* fcin, int ldf, void* stream) {
  ES_REQUIRE(p1 && w && bias && gamma && beta && cond && grp && y2 && stats && fcin, "null pointer");
  ES_REQUIRE(H1 >= 4 && W1 >= 4 && (H1 - 2) * (W1 - 2) <= kS2Threads * 4, "bad sizes");
  ES_REQUIRE(pool_kw == 1 || pool_kw == 2, "pool window must be (2,1) or (2,2)");
  ES_REQUIRE(n_groups >= 1 && n_groups <= kMaxGroups && total_rows > 0, "bad sizes");
  const int flat = 16 * ((H1 - 2) / 2) * ((W1 - 2) / pool_kw);
  ES_REQUIRE(ldf >= flat + 9, "fcin row too short");
  const size_t smem = stage2_fwd_smem(H1, W1);
  ES_REQUIRE(smem <= 220 * 1024, "pooled map too large for the fused stage");
  cudaStream_t st = as_stream(stream);
  if (pool_kw == 1) {
    ES_CUDA(cudaFuncSetAttribute(disc_stage2_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    disc_stage2_fwd_kernel<1><<<total_rows, kS2Threads, smem, st>>>(p1, w, slot_stride_w, bias, slot_stride_b, gamma, beta,
        slot_stride_n, cond, H1, W1, grp, n_groups, y2, stats, fcin, ldf);
  } else {
    ES_CUDA(cudaFuncSetAttribute(disc_stage2_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    disc_stage2_fwd_kernel<2><<<total_rows, kS2Threads, smem, st>>>(p1, w, slot_stride_w, bias, slot_stride_b, gamma, beta,
        slot_stride_n, cond, H1, W1, grp, n_groups, y2, stats, fcin, ldf);
  }
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_disc_stage2_bwd(const float* dfc, int ldf, const float* y2, const float* stats, const float* p1,
                                  const float* w, long slot_stride_w, const float* gamma, const float* beta,
                                  long slot_stride_n, int H1, int W1, int pool_kw, const es_group* grp, int n_groups,
                                  int total_rows, float* dp1, float* dw, long slot_stride_dw, float* dbias,
                                  long slot_stride_b, float* dgamma, float* dbeta, void* stream) {
  ES_REQUIRE(dfc && y2 && stats && p1 && w && gamma && beta && grp && dp1, "null pointer");
  ES_REQUIRE(H1 >= 4 && W1 >= 4 && pool_kw >= 1 && pool_kw <= 2, "bad sizes");
  ES_REQUIRE(n_groups >= 1 && n_groups <= kMaxGroups && total_rows > 0, "bad sizes");
  const bool want_w = dw != nullptr;
  ES_REQUIRE(!want_w || (dbias && dgamma && dbeta), "weight gradients need dbias, dgamma and dbeta too");
  const size_t smem = stage2_bwd_smem(H1, W1);
  ES_REQUIRE(smem <= 220 * 1024, "pooled map too large for the fused stage");
  const int per = 2;
  const int ctas = ceil_div(total_rows, per) + n_groups;
  cudaStream_t st = as_stream(stream);
#define ES_S2_BWD(KW, WW)                                                                                              \
  {                                                                                                                    \
    ES_CUDA(cudaFuncSetAttribute(disc_stage2_bwd_kernel<KW, WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    disc_stage2_bwd_kernel<KW, WW><<<ctas, kS2Threads, smem, st>>>(dfc, ldf, y2, stats, p1, w, slot_stride_w, gamma, beta, \
        slot_stride_n, H1, W1, grp, n_groups, per, dp1, dw, slot_stride_dw, dbias, slot_stride_b, dgamma, dbeta);      \
  }
  if (pool_kw == 1) { if (want_w) ES_S2_BWD(1, true) else ES_S2_BWD(1, false) }
  else { if (want_w) ES_S2_BWD(2, true) else ES_S2_BWD(2, false) }
#undef ES_S2_BWD
  ES_LAUNCH_CHECK();
  return ES_OK;
}
