// K3 — building blocks of the discriminator and the auxiliary coordinate regressor, forward and backward
// (fp32, NCHW, grouped by expert through the es_group table).  Reference: Discriminator.forward
// (expertsim/models/proton/discriminator.py:121-155), AuxReg / FeatureExtractor / ResidualBlock
// (expertsim/models/proton/aux_reg.py:11-131) and the hook-based spectral norm of torch.
//
// The convolutions live in conv_simt.cu (fp32 SIMT implicit GEMMs); this file keeps the few-tap weight-gradient kernel,
// the norms, pools, linears, spectral norm, Adam and the elementwise helpers.
#include "common.cuh"

namespace es {

constexpr int kCoT = 8;

// (group, chunk) scheduler for kernels whose CTA handles `per` consecutive rows of ONE group
__device__ __forceinline__ bool chunk_of(const es_group* grp, int n_groups, int per, int cta, int& g, int& row0, int& nrows) {
  for (int i = 0; i < n_groups; ++i) {
    const int ch = ceil_div(grp[i].rows, per);
    if (cta < ch) {
      g = i;
      row0 = grp[i].row_start + cta * per;
      nrows = min(per, grp[i].rows - cta * per);
      return true;
    }
    cta -= ch;
  }
  return false;
}

// ------------------------------------------------------------------------------------------------ conv data gradient, single input channel (direct)
// STAGED = false: the gradient sample does not fit in shared memory (neutron aux conv1: 32 x 42 x 42 floats); dy is then
// read straight from global memory / L2 and only the weight tile is staged.
template <int CIT, bool STAGED = true>
__global__ void __launch_bounds__(256)
conv2d_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ w, long sw, es_conv2d g,
                       const es_group* __restrict__ grp, int n_groups, int S, float* __restrict__ dx, int accumulate) {
  extern __shared__ float sm[];
  int gi, row0, ns;
  if (!chunk_of(grp, n_groups, S, blockIdx.x, gi, row0, ns)) return;
  const int slot = grp[gi].slot, ci0 = blockIdx.y * CIT;
  const int out_sz = g.Co * g.Ho * g.Wo, ktaps = g.KH * g.KW, HWi = g.Hi * g.Wi;
  float* s_dy = sm;                    // [S][Co][Ho][Wo]
  float* s_w = sm + (STAGED ? S * out_sz : 0);        // [Co][KH][KW][CIT]
  if (STAGED) for (int i = threadIdx.x; i < ns * out_sz; i += blockDim.x) s_dy[i] = dy[(size_t)row0 * out_sz + i];
  for (int i = threadIdx.x; i < g.Co * ktaps * CIT; i += blockDim.x) {
    const int c = i % CIT, t = (i / CIT) % ktaps, co = i / (CIT * ktaps);
    s_w[i] = w[slot * sw + ((size_t)co * g.Ci + ci0 + c) * ktaps + t];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < ns * HWi; idx += blockDim.x) {
    const int s = idx / HWi, p = idx % HWi, iy = p / g.Wi, ix = p % g.Wi;
    float acc[CIT];
#pragma unroll
    for (int c = 0; c < CIT; ++c) acc[c] = 0.f;
    const float* ds = STAGED ? s_dy + s * out_sz : dy + (size_t)(row0 + s) * out_sz;
    for (int ky = 0; ky < g.KH; ++ky) {
      const int ty = iy + g.pad - ky;
      if (ty < 0 || ty % g.stride != 0) continue;
      const int oy = ty / g.stride;
      if (oy >= g.Ho) continue;
      for (int kx = 0; kx < g.KW; ++kx) {
        const int tx = ix + g.pad - kx;
        if (tx < 0 || tx % g.stride != 0) continue;
        const int ox = tx / g.stride;
        if (ox >= g.Wo) continue;
        for (int co = 0; co < g.Co; ++co) {
          const float d = ds[(co * g.Ho + oy) * g.Wo + ox];
          const float* wv = s_w + ((co * g.KH + ky) * g.KW + kx) * CIT;
#pragma unroll
          for (int c = 0; c < CIT; ++c) acc[c] = fmaf(d, wv[c], acc[c]);
        }
      }
    }
    float* o = dx + ((size_t)(row0 + s) * g.Ci + ci0) * HWi + p;
#pragma unroll
    for (int c = 0; c < CIT; ++c) {
      if (accumulate) o[(size_t)c * HWi] += acc[c];
      else o[(size_t)c * HWi] = acc[c];
    }
  }
}

// Few-tap variant (Ci*KH*KW <= 9, e.g. the discriminator's first conv 1->32 k3): the generic kernel gives one thread per
// tap, i.e. 9 busy threads.  Here threads own PIXELS, keep all T x 8 partial sums in registers over the CTA's samples and
// combine them once at the end (warp shuffles + shared memory), then one atomic per weight per CTA.
template <int T>
__global__ void __launch_bounds__(256)
conv2d_bwd_weight_fewtaps_kernel(const float* __restrict__ x, const float* __restrict__ dy, es_conv2d g,
                                 const es_group* __restrict__ grp, int n_groups, int per, float* __restrict__ dw,
                                 float* __restrict__ db, long sw, long sb) {
  extern __shared__ float sm[];
  __shared__ float s_red[8][(T + 1) * kCoT];
  int gi, row0, ns;
  if (!chunk_of(grp, n_groups, per, blockIdx.x, gi, row0, ns)) return;
  const int slot = grp[gi].slot, co0 = blockIdx.y * kCoT;
  const int in_sz = g.Ci * g.Hi * g.Wi, HWo = g.Ho * g.Wo, taps = g.Ci * g.KH * g.KW;
  float* s_x = sm;
  float* s_d = sm + ((in_sz + 3) & ~3);
  float acc[T + 1][kCoT];
#pragma unroll
  for (int j = 0; j <= T; ++j)
#pragma unroll
    for (int c = 0; c < kCoT; ++c) acc[j][c] = 0.f;
  for (int s = 0; s < ns; ++s) {
    __syncthreads();
    const size_t row = row0 + s;
    for (int i = threadIdx.x; i < in_sz; i += blockDim.x) s_x[i] = x[row * in_sz + i];
    for (int i = threadIdx.x; i < HWo * kCoT; i += blockDim.x) {
      const int c = i / HWo, p = i % HWo;
      s_d[p * kCoT + c] = dy[(row * g.Co + co0 + c) * HWo + p];
    }
    __syncthreads();
    for (int p = threadIdx.x; p < HWo; p += blockDim.x) {
      const int oy = p / g.Wo, ox = p - oy * g.Wo;
      const float4* dv = reinterpret_cast<const float4*>(s_d + p * kCoT);
      const float4 d0 = dv[0], d1 = dv[1];
      const float d[kCoT] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int c = 0; c < kCoT; ++c) acc[T][c] += d[c];
#pragma unroll
      for (int t = 0; t < T; ++t) {
        if (t < taps) {
          const int kx = t % g.KW, ky = (t / g.KW) % g.KH, ci = t / (g.KW * g.KH);
          const int iy = oy * g.stride + ky - g.pad, ix = ox * g.stride + kx - g.pad;
          const float xv = (iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi) ? s_x[(ci * g.Hi + iy) * g.Wi + ix] : 0.f;
#pragma unroll
          for (int c = 0; c < kCoT; ++c) acc[t][c] = fmaf(xv, d[c], acc[t][c]);
        }
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j <= T; ++j)
#pragma unroll
    for (int c = 0; c < kCoT; ++c) {
      const float v = warp_sum(acc[j][c]);
      if (lane == 0) s_red[warp][j * kCoT + c] = v;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < (T + 1) * kCoT; i += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += s_red[w][i];
    const int j = i / kCoT, c = i % kCoT;
    if (j < taps) atomicAdd(&dw[slot * sw + (size_t)(co0 + c) * taps + j], v);
    else if (j == T && db) atomicAdd(&db[slot * sb + co0 + c], v);
  }
}

// ------------------------------------------------------------------------------------------------ GroupNorm (NCHW fp32)
__global__ void __launch_bounds__(128)
groupnorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     long ss, int C, int HW, int groups, int act, const es_group* __restrict__ grp, int n_groups,
                     float* __restrict__ y, float* __restrict__ stats) {
  __shared__ float red[32];
  const int r = blockIdx.x / groups, gq = blockIdx.x % groups;
  const int gi = find_group(grp, n_groups, r);
  if (gi < 0) return;
  const int slot = grp[gi].slot, cpg = C / groups, n = cpg * HW;
  const float* xs = x + ((size_t)r * C + gq * cpg) * HW;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += xs[i];
  const float mean = block_sum(s, red) / n;
  float q = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { const float d = xs[i] - mean; q += d * d; }
  const float rstd = rsqrtf(block_sum(q, red) / n + kNormEps);
  if (threadIdx.x == 0) { stats[(size_t)blockIdx.x * 2] = mean; stats[(size_t)blockIdx.x * 2 + 1] = rstd; }
  float* ys = y + ((size_t)r * C + gq * cpg) * HW;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = gq * cpg + i / HW;
    ys[i] = act_fwd((xs[i] - mean) * rstd * gamma[slot * ss + c] + beta[slot * ss + c], act);
  }
}

__global__ void __launch_bounds__(128)
groupnorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ stats,
                     const float* __restrict__ gamma, const float* __restrict__ beta, long ss, int C, int HW, int groups,
                     int act, const es_group* __restrict__ grp, int n_groups, float* __restrict__ dx,
                     float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float red[32];
  const int r = blockIdx.x / groups, gq = blockIdx.x % groups;
  const int gi = find_group(grp, n_groups, r);
  if (gi < 0) return;
  const int slot = grp[gi].slot, cpg = C / groups, n = cpg * HW;
  const float mean = stats[(size_t)blockIdx.x * 2], rstd = stats[(size_t)blockIdx.x * 2 + 1];
  const size_t off = ((size_t)r * C + gq * cpg) * HW;
  float s1 = 0.f, s2 = 0.f;
  for (int cc = 0; cc < cpg; ++cc) {
    const int c = gq * cpg + cc;
    const float ga = gamma[slot * ss + c], be = beta[slot * ss + c];
    float a = 0.f, b = 0.f;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      const float xh = (x[off + cc * HW + i] - mean) * rstd;
      const float d = dy[off + cc * HW + i] * act_grad(xh * ga + be, act);
      a += d * xh;
      b += d;
    }
    a = block_sum(a, red);
    b = block_sum(b, red);
    if (threadIdx.x == 0 && dgamma) { atomicAdd(&dgamma[slot * ss + c], a); atomicAdd(&dbeta[slot * ss + c], b); }
    s1 += b * ga;   // every thread holds the block totals
    s2 += a * ga;
  }
  s1 /= n;
  s2 /= n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = gq * cpg + i / HW;
    const float ga = gamma[slot * ss + c], be = beta[slot * ss + c];
    const float xh = (x[off + i] - mean) * rstd;
    const float d = dy[off + i] * act_grad(xh * ga + be, act) * ga;
    dx[off + i] = rstd * (d - s1 - xh * s2);
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm (rows)
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, long ss,
                     int F, int act, const es_group* __restrict__ grp, int n_groups, int total_rows, float* __restrict__ y,
                     float* __restrict__ stats) {
  const int lane = threadIdx.x & 31, r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= total_rows) return;
  const int gi = find_group(grp, n_groups, r);
  if (gi < 0) return;
  const int slot = grp[gi].slot;
  const float* xs = x + (size_t)r * F;
  float s = 0.f;
  for (int i = lane; i < F; i += 32) s += xs[i];
  const float mean = warp_sum(s) / F;
  float q = 0.f;
  for (int i = lane; i < F; i += 32) { const float d = xs[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) / F + kNormEps);
  if (lane == 0) { stats[2 * r] = mean; stats[2 * r + 1] = rstd; }
  for (int i = lane; i < F; i += 32)
    y[(size_t)r * F + i] = act_fwd((xs[i] - mean) * rstd * gamma[slot * ss + i] + beta[slot * ss + i], act);
}

__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ stats,
                     const float* __restrict__ gamma, const float* __restrict__ beta, long ss, int F, int act,
                     const es_group* __restrict__ grp, int n_groups, int total_rows, float* __restrict__ dx,
                     float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int lane = threadIdx.x & 31, r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= total_rows) return;
  const int gi = find_group(grp, n_groups, r);
  if (gi < 0) return;
  const int slot = grp[gi].slot;
  const float mean = stats[2 * r], rstd = stats[2 * r + 1];
  float s1 = 0.f, s2 = 0.f;
  for (int i = lane; i < F; i += 32) {
    const float ga = gamma[slot * ss + i];
    const float xh = (x[(size_t)r * F + i] - mean) * rstd;
    const float d = dy[(size_t)r * F + i] * act_grad(xh * ga + beta[slot * ss + i], act);
    if (dgamma) { atomicAdd(&dgamma[slot * ss + i], d * xh); atomicAdd(&dbeta[slot * ss + i], d); }
    s1 += d * ga;
    s2 += d * ga * xh;
  }
  s1 = warp_sum(s1) / F;
  s2 = warp_sum(s2) / F;
  for (int i = lane; i < F; i += 32) {
    const float ga = gamma[slot * ss + i];
    const float xh = (x[(size_t)r * F + i] - mean) * rstd;
    const float d = dy[(size_t)r * F + i] * act_grad(xh * ga + beta[slot * ss + i], act) * ga;
    dx[(size_t)r * F + i] = rstd * (d - s1 - xh * s2);
  }
}

// ------------------------------------------------------------------------------------------------ max pooling
// grid = (rows*C planes, plane chunks): 32-bit index math only (a flat 64-bit index costs three 64-bit divisions per thread)
__global__ void maxpool_fwd_kernel(const float* __restrict__ x, int Hi, int Wi, int kh, int kw, int sh, int sw_,
                                   int Ho, int Wo, float* __restrict__ y, uint8_t* __restrict__ idx) {
  const int p = blockIdx.y * blockDim.x + threadIdx.x;
  if (p >= Ho * Wo) return;
  const int oy = p / Wo, ox = p - oy * Wo;
  const size_t rc = blockIdx.x;
  const float* xs = x + rc * Hi * Wi;
  float best = -INFINITY;
  int bi = 0;
  for (int ky = 0; ky < kh; ++ky)
    for (int kx = 0; kx < kw; ++kx) {
      const float v = xs[(oy * sh + ky) * Wi + ox * sw_ + kx];
      if (v > best) { best = v; bi = ky * kw + kx; }
    }
  y[rc * Ho * Wo + p] = best;
  idx[rc * Ho * Wo + p] = (uint8_t)bi;
}

// window geometry as template constants: the stride tests become shifts/masks (runtime `%` and `/` made this kernel
// integer-division-bound: 14 divisions per element, 0.7 ms for 50 M elements)
template <int KH, int KW, int SH, int SW>
__global__ void maxpool_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ idx, int Hi, int Wi, int Ho,
                                   int Wo, float* __restrict__ dx) {
  const int p = blockIdx.y * blockDim.x + threadIdx.x;
  if (p >= Hi * Wi) return;
  const int iy = p / Wi, ix = p - iy * Wi;
  const size_t rc = blockIdx.x;
  const float* dys = dy + rc * Ho * Wo;
  const uint8_t* ids = idx + rc * Ho * Wo;
  float acc = 0.f;
#pragma unroll
  for (int ky = 0; ky < KH; ++ky) {
    const int ty = iy - ky;
    if (ty < 0 || ty % SH != 0) continue;
    const int oy = ty / SH;
    if (oy >= Ho) continue;
#pragma unroll
    for (int kx = 0; kx < KW; ++kx) {
      const int tx = ix - kx;
      if (tx < 0 || tx % SW != 0) continue;
      const int ox = tx / SW;
      if (ox >= Wo) continue;
      const int o = oy * Wo + ox;
      if (ids[o] == ky * KW + kx) acc += dys[o];
    }
  }
  dx[rc * Hi * Wi + p] = acc;
}

// ------------------------------------------------------------------------------------------------ small SGEMMs (linear)
// C[m,n] = sum_k A(m,k) B(k,n) over one 64x64 tile; element access through functors (bounds -> 0).
template <class FA, class FB, class FC>
__device__ __forceinline__ void sgemm_tile(int M, int N, int K, int m0, int n0, FA fa, FB fb, FC fc) {
  __shared__ float sA[16][65], sB[16][65];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      const int kk = i % 16, mm = i / 16;
      sA[kk][mm] = (m0 + mm < M && k0 + kk < K) ? fa(m0 + mm, k0 + kk) : 0.f;
      sB[kk][mm] = (n0 + mm < N && k0 + kk < K) ? fb(k0 + kk, n0 + mm) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[kk][ty * 4 + i]; b[i] = sB[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (m0 + ty * 4 + i < M && n0 + tx * 4 + j < N) fc(m0 + ty * 4 + i, n0 + tx * 4 + j, acc[i][j]);
}

__global__ void __launch_bounds__(256)
linear_fwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w, const float* __restrict__ b, long sw,
                  long sb, int I, int O, const es_group* __restrict__ grp, int n_groups, float* __restrict__ y) {
  int gi, row0, ns;
  if (!chunk_of(grp, n_groups, 64, blockIdx.x, gi, row0, ns)) return;
  const int slot = grp[gi].slot;
  const float* W = w + slot * sw;
  sgemm_tile(ns, O, I, 0, blockIdx.y * 64,
             [&](int m, int k) { return x[(size_t)(row0 + m) * ldx + k]; },
             [&](int k, int n) { return W[(size_t)n * I + k]; },
             [&](int m, int n, float v) { y[(size_t)(row0 + m) * O + n] = v + (b ? b[slot * sb + n] : 0.f); });
}

// split-K variant for long reductions (discriminator fc1: I = 2313, O = 128, only 2 x 16 output tiles): grid.z splits
// the reduction, partial tiles meet in fp32 atomics on a zeroed y; split 0 adds the bias.
__global__ void __launch_bounds__(256)
linear_fwd_splitk_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w, const float* __restrict__ b,
                         long sw, long sb, int I, int O, int kper, const es_group* __restrict__ grp, int n_groups,
                         float* __restrict__ y) {
  int gi, row0, ns;
  if (!chunk_of(grp, n_groups, 64, blockIdx.x, gi, row0, ns)) return;
  const int slot = grp[gi].slot;
  const int kb = blockIdx.z * kper, kn = min(kper, I - kb);
  if (kn <= 0) return;
  const float* W = w + slot * sw + kb;
  const float* X = x + kb;
  const bool first = blockIdx.z == 0;
  sgemm_tile(ns, O, kn, 0, blockIdx.y * 64,
             [&](int m, int k) { return X[(size_t)(row0 + m) * ldx + k]; },
             [&](int k, int n) { return W[(size_t)n * I + k]; },
             [&](int m, int n, float v) { atomicAdd(&y[(size_t)(row0 + m) * O + n], v + ((first && b) ? b[slot * sb + n] : 0.f)); });
}

__global__ void __launch_bounds__(256)
linear_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ w, long sw, int I, int O,
                       const es_group* __restrict__ grp, int n_groups, float* __restrict__ dx, int lddx) {
  int gi, row0, ns;
  if (!chunk_of(grp, n_groups, 64, blockIdx.x, gi, row0, ns)) return;
  const float* W = w + grp[gi].slot * sw;
  sgemm_tile(ns, I, O, 0, blockIdx.y * 64,
             [&](int m, int k) { return dy[(size_t)(row0 + m) * O + k]; },
             [&](int k, int n) { return W[(size_t)k * I + n]; },
             [&](int m, int n, float v) { dx[(size_t)(row0 + m) * lddx + n] = v; });
}

// dW[o,i] += sum_r dy[r,o] x[r,i] over a chunk of 256 rows of one group; grid = (chunks, O tiles, I tiles)
__global__ void __launch_bounds__(256)
linear_bwd_weight_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ dy, int I, int O,
                         const es_group* __restrict__ grp, int n_groups, float* __restrict__ dw, float* __restrict__ db,
                         long sw, long sb) {
  int gi, row0, ns;
  if (!chunk_of(grp, n_groups, 256, blockIdx.x, gi, row0, ns)) return;
  const int slot = grp[gi].slot;
  float* DW = dw + slot * sw;
  sgemm_tile(O, I, ns, blockIdx.y * 64, blockIdx.z * 64,
             [&](int m, int k) { return dy[(size_t)(row0 + k) * O + m]; },
             [&](int k, int n) { return x[(size_t)(row0 + k) * ldx + n]; },
             [&](int m, int n, float v) { atomicAdd(&DW[(size_t)m * I + n], v); });
  if (db && blockIdx.z == 0) {
    const int o = blockIdx.y * 64 + threadIdx.x;
    if (threadIdx.x < 64 && o < O) {
      float s = 0.f;
      for (int k = 0; k < ns; ++k) s += dy[(size_t)(row0 + k) * O + o];
      atomicAdd(&db[slot * sb + o], s);
    }
  }
}

// ------------------------------------------------------------------------------------------------ spectral norm
__global__ void __launch_bounds__(256)
spectral_norm_fwd_kernel(const float* __restrict__ w_orig, float* __restrict__ u, float* __restrict__ v, long sw, long su,
                         long sv, int O, int I, int do_iter, const es_group* __restrict__ grp, float* __restrict__ w_sn,
                         long ssn, float* __restrict__ sigma_out, float* __restrict__ u_used, float* __restrict__ v_used) {
  extern __shared__ float sm[];   // v[I], u[O], wv[O]
  __shared__ float red[32];
  const int slot = blockIdx.x;
  if (grp && grp[slot].rows == 0) return;
  const float* W = w_orig + slot * sw;
  float* U = u + slot * su;
  float* V = v + slot * sv;
  float* s_v = sm;
  float* s_u = sm + I;
  float* s_wv = s_u + O;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = threadIdx.x; o < O; o += blockDim.x) s_u[o] = U[o];
  for (int i = threadIdx.x; i < I; i += blockDim.x) s_v[i] = V[i];
  __syncthreads();
  if (do_iter) {
    // v = normalize(W^T u)
    float nrm = 0.f;
    for (int i = threadIdx.x; i < I; i += blockDim.x) {
      float a = 0.f;
      for (int o = 0; o < O; ++o) a = fmaf(W[(size_t)o * I + i], s_u[o], a);
      s_v[i] = a;
      nrm += a * a;
    }
    nrm = sqrtf(block_sum(nrm, red));
    const float inv = 1.f / fmaxf(nrm, 1e-12f);
    for (int i = threadIdx.x; i < I; i += blockDim.x) { s_v[i] *= inv; V[i] = s_v[i]; }
    __syncthreads();
  }
  // wv = W v (one warp per row)
  for (int o = warp; o < O; o += 8) {
    float a = 0.f;
    for (int i = lane; i < I; i += 32) a = fmaf(W[(size_t)o * I + i], s_v[i], a);
    a = warp_sum(a);
    if (lane == 0) s_wv[o] = a;
  }
  __syncthreads();
  if (do_iter) {
    float nrm = 0.f;
    for (int o = threadIdx.x; o < O; o += blockDim.x) nrm += s_wv[o] * s_wv[o];
    nrm = sqrtf(block_sum(nrm, red));
    const float inv = 1.f / fmaxf(nrm, 1e-12f);
    for (int o = threadIdx.x; o < O; o += blockDim.x) { s_u[o] = s_wv[o] * inv; U[o] = s_u[o]; }
    __syncthreads();
  }
  float sg = 0.f;
  for (int o = threadIdx.x; o < O; o += blockDim.x) sg += s_u[o] * s_wv[o];
  sg = block_sum(sg, red);
  if (threadIdx.x == 0) sigma_out[slot] = sg;
  const float inv = 1.f / sg;
  float* WS = w_sn + slot * ssn;
  for (int i = threadIdx.x; i < O * I; i += blockDim.x) WS[i] = W[i] * inv;
  if (u_used) for (int o = threadIdx.x; o < O; o += blockDim.x) u_used[(size_t)slot * O + o] = s_u[o];
  if (v_used) for (int i = threadIdx.x; i < I; i += blockDim.x) v_used[(size_t)slot * I + i] = s_v[i];
}

// Multi-CTA power iteration for the large matrices (fc1.0: 128 x 2313 per expert): three launches, no inter-CTA waits.
//   phase 1: t = W^T u      (threads along i: coalesced rows of W)
//   phase 2: s = W t        (one warp per row)
//   phase 3: v = t/|t|, u = (s/|t|)/|s/|t||, sigma = u.(W v) = |s|/|t|, w_sn = W / sigma (elementwise, all CTAs)
// scratch per slot: t[I], s[O], n2[2].
__global__ void __launch_bounds__(256)
sn_phase1_kernel(const float* __restrict__ w_orig, const float* __restrict__ u, long sw, long su, int O, int I,
                 const es_group* __restrict__ grp, float* __restrict__ scratch) {
  __shared__ float s_u[512];
  const int slot = blockIdx.y;
  if (grp && grp[slot].rows == 0) return;
  const float* W = w_orig + slot * sw;
  float* T = scratch + (size_t)slot * (I + O + 2);
  for (int o = threadIdx.x; o < O; o += blockDim.x) s_u[o] = u[slot * su + o];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (i < I) {
    int o = 0;
    for (; o + 4 <= O; o += 4) {
      a0 = fmaf(W[(size_t)o * I + i], s_u[o], a0);
      a1 = fmaf(W[(size_t)(o + 1) * I + i], s_u[o + 1], a1);
      a2 = fmaf(W[(size_t)(o + 2) * I + i], s_u[o + 2], a2);
      a3 = fmaf(W[(size_t)(o + 3) * I + i], s_u[o + 3], a3);
    }
    for (; o < O; ++o) a0 = fmaf(W[(size_t)o * I + i], s_u[o], a0);
    a0 = (a0 + a1) + (a2 + a3);
    T[i] = a0;
  }
}

__global__ void __launch_bounds__(256)
sn_phase2_kernel(const float* __restrict__ w_orig, long sw, int O, int I, const es_group* __restrict__ grp,
                 float* __restrict__ scratch) {
  const int slot = blockIdx.y;
  if (grp && grp[slot].rows == 0) return;
  const float* W = w_orig + slot * sw;
  float* T = scratch + (size_t)slot * (I + O + 2);
  const int lane = threadIdx.x & 31, o = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (o >= O) return;
  float a = 0.f;
  for (int i = lane; i < I; i += 32) a = fmaf(W[(size_t)o * I + i], T[i], a);
  a = warp_sum(a);
  if (lane == 0) T[I + o] = a;
}

__global__ void __launch_bounds__(256)
sn_phase3_kernel(const float* __restrict__ w_orig, float* __restrict__ u, float* __restrict__ v, long sw, long su, long sv,
                 int O, int I, const es_group* __restrict__ grp, const float* __restrict__ scratch, float* __restrict__ w_sn,
                 long ssn, float* __restrict__ sigma_out, float* __restrict__ u_used, float* __restrict__ v_used) {
  const int slot = blockIdx.y;
  if (grp && grp[slot].rows == 0) return;
  const float* T = scratch + (size_t)slot * (I + O + 2);
  // |t|^2 and |s|^2 are summed HERE, by every CTA, in one fixed order (strided per thread, then the shuffle tree of block_sum):
  // bit-identical across CTAs, runs and data-parallel replicas.  (Cross-CTA atomics made u / v depend on the arrival
  // order, and u / v are STATE: replicas drifted apart in the last bit — found by the bit-identity check of dp_parity.)
  __shared__ float red[32];
  float pt = 0.f, ps = 0.f;
  for (int i = threadIdx.x; i < I; i += blockDim.x) pt = fmaf(T[i], T[i], pt);
  for (int o = threadIdx.x; o < O; o += blockDim.x) ps = fmaf(T[I + o], T[I + o], ps);
  const float nt2 = block_sum(pt, red);
  const float ns2 = block_sum(ps, red);
  const float nt = fmaxf(sqrtf(nt2), 1e-12f);               // |W^T u|
  const float nsp = sqrtf(ns2) / nt;                         // |W v|
  const float inv_u = 1.f / (fmaxf(nsp, 1e-12f) * nt);       // u = s * inv_u
  const float sg = nsp * nsp / fmaxf(nsp, 1e-12f);           // sigma = u . (W v)
  const float inv = 1.f / sg;
  const float* W = w_orig + slot * sw;
  float* WS = w_sn + slot * ssn;
  const int n = O * I;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) WS[i] = W[i] * inv;
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) sigma_out[slot] = sg;
    for (int i = threadIdx.x; i < I; i += blockDim.x) {
      const float x = T[i] / nt;
      v[slot * sv + i] = x;
      if (v_used) v_used[(size_t)slot * I + i] = x;
    }
    for (int o = threadIdx.x; o < O; o += blockDim.x) {
      const float x = T[I + o] * inv_u;
      u[slot * su + o] = x;
      if (u_used) u_used[(size_t)slot * O + o] = x;
    }
  }
}

// backward, multi-CTA: dot[slot] = <dw_sn, w_sn> by atomics, then the elementwise update
__global__ void __launch_bounds__(256)
sn_bwd_dot_kernel(const float* __restrict__ dw_sn, const float* __restrict__ w_sn, long ssn, int n,
                  const es_group* __restrict__ grp, float* __restrict__ dot) {
  __shared__ float red[32];
  const int slot = blockIdx.y;
  if (grp && grp[slot].rows == 0) return;
  const float* D = dw_sn + slot * ssn;
  const float* WS = w_sn + slot * ssn;
  float a = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) a = fmaf(D[i], WS[i], a);
  a = block_sum(a, red);
  if (threadIdx.x == 0) atomicAdd(&dot[slot], a);
}
__global__ void __launch_bounds__(256)
sn_bwd_apply_kernel(const float* __restrict__ dw_sn, const float* __restrict__ u_used, const float* __restrict__ v_used,
                    const float* __restrict__ sigma, const float* __restrict__ dot, long ssn, int O, int I,
                    float* __restrict__ dw_orig, long sw, const es_group* __restrict__ grp) {
  const int slot = blockIdx.y;
  if (grp && grp[slot].rows == 0) return;
  const float* D = dw_sn + slot * ssn;
  const float inv = 1.f / sigma[slot], dt = dot[slot];
  const int n = O * I;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int o = i / I, k = i - o * I;
    atomicAdd(&dw_orig[slot * sw + i], (D[i] - dt * u_used[(size_t)slot * O + o] * v_used[(size_t)slot * I + k]) * inv);
  }
}

__global__ void __launch_bounds__(256)
spectral_norm_bwd_kernel(const float* __restrict__ dw_sn, const float* __restrict__ w_sn, const float* __restrict__ u_used,
                         const float* __restrict__ v_used, const float* __restrict__ sigma, long ssn, int O, int I,
                         float* __restrict__ dw_orig, long sw, const es_group* __restrict__ grp) {
  __shared__ float red[32];
  const int slot = blockIdx.x;
  if (grp && grp[slot].rows == 0) return;   // skipped expert: its w_sn / sigma were never produced
  const float* D = dw_sn + slot * ssn;
  const float* WS = w_sn + slot * ssn;
  float dot = 0.f;
  for (int i = threadIdx.x; i < O * I; i += blockDim.x) dot += D[i] * WS[i];
  dot = block_sum(dot, red);
  const float inv = 1.f / sigma[slot];
  for (int i = threadIdx.x; i < O * I; i += blockDim.x) {
    const int o = i / I, k = i % I;
    // atomic: the two backward passes of a discriminator step (real / fake batch) may run concurrently on two streams
    atomicAdd(&dw_orig[slot * sw + i], (D[i] - dot * u_used[(size_t)slot * O + o] * v_used[(size_t)slot * I + k]) * inv);
  }
}

// ------------------------------------------------------------------------------------------------ elementwise
__global__ void add_relu_kernel(const float* a, const float* b, long n, float* y) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n) y[i] = fmaxf(a[i] + b[i], 0.f);
}
__global__ void relu_bwd_kernel(const float* dy, const float* y, long n, float* dx) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n) dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}
__global__ void gap_fwd_kernel(const float* x, int HW, long rc, float* y) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= rc) return;
  float s = 0.f;
  for (int p = 0; p < HW; ++p) s += x[i * HW + p];
  y[i] = s / HW;
}
__global__ void gap_bwd_kernel(const float* dy, int HW, long total, float* dx) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < total) dx[i] = dy[i / HW] / HW;
}
__global__ void dropout_kernel(const float* x, const float* m, float scale, long n, float* y) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n) y[i] = x[i] * m[i] * scale;
}
__global__ void axpy_kernel(float a, const float* x, long n, float* y) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n) y[i] += a * x[i];
}
__global__ void copy_cols_kernel(const float* src, int lds, int cols, long total, float* dst, int ldd, int col0) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long r = i / cols;
  const int c = i % cols;
  dst[r * ldd + col0 + c] = src[r * lds + c];
}

// ------------------------------------------------------------------------------------------------ Adam
__global__ void adam_bump_kernel(int32_t* step_count, const es_group* grp, int slots) {
  const int s = threadIdx.x;
  if (s < slots && (!grp || grp[s].rows > 0)) step_count[s] += 1;
}
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long n,
            long slot_stride, float lr, float b1, float b2, float eps, const int32_t* __restrict__ step_count,
            const es_group* __restrict__ grp) {
  const int slot = blockIdx.y;
  if (grp && grp[slot].rows == 0) return;
  const int t = step_count[slot];
  const double bc1 = 1.0 - pow((double)b1, (double)t);
  const float bc2s = (float)sqrt(1.0 - pow((double)b2, (double)t));
  const float step = (float)((double)lr / bc1);
  const long base = slot * slot_stride;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float gg = g[base + i];
    const float mm = m[base + i] + (gg - m[base + i]) * (1.f - b1);
    const float vv = v[base + i] * b2 + gg * gg * (1.f - b2);
    m[base + i] = mm;
    v[base + i] = vv;
    p[base + i] -= step * mm / (sqrtf(vv) / bc2s + eps);
  }
}

// float4 variant (n, slot_stride multiples of 4, 16-byte aligned bases — the arenas guarantee it): 7 x 16-byte accesses per
// thread and iteration instead of 7 x 4, which is what lets this 28 B/parameter stream approach the HBM rate
__global__ void __launch_bounds__(256)
adam_vec4_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v, long n4,
                 long slot_stride4, float lr, float b1, float b2, float eps, const int32_t* __restrict__ step_count,
                 const es_group* __restrict__ grp) {
  const int slot = blockIdx.y;
  if (grp && grp[slot].rows == 0) return;
  const int t = step_count[slot];
  const double bc1 = 1.0 - pow((double)b1, (double)t);
  const float bc2s = (float)sqrt(1.0 - pow((double)b2, (double)t));
  const float step = (float)((double)lr / bc1);
  const long base = slot * slot_stride4;
  const float c1 = 1.f - b1, c2 = 1.f - b2;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 gg = __ldg(g + base + i);
    float4 mm = m[base + i], vv = v[base + i], pp = p[base + i];
#define ES_ADAM1(c)                                         \
    mm.c = mm.c + (gg.c - mm.c) * c1;                       \
    vv.c = vv.c * b2 + gg.c * gg.c * c2;                    \
    pp.c -= step * mm.c / (sqrtf(vv.c) / bc2s + eps);
    ES_ADAM1(x) ES_ADAM1(y) ES_ADAM1(z) ES_ADAM1(w)
#undef ES_ADAM1
    m[base + i] = mm;
    v[base + i] = vv;
    p[base + i] = pp;
  }
}

static inline unsigned blocks_for(long n) { return (unsigned)ceil_div_l(n, 256); }

}  // namespace es

using namespace es;

static bool conv_ok(const es_conv2d* g) {
  return g && g->Ci > 0 && g->Co > 0 && g->stride > 0 && g->Ho == (g->Hi + 2 * g->pad - g->KH) / g->stride + 1 &&
         g->Wo == (g->Wi + 2 * g->pad - g->KW) / g->stride + 1;
}
static int pick_samples(int per_sample_floats, int pixels, int extra_floats) {
  int S = ceil_div(256, pixels);
  const int budget = (200 * 1024 - extra_floats * 4) / 4;
  if (S * per_sample_floats > budget) S = budget / per_sample_floats;
  return S < 1 ? 1 : S;
}

namespace es {
// direct data gradient for Ci == 1 (gradient w.r.t. the image); called by es_conv2d_bwd_data (conv_simt.cu)
int conv2d_bwd_data_ci1(const float* dy, const float* w, long slot_stride_w, const es_conv2d* g,
                                  const es_group* grp, int n_groups, int total_rows, float* dx, int accumulate,
                                  void* stream) {
  ES_REQUIRE(dy && w && grp && dx, "null pointer");
  ES_REQUIRE(conv_ok(g) && g->Ci == 1, "direct data-gradient path is for Ci == 1");
  ES_REQUIRE(n_groups >= 1 && n_groups <= kMaxGroups && total_rows > 0, "bad sizes");
  const int cit = g->Ci == 1 ? 1 : 8;
  const int out_sz = g->Co * g->Ho * g->Wo, wt = g->Co * g->KH * g->KW * cit;
  const int S = pick_samples(out_sz, g->Hi * g->Wi, wt);
  const size_t smem = ((size_t)S * out_sz + wt) * sizeof(float);
  if (smem > 220 * 1024) {      // unstaged variant
    ES_REQUIRE(wt * sizeof(float) <= 48 * 1024, "weight tile does not fit in shared memory");
    const dim3 grid1(total_rows + n_groups, g->Ci / cit);
    if (cit == 1) conv2d_bwd_data_kernel<1, false><<<grid1, 256, wt * sizeof(float), as_stream(stream)>>>(dy, w, slot_stride_w, *g, grp, n_groups, 1, dx, accumulate);
    else conv2d_bwd_data_kernel<8, false><<<grid1, 256, wt * sizeof(float), as_stream(stream)>>>(dy, w, slot_stride_w, *g, grp, n_groups, 1, dx, accumulate);
    ES_LAUNCH_CHECK();
    return ES_OK;
  }
  const dim3 grid(ceil_div(total_rows, S) + n_groups, g->Ci / cit);
  if (cit == 1) {
    ES_CUDA(cudaFuncSetAttribute(conv2d_bwd_data_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    conv2d_bwd_data_kernel<1><<<grid, 256, smem, as_stream(stream)>>>(dy, w, slot_stride_w, *g, grp, n_groups, S, dx, accumulate);
  } else {
    ES_CUDA(cudaFuncSetAttribute(conv2d_bwd_data_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    conv2d_bwd_data_kernel<8><<<grid, 256, smem, as_stream(stream)>>>(dy, w, slot_stride_w, *g, grp, n_groups, S, dx, accumulate);
  }
  ES_LAUNCH_CHECK();
  return ES_OK;
}
}  // namespace es

// few-tap weight gradient (Ci*KH*KW <= 9); called by es_conv2d_bwd_weight (conv_simt.cu)
namespace es {
int conv2d_bwd_weight_fewtaps(const float* x, const float* dy, const es_conv2d* g, const es_group* grp, int n_groups,
                              int total_rows, float* dw, float* db, long slot_stride_w, long slot_stride_b, void* stream) {
  ES_REQUIRE(g->Co % kCoT == 0 && g->Ci * g->KH * g->KW <= 9, "few-tap path needs Co % 8 == 0 and <= 9 taps");
  const size_t smem = ((size_t)g->Ci * g->Hi * g->Wi + (size_t)g->Ho * g->Wo * kCoT + 4) * sizeof(float);
  ES_REQUIRE(smem <= 220 * 1024, "sample does not fit in shared memory");
  int per = ceil_div(total_rows * (g->Co / kCoT), 4 * 148);
  if (per < 1) per = 1;
  ES_CUDA(cudaFuncSetAttribute(conv2d_bwd_weight_fewtaps_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  conv2d_bwd_weight_fewtaps_kernel<9><<<dim3(ceil_div(total_rows, per) + n_groups, g->Co / kCoT), 256, smem, as_stream(stream)>>>(
      x, dy, *g, grp, n_groups, per, dw, db, slot_stride_w, slot_stride_b);
  ES_LAUNCH_CHECK();
  return ES_OK;
}
}  // namespace es

extern "C" int es_groupnorm_fwd(const float* x, const float* gamma, const float* beta, long slot_stride, int C, int HW,
                                int groups, int act, const es_group* grp, int n_groups, int total_rows, float* y,
                                float* stats, void* stream) {
  ES_REQUIRE(x && gamma && beta && grp && y && stats, "null pointer");
  ES_REQUIRE(C > 0 && groups > 0 && C % groups == 0 && HW > 0 && total_rows > 0 && n_groups <= kMaxGroups, "bad sizes");
  groupnorm_fwd_kernel<<<total_rows * groups, 128, 0, as_stream(stream)>>>(x, gamma, beta, slot_stride, C, HW, groups, act,
                                                                          grp, n_groups, y, stats);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_groupnorm_bwd(const float* dy, const float* x, const float* stats, const float* gamma,
                                const float* beta, long slot_stride, int C, int HW, int groups, int act,
                                const es_group* grp, int n_groups, int total_rows, float* dx, float* dgamma,
                                float* dbeta, void* stream) {
  ES_REQUIRE(dy && x && stats && gamma && beta && grp && dx, "null pointer");
  ES_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "dgamma and dbeta go together");
  ES_REQUIRE(C > 0 && groups > 0 && C % groups == 0 && HW > 0 && total_rows > 0 && n_groups <= kMaxGroups, "bad sizes");
  groupnorm_bwd_kernel<<<total_rows * groups, 128, 0, as_stream(stream)>>>(dy, x, stats, gamma, beta, slot_stride, C, HW,
                                                                          groups, act, grp, n_groups, dx, dgamma, dbeta);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_layernorm_fwd(const float* x, const float* gamma, const float* beta, long slot_stride, int F, int act,
                                const es_group* grp, int n_groups, int total_rows, float* y, float* stats,
                                void* stream) {
  ES_REQUIRE(x && gamma && beta && grp && y && stats && F > 0 && total_rows > 0 && n_groups <= kMaxGroups, "bad arguments");
  layernorm_fwd_kernel<<<ceil_div(total_rows, 8), 256, 0, as_stream(stream)>>>(x, gamma, beta, slot_stride, F, act, grp,
                                                                              n_groups, total_rows, y, stats);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_layernorm_bwd(const float* dy, const float* x, const float* stats, const float* gamma,
                                const float* beta, long slot_stride, int F, int act, const es_group* grp, int n_groups,
                                int total_rows, float* dx, float* dgamma, float* dbeta, void* stream) {
  ES_REQUIRE(dy && x && stats && gamma && beta && grp && dx && F > 0 && total_rows > 0 && n_groups <= kMaxGroups, "bad arguments");
  ES_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "dgamma and dbeta go together");
  layernorm_bwd_kernel<<<ceil_div(total_rows, 8), 256, 0, as_stream(stream)>>>(dy, x, stats, gamma, beta, slot_stride, F,
                                                                              act, grp, n_groups, total_rows, dx, dgamma,
                                                                              dbeta);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_maxpool_fwd(const float* x, int C, int Hi, int Wi, int kh, int kw, int sh, int sw, int total_rows,
                              float* y, uint8_t* idx, void* stream) {
  ES_REQUIRE(x && y && idx && C > 0 && kh > 0 && kw > 0 && sh > 0 && sw > 0 && total_rows > 0, "bad arguments");
  const int Ho = (Hi - kh) / sh + 1, Wo = (Wi - kw) / sw + 1;
  ES_REQUIRE((long)total_rows * C < 2147483647L, "too many planes");
  maxpool_fwd_kernel<<<dim3(total_rows * C, ceil_div(Ho * Wo, 128)), 128, 0, as_stream(stream)>>>(x, Hi, Wi, kh, kw, sh, sw, Ho, Wo, y, idx);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_maxpool_bwd(const float* dy, const uint8_t* idx, int C, int Hi, int Wi, int kh, int kw, int sh, int sw,
                              int total_rows, float* dx, void* stream) {
  ES_REQUIRE(dy && dx && idx && C > 0 && total_rows > 0, "bad arguments");
  const int Ho = (Hi - kh) / sh + 1, Wo = (Wi - kw) / sw + 1;
  ES_REQUIRE((long)total_rows * C < 2147483647L, "too many planes");
  const dim3 grid(total_rows * C, ceil_div(Hi * Wi, 128));
  cudaStream_t st = as_stream(stream);
  if (kh == 2 && kw == 2 && sh == 2 && sw == 2) maxpool_bwd_kernel<2, 2, 2, 2><<<grid, 128, 0, st>>>(dy, idx, Hi, Wi, Ho, Wo, dx);
  else if (kh == 2 && kw == 1 && sh == 2 && sw == 1) maxpool_bwd_kernel<2, 1, 2, 1><<<grid, 128, 0, st>>>(dy, idx, Hi, Wi, Ho, Wo, dx);
  else if (kh == 2 && kw == 2 && sh == 1 && sw == 1) maxpool_bwd_kernel<2, 2, 1, 1><<<grid, 128, 0, st>>>(dy, idx, Hi, Wi, Ho, Wo, dx);
  else { set_error("es_maxpool_bwd: supported windows are (2,2)/s(2,2), (2,1)/s(2,1), (2,2)/s(1,1)"); return ES_ERR_UNSUPPORTED; }
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_linear_fwd(const float* x, int ldx, const float* w, const float* b, long slot_stride_w,
                             long slot_stride_b, int I, int O, const es_group* grp, int n_groups, int total_rows, float* y,
                             void* stream) {
  ES_REQUIRE(x && w && grp && y && I > 0 && O > 0 && ldx >= I && total_rows > 0 && n_groups <= kMaxGroups, "bad arguments");
  const int chunks = ceil_div(total_rows, 64) + n_groups, ntiles = ceil_div(O, 64);
  if (I >= 512 && chunks * ntiles < 2 * 148) {
    int splits = min(ceil_div(4 * 148, chunks * ntiles), ceil_div(I, 128));
    const int kper = ceil_div(ceil_div(I, splits), 16) * 16;
    splits = ceil_div(I, kper);
    ES_CUDA(cudaMemsetAsync(y, 0, (size_t)total_rows * O * sizeof(float), as_stream(stream)));
    linear_fwd_splitk_kernel<<<dim3(chunks, ntiles, splits), 256, 0, as_stream(stream)>>>(
        x, ldx, w, b, slot_stride_w, slot_stride_b, I, O, kper, grp, n_groups, y);
    ES_LAUNCH_CHECK();
    return ES_OK;
  }
  linear_fwd_kernel<<<dim3(chunks, ntiles), 256, 0, as_stream(stream)>>>(
      x, ldx, w, b, slot_stride_w, slot_stride_b, I, O, grp, n_groups, y);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_linear_bwd_data(const float* dy, const float* w, long slot_stride_w, int I, int O, const es_group* grp,
                                  int n_groups, int total_rows, float* dx, int lddx, void* stream) {
  ES_REQUIRE(dy && w && grp && dx && I > 0 && O > 0 && lddx >= I && total_rows > 0 && n_groups <= kMaxGroups, "bad arguments");
  linear_bwd_data_kernel<<<dim3(ceil_div(total_rows, 64) + n_groups, ceil_div(I, 64)), 256, 0, as_stream(stream)>>>(
      dy, w, slot_stride_w, I, O, grp, n_groups, dx, lddx);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_linear_bwd_weight(const float* x, int ldx, const float* dy, int I, int O, const es_group* grp,
                                    int n_groups, int total_rows, float* dw, float* db, long slot_stride_w,
                                    long slot_stride_b, void* stream) {
  ES_REQUIRE(x && dy && grp && dw && I > 0 && O > 0 && total_rows > 0 && n_groups <= kMaxGroups, "bad arguments");
  linear_bwd_weight_kernel<<<dim3(ceil_div(total_rows, 256) + n_groups, ceil_div(O, 64), ceil_div(I, 64)), 256, 0,
                             as_stream(stream)>>>(x, ldx, dy, I, O, grp, n_groups, dw, db, slot_stride_w, slot_stride_b);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_spectral_norm_fwd(const float* w_orig, float* u, float* v, long slot_stride_w, long slot_stride_u,
                                    long slot_stride_v, int slots, int O, int I, int do_power_iter, const es_group* grp,
                                    float* w_sn, long slot_stride_sn, float* sigma_out, float* u_used, float* v_used,
                                    float* scratch, void* stream) {
  ES_REQUIRE(w_orig && u && v && w_sn && sigma_out && slots >= 1 && O > 0 && I > 0, "bad arguments");
  if (scratch && do_power_iter && O <= 512 && (long)O * I >= 16384) {
    cudaStream_t st = as_stream(stream);
    const size_t per = (size_t)I + O + 2;
    (void)per;      // every scratch element that is read is written first (t by phase 1, s by phase 2): no memset
    sn_phase1_kernel<<<dim3(ceil_div(I, 256), slots), 256, 0, st>>>(w_orig, u, slot_stride_w, slot_stride_u, O, I, grp, scratch);
    sn_phase2_kernel<<<dim3(ceil_div(O, 8), slots), 256, 0, st>>>(w_orig, slot_stride_w, O, I, grp, scratch);
    const int nb = min(64, ceil_div(O * I, 2048));
    sn_phase3_kernel<<<dim3(nb, slots), 256, 0, st>>>(w_orig, u, v, slot_stride_w, slot_stride_u, slot_stride_v, O, I, grp,
                                                       scratch, w_sn, slot_stride_sn, sigma_out, u_used, v_used);
    ES_LAUNCH_CHECK();
    return ES_OK;
  }
  const size_t smem = ((size_t)I + 2 * O) * sizeof(float);
  ES_REQUIRE(smem <= 48 * 1024, "spectral norm vectors do not fit in shared memory");
  spectral_norm_fwd_kernel<<<slots, 256, smem, as_stream(stream)>>>(w_orig, u, v, slot_stride_w, slot_stride_u,
                                                                   slot_stride_v, O, I, do_power_iter, grp, w_sn,
                                                                   slot_stride_sn, sigma_out, u_used, v_used);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_spectral_norm_bwd(const float* dw_sn, const float* w_sn, const float* u_used, const float* v_used,
                                    const float* sigma, long slot_stride_sn, int slots, int O, int I, float* dw_orig,
                                    long slot_stride_w, const es_group* grp, float* scratch, void* stream) {
  ES_REQUIRE(dw_sn && w_sn && u_used && v_used && sigma && dw_orig && slots >= 1, "bad arguments");
  if (scratch && (long)O * I >= 16384) {
    cudaStream_t st = as_stream(stream);
    ES_CUDA(cudaMemsetAsync(scratch, 0, slots * sizeof(float), st));
    const int nb = min(64, ceil_div(O * I, 2048));
    sn_bwd_dot_kernel<<<dim3(nb, slots), 256, 0, st>>>(dw_sn, w_sn, slot_stride_sn, O * I, grp, scratch);
    sn_bwd_apply_kernel<<<dim3(nb, slots), 256, 0, st>>>(dw_sn, u_used, v_used, sigma, scratch, slot_stride_sn, O, I, dw_orig,
                                                          slot_stride_w, grp);
    ES_LAUNCH_CHECK();
    return ES_OK;
  }
  spectral_norm_bwd_kernel<<<slots, 256, 0, as_stream(stream)>>>(dw_sn, w_sn, u_used, v_used, sigma, slot_stride_sn, O, I,
                                                                dw_orig, slot_stride_w, grp);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_add_relu_fwd(const float* a, const float* b, long n, float* y, void* stream) {
  ES_REQUIRE(a && b && y && n > 0, "bad arguments");
  add_relu_kernel<<<blocks_for(n), 256, 0, as_stream(stream)>>>(a, b, n, y);
  ES_LAUNCH_CHECK();
  return ES_OK;
}
extern "C" int es_relu_bwd(const float* dy, const float* y, long n, float* dx, void* stream) {
  ES_REQUIRE(dy && y && dx && n > 0, "bad arguments");
  relu_bwd_kernel<<<blocks_for(n), 256, 0, as_stream(stream)>>>(dy, y, n, dx);
  ES_LAUNCH_CHECK();
  return ES_OK;
}
extern "C" int es_gap_fwd(const float* x, int C, int HW, int total_rows, float* y, void* stream) {
  ES_REQUIRE(x && y && C > 0 && HW > 0 && total_rows > 0, "bad arguments");
  gap_fwd_kernel<<<blocks_for((long)total_rows * C), 256, 0, as_stream(stream)>>>(x, HW, (long)total_rows * C, y);
  ES_LAUNCH_CHECK();
  return ES_OK;
}
extern "C" int es_gap_bwd(const float* dy, int C, int HW, int total_rows, float* dx, void* stream) {
  ES_REQUIRE(dy && dx && C > 0 && HW > 0 && total_rows > 0, "bad arguments");
  const long total = (long)total_rows * C * HW;
  gap_bwd_kernel<<<blocks_for(total), 256, 0, as_stream(stream)>>>(dy, HW, total, dx);
  ES_LAUNCH_CHECK();
  return ES_OK;
}
extern "C" int es_dropout(const float* x, const float* keep_mask, float p, long n, float* y, void* stream) {
  ES_REQUIRE(x && keep_mask && y && n > 0 && p >= 0.f && p < 1.f, "bad arguments");
  dropout_kernel<<<blocks_for(n), 256, 0, as_stream(stream)>>>(x, keep_mask, 1.f / (1.f - p), n, y);
  ES_LAUNCH_CHECK();
  return ES_OK;
}
extern "C" int es_axpy(float alpha, const float* x, long n, float* y, void* stream) {
  ES_REQUIRE(x && y && n > 0, "bad arguments");
  axpy_kernel<<<blocks_for(n), 256, 0, as_stream(stream)>>>(alpha, x, n, y);
  ES_LAUNCH_CHECK();
  return ES_OK;
}
extern "C" int es_copy_cols(const float* src, int lds, int cols, int rows, float* dst, int ldd, int col0, void* stream) {
  ES_REQUIRE(src && dst && cols > 0 && rows > 0 && lds >= cols && ldd >= col0 + cols, "bad arguments");
  const long total = (long)rows * cols;
  copy_cols_kernel<<<blocks_for(total), 256, 0, as_stream(stream)>>>(src, lds, cols, total, dst, ldd, col0);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_adam_step(float* p, const float* g, float* m, float* v, long n, long slot_stride, int slots, float lr,
                            float beta1, float beta2, float eps, int32_t* step_count, const es_group* grp, void* stream) {
  ES_REQUIRE(p && g && m && v && step_count && n > 0 && slots >= 1 && slots <= 1024, "bad arguments");
  adam_bump_kernel<<<1, 1024, 0, as_stream(stream)>>>(step_count, grp, slots);
  ES_LAUNCH_CHECK();
  if (n % 4 == 0 && slot_stride % 4 == 0 && (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0) {
    const unsigned bx4 = (unsigned)min(ceil_div_l(n / 4, 256 * 2), 148L * 8);
    adam_vec4_kernel<<<dim3(bx4 < 1 ? 1 : bx4, slots), 256, 0, as_stream(stream)>>>(
        (float4*)p, (const float4*)g, (float4*)m, (float4*)v, n / 4, slot_stride / 4, lr, beta1, beta2, eps, step_count, grp);
    ES_LAUNCH_CHECK();
    return ES_OK;
  }
  const unsigned bx = (unsigned)min(ceil_div_l(n, 256 * 4), 148L * 8);
  adam_kernel<<<dim3(bx < 1 ? 1 : bx, slots), 256, 0, as_stream(stream)>>>(p, g, m, v, n, slot_stride, lr, beta1, beta2, eps,
                                                                         step_count, grp);
  ES_LAUNCH_CHECK();
  return ES_OK;
}
