// K2 companions — everything in the generator that is not a tensor-core GEMM: the 19->256 head, the fused
// LayerNorm/GroupNorm + LeakyReLU passes (bf16 NHWC activations, fp32 statistics) with their backward, the final
// 64->1 convolution + ReLU, weight (re)packing, and CUDA-core (SIMT) reference versions of the implicit GEMMs.
// Reference: Generator.forward (expertsim/models/proton/generator.py:13-52).
#include "common.cuh"
#include "gen_common.cuh"

namespace es {

// ----------------------------------------------------------------------------------------------- fc1 (19 -> 256)
constexpr int kF1 = 256, kIn = 19, kNz = 10, kNc = 9;

__global__ void __launch_bounds__(256)
gen_fc1_fwd_kernel(const float* __restrict__ z1, const float* __restrict__ z2, const float* __restrict__ cond,
                   const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ gamma,
                   const float* __restrict__ beta, long sw, long sv, const es_group* __restrict__ grp, int E,
                   int total_rows, int two_pass, float* __restrict__ x0, float* __restrict__ lin,
                   __nv_bfloat16* __restrict__ h) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = blockIdx.x * 8 + warp;
  if (r >= total_rows) return;
  const RowMap m = map_row(grp, E, r, two_pass);
  if (m.g < 0) return;
  const int slot = grp[m.g].slot;
  const float* zz = m.pass ? z2 : z1;
  float xin = 0.f;
  if (lane < kNz) xin = zz[(size_t)m.j * kNz + lane];
  else if (lane < kIn) xin = cond[(size_t)m.j * kNc + lane - kNz];
  if (lane < kIn) x0[(size_t)r * kIn + lane] = xin;
  const float* W = w + slot * sw;
  float v[8];
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int o = lane + 32 * q;
    float acc = b[slot * sv + o];
    for (int k = 0; k < kIn; ++k) acc = fmaf(W[o * kIn + k], __shfl_sync(0xffffffffu, xin, k), acc);
    v[q] = acc;
    s += acc;
    if (lin) lin[(size_t)r * kF1 + o] = acc;
  }
  if (!gamma) {     // linear head only (neutron: BatchNorm follows as its own pass)
#pragma unroll
    for (int q = 0; q < 8; ++q) h[(size_t)r * kF1 + lane + 32 * q] = f2bf(v[q]);
    return;
  }
  const float mean = warp_sum(s) * (1.f / kF1);
  float ss = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) ss += (v[q] - mean) * (v[q] - mean);
  const float rstd = rsqrtf(warp_sum(ss) * (1.f / kF1) + kNormEps);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int o = lane + 32 * q;
    const float y = (v[q] - mean) * rstd * gamma[slot * sv + o] + beta[slot * sv + o];
    h[(size_t)r * kF1 + o] = f2bf(lrelu(y));
  }
}

// one CTA = (group, chunk): rows of the group strided by chunks; gradients staged in shared memory
__global__ void __launch_bounds__(256)
gen_fc1_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ x0, const float* __restrict__ lin,
                   const float* __restrict__ gamma, const float* __restrict__ beta, long sw, long sv,
                   const es_group* __restrict__ grp, int chunks, float* __restrict__ dw, float* __restrict__ db,
                   float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float s_dw[kF1 * kIn];
  __shared__ float s_v[3][kF1];
  const int g = blockIdx.x / chunks, ch = blockIdx.x % chunks;
  const es_group G = grp[g];
  if (G.rows == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kF1 * kIn; i += 256) s_dw[i] = 0.f;
  for (int i = threadIdx.x; i < 3 * kF1; i += 256) (&s_v[0][0])[i] = 0.f;
  __syncthreads();
  const int slot = G.slot;
  for (int i = ch * 8 + warp; i < G.rows; i += chunks * 8) {
    const int r = G.row_start + i;
    if (!gamma) {   // linear head only: dh is the gradient w.r.t. the linear output
      const float xin0 = lane < kIn ? x0[(size_t)r * kIn + lane] : 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int o = lane + 32 * q;
        const float dl = dh[(size_t)r * kF1 + o];
        atomicAdd(&s_v[0][o], dl);
        for (int k = 0; k < kIn; ++k) atomicAdd(&s_dw[o * kIn + k], dl * __shfl_sync(0xffffffffu, xin0, k));
      }
      continue;
    }
    float v[8], xh[8], dyn[8];
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) { v[q] = lin[(size_t)r * kF1 + lane + 32 * q]; s += v[q]; }
    const float mean = warp_sum(s) * (1.f / kF1);
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) ss += (v[q] - mean) * (v[q] - mean);
    const float rstd = rsqrtf(warp_sum(ss) * (1.f / kF1) + kNormEps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int o = lane + 32 * q;
      xh[q] = (v[q] - mean) * rstd;
      const float ga = gamma[slot * sv + o];
      const float y = xh[q] * ga + beta[slot * sv + o];
      const float d = dh[(size_t)r * kF1 + o] * (y > 0.f ? 1.f : kLReLU);
      atomicAdd(&s_v[1][o], d * xh[q]);   // dgamma
      atomicAdd(&s_v[2][o], d);           // dbeta
      dyn[q] = d * ga;
      s1 += dyn[q];
      s2 += dyn[q] * xh[q];
    }
    s1 = warp_sum(s1) * (1.f / kF1);
    s2 = warp_sum(s2) * (1.f / kF1);
    const float xin = lane < kIn ? x0[(size_t)r * kIn + lane] : 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int o = lane + 32 * q;
      const float dl = rstd * (dyn[q] - s1 - xh[q] * s2);
      atomicAdd(&s_v[0][o], dl);          // dbias
      for (int k = 0; k < kIn; ++k) atomicAdd(&s_dw[o * kIn + k], dl * __shfl_sync(0xffffffffu, xin, k));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kF1 * kIn; i += 256) atomicAdd(&dw[slot * sw + i], s_dw[i]);
  for (int i = threadIdx.x; i < kF1; i += 256) {
    atomicAdd(&db[slot * sv + i], s_v[0][i]);
    if (gamma) {
      atomicAdd(&dgamma[slot * sv + i], s_v[1][i]);
      atomicAdd(&dbeta[slot * sv + i], s_v[2][i]);
    }
  }
}

// ----------------------------------------------------------------------------------------------- big LayerNorm
// one CTA per row, three sweeps (mean, centred second moment, normalise); sweeps 2-3 hit L2.
__global__ void __launch_bounds__(512)
ln_lrelu_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                    long slot_stride, int F, const es_group* __restrict__ grp, int n_groups,
                    __nv_bfloat16* __restrict__ y, float* __restrict__ stats) {
  __shared__ float red[32];
  const int r = blockIdx.x;
  const int g = find_group(grp, n_groups, r);
  if (g < 0) return;
  const int slot = grp[g].slot;
  const uint4* x4 = reinterpret_cast<const uint4*>(x + (size_t)r * F);
  const int n4 = F / 8;
  float f[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    unpack8(__ldg(x4 + i), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) s += f[k];
  }
  const float mean = block_sum(s, red) / (float)F;
  float ss = 0.f;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    unpack8(__ldg(x4 + i), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) ss += (f[k] - mean) * (f[k] - mean);
  }
  const float rstd = rsqrtf(block_sum(ss, red) / (float)F + kNormEps);
  if (threadIdx.x == 0) { stats[2 * r] = mean; stats[2 * r + 1] = rstd; }
  const float4* g4 = reinterpret_cast<const float4*>(gamma + slot * slot_stride);
  const float4* b4 = reinterpret_cast<const float4*>(beta + slot * slot_stride);
  uint4* y4 = reinterpret_cast<uint4*>(y + (size_t)r * F);
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    unpack8(__ldg(x4 + i), f);
    const float4 ga = __ldg(g4 + 2 * i), gb = __ldg(g4 + 2 * i + 1), ba = __ldg(b4 + 2 * i), bb = __ldg(b4 + 2 * i + 1);
    const float gg[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
    const float be[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = lrelu((f[k] - mean) * rstd * gg[k] + be[k]);
    y4[i] = pack8(f);
  }
}

template <bool FAN>
__global__ void __launch_bounds__(512)
ln_lrelu_bwd_kernel(const __nv_bfloat16* __restrict__ dy_up, int Hs, int Ws, int Hu, int Wu, int C,
                    const __nv_bfloat16* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                    const float* __restrict__ beta, long slot_stride, const es_group* __restrict__ grp, int n_groups,
                    __nv_bfloat16* __restrict__ dx) {
  __shared__ float red[32];
  __shared__ int ylo[64], yhi[64], xlo[64], xhi[64];
  const int r = blockIdx.x;
  const int g = find_group(grp, n_groups, r);
  if (g < 0) {
    // rows of skipped experts: fc2's weight-gradient kernel streams this tensor through TMA boxes that may straddle a
    // group's end, where the rows meet a zero-padded operand — they must be finite (0 * NaN would poison the sum)
    uint4* o4 = reinterpret_cast<uint4*>(dx + (size_t)r * Hs * Ws * C);
    for (int i = threadIdx.x; i < Hs * Ws * C / 8; i += blockDim.x) o4[i] = make_uint4(0, 0, 0, 0);
    return;
  }
  const int slot = grp[g].slot;
  if (FAN) {
    if (threadIdx.x == 0) { build_fanin(Hs, Hu, ylo, yhi); build_fanin(Ws, Wu, xlo, xhi); }
    __syncthreads();
  }
  const int F = Hs * Ws * C, n4 = F / 8, c4 = C / 8;
  const float mean = stats[2 * r], rstd = stats[2 * r + 1];
  const __nv_bfloat16* dyr = dy_up + (size_t)r * Hu * Wu * C;
  const uint4* d4 = reinterpret_cast<const uint4*>(dyr);
  const uint4* x4 = reinterpret_cast<const uint4*>(x + (size_t)r * F);
  const float4* ga4 = reinterpret_cast<const float4*>(gamma + slot * slot_stride);
  const float4* be4 = reinterpret_cast<const float4*>(beta + slot * slot_stride);
  float s1 = 0.f, s2 = 0.f;
  // U elements of 8 features are fetched before any is consumed (one 16-byte load in flight per thread left this kernel
  // latency-bound); with the x2 upsample folded into conv1's data gradient the fan-in path is only used by tests
  constexpr int U = FAN ? 1 : 4;
  uint4 qx[U], qd[U];
  float4 qg[U][2], qb[U][2];
  float da[8], xv[8];
  auto fetch = [&](int i0) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < n4) {
        qx[u] = __ldg(x4 + i);
        if (!FAN) qd[u] = __ldg(d4 + i);
        qg[u][0] = __ldg(ga4 + 2 * i); qg[u][1] = __ldg(ga4 + 2 * i + 1);
        qb[u][0] = __ldg(be4 + 2 * i); qb[u][1] = __ldg(be4 + 2 * i + 1);
      }
    }
  };
  auto grad_of = [&](int u, int i) {
    if (FAN) {
      const int pix = i / c4, c8 = (i % c4) * 8;
      load_da8(dyr, Wu, C, c8, ylo, yhi, xlo, xhi, pix / Ws, pix % Ws, da);
    } else {
      unpack8(qd[u], da);
    }
  };
  for (int i0 = threadIdx.x; i0 < n4; i0 += U * blockDim.x) {
    fetch(i0);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i >= n4) continue;
      grad_of(u, i);
      unpack8(qx[u], xv);
      const float gk8[8] = {qg[u][0].x, qg[u][0].y, qg[u][0].z, qg[u][0].w, qg[u][1].x, qg[u][1].y, qg[u][1].z, qg[u][1].w};
      const float bk8[8] = {qb[u][0].x, qb[u][0].y, qb[u][0].z, qb[u][0].w, qb[u][1].x, qb[u][1].y, qb[u][1].z, qb[u][1].w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xh = (xv[k] - mean) * rstd;
        const float yv = xh * gk8[k] + bk8[k];
        const float d = da[k] * (yv > 0.f ? 1.f : kLReLU) * gk8[k];
        s1 += d;
        s2 += d * xh;
      }
    }
  }
  s1 = block_sum(s1, red) / (float)F;
  s2 = block_sum(s2, red) / (float)F;
  uint4* dx4 = reinterpret_cast<uint4*>(dx + (size_t)r * F);
  for (int i0 = threadIdx.x; i0 < n4; i0 += U * blockDim.x) {
    fetch(i0);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i >= n4) continue;
      grad_of(u, i);
      unpack8(qx[u], xv);
      const float gk8[8] = {qg[u][0].x, qg[u][0].y, qg[u][0].z, qg[u][0].w, qg[u][1].x, qg[u][1].y, qg[u][1].z, qg[u][1].w};
      const float bk8[8] = {qb[u][0].x, qb[u][0].y, qb[u][0].z, qb[u][0].w, qb[u][1].x, qb[u][1].y, qb[u][1].z, qb[u][1].w};
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xh = (xv[k] - mean) * rstd;
        const float yv = xh * gk8[k] + bk8[k];
        const float d = da[k] * (yv > 0.f ? 1.f : kLReLU) * gk8[k];
        o[k] = rstd * (d - s1 - xh * s2);
      }
      dx4[i] = pack8(o);
    }
  }
}

// column reductions for the big LayerNorm: each thread owns 8 consecutive features and walks the rows of one group
template <bool FAN>
__global__ void __launch_bounds__(128)
ln_affine_bwd_kernel(const __nv_bfloat16* __restrict__ dy_up, int Hs, int Ws, int Hu, int Wu, int C,
                     const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dx,
                     const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                     long slot_stride, const es_group* __restrict__ grp, const int* __restrict__ row_map,
                     long out_stride, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias) {
  __shared__ int ylo[64], yhi[64], xlo[64], xhi[64];
  const es_group G = grp[blockIdx.y];
  if (G.rows == 0) return;
  if (FAN) {
    if (threadIdx.x == 0) { build_fanin(Hs, Hu, ylo, yhi); build_fanin(Ws, Wu, xlo, xhi); }
    __syncthreads();
  }
  const int F = Hs * Ws * C, n4 = F / 8, c4 = C / 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int pix = i / c4, c8 = (i % c4) * 8;
  const float* ga = gamma + G.slot * slot_stride + (size_t)i * 8;
  const float* be = beta + G.slot * slot_stride + (size_t)i * 8;
  float gk[8], bk[8], ag[8], ab[8], al[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { gk[k] = ga[k]; bk[k] = be[k]; ag[k] = ab[k] = al[k] = 0.f; }
  float xv[8], da[8], dv[8];
  constexpr int U = FAN ? 1 : 4;      // rows in flight per thread (three 16-byte loads each)
  for (int r0 = 0; r0 < G.rows; r0 += U) {
    uint4 qx[U], qv[U], qd[U];
    float mu[U], rs[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (r0 + u < G.rows) {
        const int r = G.row_start + r0 + u;
        mu[u] = stats[2 * r]; rs[u] = stats[2 * r + 1];
        qx[u] = __ldg(reinterpret_cast<const uint4*>(x + (size_t)r * F) + i);
        qv[u] = __ldg(reinterpret_cast<const uint4*>(dx + (size_t)r * F) + i);
        if (!FAN) qd[u] = __ldg(reinterpret_cast<const uint4*>(dy_up + (size_t)r * F) + i);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (r0 + u >= G.rows) continue;
      if (FAN) load_da8(dy_up + (size_t)(G.row_start + r0 + u) * Hu * Wu * C, Wu, C, c8, ylo, yhi, xlo, xhi, pix / Ws, pix % Ws, da);
      else unpack8(qd[u], da);
      unpack8(qx[u], xv);
      unpack8(qv[u], dv);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xh = (xv[k] - mu[u]) * rs[u];
        const float yv = xh * gk[k] + bk[k];
        const float d = da[k] * (yv > 0.f ? 1.f : kLReLU);
        ag[k] += d * xh;
        ab[k] += d;
        al[k] += dv[k];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int f = i * 8 + k;
    const size_t o = G.slot * out_stride + (row_map ? row_map[f] : f);
    dgamma[o] += ag[k];
    dbeta[o] += ab[k];
    dbias[o] += al[k];
  }
}

// ----------------------------------------------------------------------------------------------- GroupNorm (NHWC bf16)
// One CTA (256 threads) per sample.  Thread t always owns channel octet (t % (C/8)), so per-channel partial sums stay in
// registers; group statistics are combined through shared memory.
// FAN = the consumer reads a nearest-upsampled copy (gradient fan-in over the upsample); without it dy is read directly
template <bool BWD, bool FAN>
__global__ void __launch_bounds__(256)
gn_lrelu_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy_up, int Hs, int Ws, int Hu,
                int Wu, const float* __restrict__ gamma, const float* __restrict__ beta, long slot_stride, int C,
                int groups, const es_group* __restrict__ grp, int n_groups, __nv_bfloat16* __restrict__ out,
                float* __restrict__ stats, float* __restrict__ dgamma, float* __restrict__ dbeta,
                float* __restrict__ dbias, int OWs = 0, int OWu = 0) {
  __shared__ float s_a[256], s_b[256];     // per-channel accumulators
  __shared__ float s_dg[256], s_db[256];   // per-channel affine-gradient accumulators (backward)
  __shared__ float s_g1[64], s_g2[64];     // per-group results
  __shared__ int ylo[64], yhi[64], xlo[64], xhi[64];
  const int r = blockIdx.x;
  const int g = find_group(grp, n_groups, r);
  if (g < 0) {
    // rows of skipped experts: the gradient tensor is read by the weight-gradient GEMM through TMA boxes that may
    // straddle a group's end, where it is multiplied by zero-filled im2col rows — it must be finite there
    // (forward: the strip kernels of the following conv's weight gradient read x strips past a group's end against zero dy)
    const int n16 = (BWD || OWu <= 0) ? Hs * Ws * C / 8 : (Hs * Ws / OWs) * OWu * C / 8;
    uint4* o4 = reinterpret_cast<uint4*>(out + (size_t)r * n16 * 8);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) o4[i] = make_uint4(0, 0, 0, 0);
    return;
  }
  const int slot = grp[g].slot;
  const int P = Hs * Ws, c4 = C / 8, cpg = C / groups;
  const int tid = threadIdx.x;
  const int cu = tid % c4, c8 = cu * 8, pstep = 256 / c4, p0 = tid / c4;
  const float cnt = (float)(cpg * P);
  if (BWD && FAN && tid == 0) { build_fanin(Hs, Hu, ylo, yhi); build_fanin(Ws, Wu, xlo, xhi); }
  if (!BWD && OWu > 0 && tid == 0) build_fanin(OWs, OWu, xlo, xhi);     // forward with the x-upsampled output layout
  s_a[tid] = 0.f; s_b[tid] = 0.f; s_dg[tid] = 0.f; s_db[tid] = 0.f;
  __syncthreads();
  const uint4* x4 = reinterpret_cast<const uint4*>(x + (size_t)r * P * C);
  float gk[8], bk[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { gk[k] = gamma[slot * slot_stride + c8 + k]; bk[k] = beta[slot * slot_stride + c8 + k]; }
  float f[8];
  if (!BWD) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int pix = p0; pix < P; pix += pstep) {
      unpack8(__ldg(x4 + (size_t)pix * c4 + cu), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += f[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&s_a[c8 + k], acc[k]);
    __syncthreads();
    if (tid < groups) {
      float s = 0.f;
      for (int k = 0; k < cpg; ++k) s += s_a[tid * cpg + k];
      s_g1[tid] = s / cnt;
    }
    __syncthreads();
    float mu[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { mu[k] = s_g1[(c8 + k) / cpg]; acc[k] = 0.f; }
    for (int pix = p0; pix < P; pix += pstep) {
      unpack8(__ldg(x4 + (size_t)pix * c4 + cu), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += (f[k] - mu[k]) * (f[k] - mu[k]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&s_b[c8 + k], acc[k]);
    __syncthreads();
    if (tid < groups) {
      float s = 0.f;
      for (int k = 0; k < cpg; ++k) s += s_b[tid * cpg + k];
      const float rstd = rsqrtf(s / cnt + kNormEps);
      s_g2[tid] = rstd;
      stats[((size_t)r * groups + tid) * 2] = s_g1[tid];
      stats[((size_t)r * groups + tid) * 2 + 1] = rstd;
    }
    __syncthreads();
    float rs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) rs[k] = s_g2[(c8 + k) / cpg];
    uint4* y4 = reinterpret_cast<uint4*>(out + (size_t)r * (OWu > 0 ? (P / OWs) * OWu : P) * C);
    for (int pix = p0; pix < P; pix += pstep) {
      unpack8(__ldg(x4 + (size_t)pix * c4 + cu), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = lrelu((f[k] - mu[k]) * rs[k] * gk[k] + bk[k]);
      if (OWu > 0) {      // nearest-upsampled along x: [P/OWs, OWu, C]
        const int h = pix / OWs, w = pix - h * OWs;
        const uint4 v = pack8(f);
        for (int xu = xlo[w]; xu < xhi[w]; ++xu) y4[((size_t)h * OWu + xu) * c4 + cu] = v;
      } else {
        y4[(size_t)pix * c4 + cu] = pack8(f);
      }
    }
  } else {
    float mu[8], rs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int gi = (c8 + k) / cpg;
      mu[k] = stats[((size_t)r * groups + gi) * 2];
      rs[k] = stats[((size_t)r * groups + gi) * 2 + 1];
    }
    const __nv_bfloat16* dyr = dy_up + (size_t)r * Hu * Wu * C;
    const uint4* d4 = reinterpret_cast<const uint4*>(dyr);
    float a1[8], a2[8];
    float ag[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ab[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // The loop is latency-bound unless several 16-byte loads per thread are in flight: U pixels are fetched before any
    // is consumed (measured: 2.2 TB/s with one pixel in flight).  FAN: a source pixel sums its <= 2 x 2 fan-out pixels.
    constexpr int U = FAN ? 2 : 4;
    constexpr int NQ = FAN ? 4 : 1;
    uint4 qx[U], qd[U][NQ];
    int nq[U];
    auto fetch = [&](int pix0) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pix = pix0 + u * pstep;
        nq[u] = 0;
        if (pix < P) {
          qx[u] = __ldg(x4 + (size_t)pix * c4 + cu);
          if (FAN) {
            const int sy = pix / Ws, sx = pix - sy * Ws;
            const int y0 = ylo[sy], ny = yhi[sy] - y0, x0 = xlo[sx], nx = xhi[sx] - x0;
            nq[u] = ny * nx;
#pragma unroll
            for (int j = 0; j < NQ; ++j)
              if (j < ny * nx) qd[u][j] = __ldg(d4 + ((size_t)(y0 + j / nx) * Wu + x0 + j % nx) * c4 + cu);
          } else {
            qd[u][0] = __ldg(d4 + (size_t)pix * c4 + cu);
          }
        }
      }
    };
    auto grad_of = [&](int u, float* da) {
      unpack8(qd[u][0], da);
      if (FAN) {
        float t[8];
#pragma unroll
        for (int j = 1; j < NQ; ++j)
          if (j < nq[u]) {
            unpack8(qd[u][j], t);
#pragma unroll
            for (int k = 0; k < 8; ++k) da[k] += t[k];
          }
      }
    };
    float da[8];
    for (int pix0 = p0; pix0 < P; pix0 += U * pstep) {
      fetch(pix0);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (pix0 + u * pstep >= P) continue;
        grad_of(u, da);
        unpack8(qx[u], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float xh = (f[k] - mu[k]) * rs[k];
          const float yv = xh * gk[k] + bk[k];
          const float d = da[k] * (yv > 0.f ? 1.f : kLReLU);
          ag[k] += d * xh;
          ab[k] += d;
        }
      }
    }
    // gamma is constant per channel: sum(d * gamma) = gamma * sum(d)
#pragma unroll
    for (int k = 0; k < 8; ++k) { a1[k] = gk[k] * ab[k]; a2[k] = gk[k] * ag[k]; }
#pragma unroll
    for (int k = 0; k < 8; ++k) { atomicAdd(&s_a[c8 + k], a1[k]); atomicAdd(&s_b[c8 + k], a2[k]); }
    // per-channel affine gradients: combined per CTA in shared memory, then ONE global atomic per channel per CTA
#pragma unroll
    for (int k = 0; k < 8; ++k) { atomicAdd(&s_dg[c8 + k], ag[k]); atomicAdd(&s_db[c8 + k], ab[k]); }
    __syncthreads();
    if (tid < C) {
      atomicAdd(&dgamma[slot * slot_stride + tid], s_dg[tid]);
      atomicAdd(&dbeta[slot * slot_stride + tid], s_db[tid]);
    }
    if (tid < groups) {
      float s1 = 0.f, s2 = 0.f;
      for (int k = 0; k < cpg; ++k) { s1 += s_a[tid * cpg + k]; s2 += s_b[tid * cpg + k]; }
      s_g1[tid] = s1 / cnt;
      s_g2[tid] = s2 / cnt;
    }
    __syncthreads();
    float m1[8], m2[8], al[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 8; ++k) { m1[k] = s_g1[(c8 + k) / cpg]; m2[k] = s_g2[(c8 + k) / cpg]; }
    uint4* dx4 = reinterpret_cast<uint4*>(out + (size_t)r * P * C);
    for (int pix0 = p0; pix0 < P; pix0 += U * pstep) {
      fetch(pix0);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pix = pix0 + u * pstep;
        if (pix >= P) continue;
        grad_of(u, da);
        unpack8(qx[u], f);
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float xh = (f[k] - mu[k]) * rs[k];
          const float yv = xh * gk[k] + bk[k];
          const float d = da[k] * (yv > 0.f ? 1.f : kLReLU) * gk[k];
          o[k] = rs[k] * (d - m1[k] - xh * m2[k]);
          al[k] += o[k];
        }
        dx4[(size_t)pix * c4 + cu] = pack8(o);
      }
    }
    if (dbias) {
      __syncthreads();                 // s_dg is free again: reuse it for the conv-bias gradient
      if (tid < C) s_dg[tid] = 0.f;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(&s_dg[c8 + k], al[k]);
      __syncthreads();
      if (tid < C) atomicAdd(&dbias[slot * slot_stride + tid], s_dg[tid]);
    }
  }
}

// ----------------------------------------------------------------------------------------------- GroupNorm, stats from the conv
// Forward GroupNorm + LeakyReLU as ONE streaming pass: the producing conv's epilogue already accumulated the per-(sample,
// channel pair) sum and sum of squares of the stored values (es_igemm_*_fwd_sums), so mean / rstd are 32 tiny reductions per
// CTA and every element crosses HBM once per direction with no dependency between CTAs — gridDim.y CTAs share a sample.
// Variance = E[x^2] - mean^2 in fp32 (clamped at 0); the sums are over <= 51k bf16 values of O(1) magnitude.
template <int U>                           // 16-byte loads in flight per thread
__global__ void __launch_bounds__(256)
gn_apply_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ pair_sums, const float* __restrict__ gamma,
                    const float* __restrict__ beta, long slot_stride, int Hs, int Ws, int Wu, int C, int groups,
                    const es_group* __restrict__ grp, int n_groups, __nv_bfloat16* __restrict__ out, float* __restrict__ stats) {
  __shared__ float s_mu[64], s_rs[64];
  __shared__ int xlo[64], xhi[64];
  const int r = blockIdx.x, tid = threadIdx.x;
  const int g = find_group(grp, n_groups, r);
  const int P = Hs * Ws, c4 = C / 8, cpg = C / groups;
  if (g < 0) {
    // rows of skipped experts: the strip kernels of the following conv's weight gradient read x strips that may run past a
    // group's end, where they meet zero-filled dy columns — the activation must be finite there (0 x NaN = NaN)
    const size_t n16 = (size_t)Hs * Wu * c4;
    const size_t lo = n16 * blockIdx.y / gridDim.y, hi = n16 * (blockIdx.y + 1) / gridDim.y;
    uint4* o4 = reinterpret_cast<uint4*>(out + (size_t)r * Hs * Wu * C);
    for (size_t i = lo + tid; i < hi; i += 256) o4[i] = make_uint4(0, 0, 0, 0);
    return;
  }
  const int slot = grp[g].slot;
  if (tid < groups) {
    const float* ps = pair_sums + ((size_t)r * (C / 2) + (size_t)tid * (cpg / 2)) * 2;
    float s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < cpg / 2; ++k) { s1 += ps[2 * k]; s2 += ps[2 * k + 1]; }
    const float cnt = (float)(cpg * P), mu = s1 / cnt;
    const float rstd = rsqrtf(fmaxf(s2 / cnt - mu * mu, 0.f) + kNormEps);
    s_mu[tid] = mu; s_rs[tid] = rstd;
    if (blockIdx.y == 0) {
      stats[((size_t)r * groups + tid) * 2] = mu;
      stats[((size_t)r * groups + tid) * 2 + 1] = rstd;
    }
  }
  if (Wu != Ws && tid == 0) build_fanin(Ws, Wu, xlo, xhi);
  __syncthreads();
  const int cu = tid % c4, c8 = cu * 8, pstep = 256 / c4;
  float gk[8], bk[8], mu[8], rs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    gk[k] = gamma[slot * slot_stride + c8 + k]; bk[k] = beta[slot * slot_stride + c8 + k];
    mu[k] = s_mu[(c8 + k) / cpg]; rs[k] = s_rs[(c8 + k) / cpg];
  }
  const int per = ceil_div(P, (int)gridDim.y), p_lo = blockIdx.y * per, p_hi = min(P, p_lo + per);
  const uint4* x4 = reinterpret_cast<const uint4*>(x + (size_t)r * P * C);
  uint4* y4 = reinterpret_cast<uint4*>(out + (size_t)r * Hs * Wu * C);
  for (int pix0 = p_lo + tid / c4; pix0 < p_hi; pix0 += U * pstep) {
    uint4 q[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (pix0 + u * pstep < p_hi) q[u] = __ldg(x4 + (size_t)(pix0 + u * pstep) * c4 + cu);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = pix0 + u * pstep;
      if (pix >= p_hi) continue;
      float f[8];
      unpack8(q[u], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = lrelu((f[k] - mu[k]) * rs[k] * gk[k] + bk[k]);
      const uint4 v = pack8(f);
      if (Wu != Ws) {                          // nearest-upsampled along x: [Hs, Wu, C]
        const int h = pix / Ws, w = pix - h * Ws;
        for (int xu = xlo[w]; xu < xhi[w]; ++xu) y4[((size_t)h * Wu + xu) * c4 + cu] = v;
      } else {
        y4[(size_t)pix * c4 + cu] = v;
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------- output conv (C -> 1)
__global__ void __launch_bounds__(256)
gen_out_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, long sw, long sb,
                   int Hs, int Ws, int C, int KH, int KW, int pad, const es_group* __restrict__ grp, int E,
                   int two_pass, float* __restrict__ img1, float* __restrict__ img2) {
  extern __shared__ float s_w[];  // [KH*KW][C]
  const int r = blockIdx.x;
  const RowMap m = map_row(grp, E, r, two_pass);
  if (m.g < 0) return;
  const int slot = grp[m.g].slot;
  for (int i = threadIdx.x; i < KH * KW * C; i += blockDim.x) {
    const int tap = i / C, c = i % C;   // reference layout [1][C][KH][KW]
    s_w[i] = w[slot * sw + (size_t)c * KH * KW + tap];
  }
  __syncthreads();
  const float bias = b[slot * sb];
  const int Ho = Hs + 2 * pad - KH + 1, Wo = Ws + 2 * pad - KW + 1;
  float* dst = (m.pass ? img2 : img1) + (size_t)m.j * Ho * Wo;
  const __nv_bfloat16* xr = x + (size_t)r * Hs * Ws * C;
  float f[8];
  for (int p = threadIdx.x; p < Ho * Wo; p += blockDim.x) {
    const int oy = p / Wo, ox = p % Wo;
    float acc = bias;
    for (int ky = 0; ky < KH; ++ky) {
      const int sy = oy + ky - pad;
      if (sy < 0 || sy >= Hs) continue;
      for (int kx = 0; kx < KW; ++kx) {
        const int sx = ox + kx - pad;
        if (sx < 0 || sx >= Ws) continue;
        const uint4* px = reinterpret_cast<const uint4*>(xr + ((size_t)sy * Ws + sx) * C);
        const float* wt = s_w + (ky * KW + kx) * C;
        for (int c8 = 0; c8 < C / 8; ++c8) {
          unpack8(__ldg(px + c8), f);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc = fmaf(f[k], wt[c8 * 8 + k], acc);
        }
      }
    }
    dst[p] = fmaxf(acc, 0.f);
  }
}

// C == 64, <= 4 taps: 8 lanes share an output pixel (one channel octet each, weights in registers); the <= 4 tap loads of
// a lane are independent 16-byte loads that coalesce to one 128-byte line per tap and pixel, and the 8 partial dot products
// meet in three shuffles.  (The one-thread-per-pixel kernel above issued 32 dependent loads per output: 0.49 ms for a
// 418 MB read.)
__global__ void __launch_bounds__(256)
gen_out_fwd64_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, long sw,
                     long sb, int Hs, int Ws, int KH, int KW, int pad, const es_group* __restrict__ grp, int E, int two_pass,
                     float* __restrict__ img1, float* __restrict__ img2) {
  constexpr int C = 64, c4 = 8;
  const int r = blockIdx.x;
  const RowMap m = map_row(grp, E, r, two_pass);
  if (m.g < 0) return;
  const int slot = grp[m.g].slot;
  const int cu = threadIdx.x & 7, pl = threadIdx.x >> 3;
  float wr[4][8];
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) wr[t][k] = t < KH * KW ? w[slot * sw + (size_t)(cu * 8 + k) * KH * KW + t] : 0.f;
  const float bias = b[slot * sb];
  const int Ho = Hs + 2 * pad - KH + 1, Wo = Ws + 2 * pad - KW + 1;
  float* dst = (m.pass ? img2 : img1) + (size_t)m.j * Ho * Wo;
  const uint4* x4 = reinterpret_cast<const uint4*>(x + (size_t)r * Hs * Ws * C);
  for (int p0 = 0; p0 < Ho * Wo; p0 += 32) {       // all 32 pixel lanes iterate together (shuffles below)
    const int p = p0 + pl;
    const int oy = p / Wo, ox = p - oy * Wo;
    uint4 q[4];
    bool ok[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int ky = t / KW, kx = t - ky * KW;
      const int sy = oy + ky - pad, sx = ox + kx - pad;
      ok[t] = p < Ho * Wo && t < KH * KW && sy >= 0 && sy < Hs && sx >= 0 && sx < Ws;
      if (ok[t]) q[t] = __ldg(x4 + (size_t)(sy * Ws + sx) * c4 + cu);
    }
    float acc = 0.f, f[8];
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (ok[t]) {
        unpack8(q[t], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fmaf(f[k], wr[t][k], acc);
      }
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (cu == 0 && p < Ho * Wo) dst[p] = fmaxf(acc + bias, 0.f);
  }
}

// dx[sy,sx,c] = sum_{ky,kx} dimg_masked[sy-ky+pad, sx-kx+pad] * w[c,ky,kx];  dw[c,ky,kx] += sum dimg_masked * x;  db += sum dimg_masked
// KT > 0: KH = KW = KT known at compile time — the tap loops unroll and the weight-gradient partials accw[tap][] stay in
// registers (with run-time bounds they were indexed dynamically, i.e. lived in local memory: 0.7 ms for this kernel)
template <int KT>
__global__ void __launch_bounds__(256)
gen_out_bwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, long sw, int Hs, int Ws, int C,
                   int KH_, int KW_, int pad, const float* __restrict__ img1, const float* __restrict__ img2,
                   const float* __restrict__ dimg1, const float* __restrict__ dimg2, const es_group* __restrict__ grp,
                   int E, int two_pass, __nv_bfloat16* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, long sb) {
  extern __shared__ float sm[];
  const int KH = KT > 0 ? KT : KH_, KW = KT > 0 ? KT : KW_;
  const int Ho = Hs + 2 * pad - KH + 1, Wo = Ws + 2 * pad - KW + 1;
  float* s_w = sm;                       // [KH*KW][C]
  float* s_dw = s_w + KH * KW * C;       // [KH*KW][C]
  float* s_d = s_dw + KH * KW * C;       // [Ho*Wo] masked dimg
  __shared__ float red[32];
  const int r = blockIdx.x;
  const RowMap m = map_row(grp, E, r, two_pass);
  if (m.g < 0) return;
  const int slot = grp[m.g].slot;
  for (int i = threadIdx.x; i < KH * KW * C; i += blockDim.x) {
    const int tap = i / C, c = i % C;
    s_w[i] = w[slot * sw + (size_t)c * KH * KW + tap];
    s_dw[i] = 0.f;
  }
  const float* im = (m.pass ? img2 : img1) + (size_t)m.j * Ho * Wo;
  const float* di = (m.pass ? dimg2 : dimg1) + (size_t)m.j * Ho * Wo;
  float dsum = 0.f;
  for (int p = threadIdx.x; p < Ho * Wo; p += blockDim.x) {
    const float d = im[p] > 0.f ? di[p] : 0.f;   // ReLU backward
    s_d[p] = d;
    dsum += d;
  }
  dsum = block_sum(dsum, red);   // contains the __syncthreads that publishes s_w / s_d
  if (threadIdx.x == 0) atomicAdd(&db[slot * sb], dsum);
  const int c4 = C / 8;
  const __nv_bfloat16* xr = x + (size_t)r * Hs * Ws * C;
  __nv_bfloat16* dxr = dx + (size_t)r * Hs * Ws * C;
  // thread owns channel octet cu (fixed) and walks pixels, so the weight-gradient partials stay in registers
  const int cu = threadIdx.x % c4, pstep = blockDim.x / c4;
  float accw[4][8];
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) accw[t][k] = 0.f;
  float f[8];
  constexpr int U = 4;      // pixels whose x octet is fetched before any is consumed (latency-bound otherwise)
  for (int pix0 = threadIdx.x / c4; pix0 < Hs * Ws; pix0 += U * pstep) {
   uint4 qx[U];
#pragma unroll
   for (int u = 0; u < U; ++u)
     if (pix0 + u * pstep < Hs * Ws) qx[u] = __ldg(reinterpret_cast<const uint4*>(xr + (size_t)(pix0 + u * pstep) * C) + cu);
#pragma unroll
   for (int u = 0; u < U; ++u) {
    const int pix = pix0 + u * pstep;
    if (pix >= Hs * Ws) continue;
    const int sy = pix / Ws, sx = pix % Ws;
    unpack8(qx[u], f);
    float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (KT > 0) {
#pragma unroll
      for (int ky = 0; ky < (KT > 0 ? KT : 1); ++ky) {
#pragma unroll
        for (int kx = 0; kx < (KT > 0 ? KT : 1); ++kx) {
          const int oy = sy - ky + pad, ox = sx - kx + pad;
          const bool okp = oy >= 0 && oy < Ho && ox >= 0 && ox < Wo;
          const float d = okp ? s_d[oy * Wo + ox] : 0.f;
          constexpr int KTT = KT > 0 ? KT : 1;
          const int tap = ky * KTT + kx;
          const float* wt = s_w + tap * C + cu * 8;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            o[k] = fmaf(d, wt[k], o[k]);
            if (tap < 4) accw[tap < 4 ? tap : 0][k] = fmaf(d, f[k], accw[tap < 4 ? tap : 0][k]);
          }
        }
      }
    } else {
      for (int ky = 0; ky < KH; ++ky) {
        const int oy = sy - ky + pad;
        if (oy < 0 || oy >= Ho) continue;
        for (int kx = 0; kx < KW; ++kx) {
          const int ox = sx - kx + pad;
          if (ox < 0 || ox >= Wo) continue;
          const float d = s_d[oy * Wo + ox];
          const int tap = ky * KW + kx;
          const float* wt = s_w + tap * C + cu * 8;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            o[k] = fmaf(d, wt[k], o[k]);
            if (tap < 4) accw[tap][k] = fmaf(d, f[k], accw[tap][k]);
          }
        }
      }
    }
    reinterpret_cast<uint4*>(dxr + (size_t)pix * C)[cu] = pack8(o);
   }
  }
  for (int tap = 0; tap < KH * KW && tap < 4; ++tap)
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&s_dw[tap * C + cu * 8 + k], accw[tap][k]);
  __syncthreads();
  for (int i = threadIdx.x; i < KH * KW * C; i += blockDim.x) {
    const int tap = i / C, c = i % C;
    atomicAdd(&dw[slot * sw + (size_t)c * KH * KW + tap], s_dw[i]);
  }
}

// ----------------------------------------------------------------------------------------------- packing
__global__ void pack_conv_kernel(const float* __restrict__ w, long slot_stride, int N, int C, int KH, int KW,
                                 __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd) {
  const int slot = blockIdx.y;
  const long total = (long)N * C * KH * KW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    // i indexes the forward packed layout [N][KH][KW][C]
    const int c = i % C;
    const int kx = (i / C) % KW, ky = (i / ((long)C * KW)) % KH, n = i / ((long)C * KW * KH);
    const float v = w[slot * slot_stride + (((long)n * C + c) * KH + ky) * KW + kx];
    if (wf) wf[slot * total + i] = f2bf(v);
    // data-gradient weights: [C][KH][KW][N] with the window flipped
    if (wd) wd[slot * total + (((long)c * KH + (KH - 1 - ky)) * KW + (KW - 1 - kx)) * N + n] = f2bf(v);
  }
}

__global__ void pack_dense_kernel(const float* __restrict__ w, long slot_stride, int N, int K,
                                  const int* __restrict__ row_map, __nv_bfloat16* __restrict__ wp) {
  const int slot = blockIdx.y;
  const long total = (long)N * K;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int n = i / K, k = i % K;
    const int nr = row_map ? row_map[n] : n;
    wp[slot * total + i] = f2bf(w[slot * slot_stride + (long)nr * K + k]);
  }
}

// 8 consecutive k per thread: two 16-byte loads, one 16-byte store (the scalar kernel ran at 1.5 TB/s)
__global__ void __launch_bounds__(256)
pack_dense_vec8_kernel(const float* __restrict__ w, long slot_stride, int N, int K8, const int* __restrict__ row_map,
                       __nv_bfloat16* __restrict__ wp) {
  const int slot = blockIdx.y;
  const long total8 = (long)N * K8;
  uint4* dst = reinterpret_cast<uint4*>(wp + (size_t)slot * total8 * 8);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total8; i += (long)gridDim.x * blockDim.x) {
    const int n = (int)(i / K8), k8 = (int)(i - (long)n * K8);
    const int nr = row_map ? row_map[n] : n;
    const float4* src = reinterpret_cast<const float4*>(w + slot * slot_stride + ((long)nr * K8 + k8) * 8);
    const float4 a = __ldg(src), b = __ldg(src + 1);
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    dst[i] = pack8(f);
  }
}

__global__ void unpack_conv_wgrad_kernel(const float* __restrict__ dwp, int N, int C, int KH, int KW,
                                         float* __restrict__ dw, long slot_stride) {
  const int slot = blockIdx.y;
  const long total = (long)N * C * KH * KW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    // i indexes the reference layout [N][C][KH][KW]
    const int kx = i % KW, ky = (i / KW) % KH, c = (i / ((long)KW * KH)) % C, n = i / ((long)KW * KH * C);
    dw[slot * slot_stride + i] = dwp[slot * total + (((long)n * KH + ky) * KW + kx) * C + c];
  }
}

__global__ void permute_features_kernel(const float* __restrict__ in, long in_stride, const int* __restrict__ row_map,
                                        int F, float* __restrict__ out, long out_stride, int inverse) {
  const int slot = blockIdx.y;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < F; f += gridDim.x * blockDim.x) {
    if (inverse) out[slot * out_stride + row_map[f]] = in[slot * in_stride + f];   // packed -> reference
    else out[slot * out_stride + f] = in[slot * in_stride + row_map[f]];           // reference -> packed
  }
}

// ----------------------------------------------------------------------------------------------- SIMT checkers
__global__ void igemm_fwd_simt_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                      const float* __restrict__ bias, long bias_stride, __nv_bfloat16* __restrict__ y, es_conv_geom g,
                                      const es_group* __restrict__ grp, int n_groups, long total) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int n = i % g.N;
  const long pixi = i / g.N;
  const int P = g.Ho * g.Wo;
  const int row = pixi / P, pix = pixi % P, oy = pix / g.Wo, ox = pix % g.Wo;
  const int gi = find_group(grp, n_groups, row);
  if (gi < 0) return;
  const int slot = grp[gi].slot;
  const long KK = (long)g.KH * g.KW * g.C;
  const float sy_sc = (float)g.Hs / (float)g.Hu, sx_sc = (float)g.Ws / (float)g.Wu;
  float acc = bias ? bias[slot * bias_stride + n] : 0.f;
  for (int ky = 0; ky < g.KH; ++ky)
    for (int kx = 0; kx < g.KW; ++kx) {
      const int uy = oy + ky - g.pad, ux = ox + kx - g.pad;
      if (uy < 0 || uy >= g.Hu || ux < 0 || ux >= g.Wu) continue;
      const int sy = min((int)floorf(uy * sy_sc), g.Hs - 1), sx = min((int)floorf(ux * sx_sc), g.Ws - 1);
      const __nv_bfloat16* xp = x + (((long)row * g.Hs + sy) * g.Ws + sx) * g.C;
      const __nv_bfloat16* wp = w + slot * g.N * KK + n * KK + (long)(ky * g.KW + kx) * g.C;
      for (int c = 0; c < g.C; ++c) acc = fmaf(bf2f(xp[c]), bf2f(wp[c]), acc);
    }
  y[i] = f2bf(acc);
}

__global__ void igemm_wgrad_simt_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                        float* __restrict__ dw, es_conv_geom g, const es_group* __restrict__ grp,
                                        int n_groups) {
  const long KK = (long)g.KH * g.KW * g.C;
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)g.N * KK) return;
  const es_group G = grp[blockIdx.y];
  if (G.rows == 0) return;
  const int n = i / KK;
  const int kk = i % KK, c = kk % g.C, tap = kk / g.C, ky = tap / g.KW, kx = tap % g.KW;
  const float sy_sc = (float)g.Hs / (float)g.Hu, sx_sc = (float)g.Ws / (float)g.Wu;
  float acc = 0.f;
  for (int rr = 0; rr < G.rows; ++rr) {
    const long row = G.row_start + rr;
    for (int oy = 0; oy < g.Ho; ++oy) {
      const int uy = oy + ky - g.pad;
      if (uy < 0 || uy >= g.Hu) continue;
      const int sy = min((int)floorf(uy * sy_sc), g.Hs - 1);
      for (int ox = 0; ox < g.Wo; ++ox) {
        const int ux = ox + kx - g.pad;
        if (ux < 0 || ux >= g.Wu) continue;
        const int sx = min((int)floorf(ux * sx_sc), g.Ws - 1);
        acc = fmaf(bf2f(dy[((row * g.Ho + oy) * g.Wo + ox) * g.N + n]), bf2f(x[((row * g.Hs + sy) * g.Ws + sx) * g.C + c]), acc);
      }
    }
  }
  dw[G.slot * g.N * KK + i] += acc;
}

}  // namespace es

using namespace es;

// ----------------------------------------------------------------------------------------------- x2-upsample folding
// A conv that follows a x2 nearest upsample sees, for output phase (py, px) = (oy & 1, ox & 1), only the distinct source
// rows a + dy with dy = floor((py + ky - pad) / 2): taps that land on the same source pixel are pre-summed.  Folded tap
// t = (py, px, dy, dx) of the table; forward layout [slot][n][t][c], data-gradient layout [slot][c][t][n] (both bf16),
// folded weight gradient [slot][n][t][c] (fp32) unfolded back to the reference's [n][c][ky][kx].
namespace es {
__global__ void fold_up2_kernel(const float* __restrict__ w, long sw, int N, int C, int KHW, es_fold_table t,
                                __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd) {
  const int slot = blockIdx.y;
  const long total = (long)N * t.n_taps * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C), tt = (int)((i / C) % t.n_taps), n = (int)(i / ((long)C * t.n_taps));
    const float* src = w + slot * sw + ((size_t)n * C + c) * KHW;
    float acc = 0.f;
    for (uint32_t m = t.mask[tt]; m; m &= m - 1) acc += src[__ffs(m) - 1];
    const __nv_bfloat16 v = f2bf(acc);
    if (wf) wf[(size_t)slot * total + ((size_t)n * t.n_taps + tt) * C + c] = v;
    if (wd) wd[(size_t)slot * total + ((size_t)c * t.n_taps + tt) * N + n] = v;
  }
}

// All fold tables of a layer in ONE pass over the weights (round 1b: 18 single-table launches re-read the fp32 weights 18
// times with one strided 4-byte load per output: 1.05 ms per step).  A thread owns one (n, c) filter: its KH*KW taps are
// 64 contiguous bytes, loaded once, and every table's pre-summed taps are produced from registers.  DGRAD = false: threads
// run along c (stores [n][t][c] coalesced); DGRAD = true: along n (stores [c][t][n] coalesced).
constexpr int kFoldMax = 13;
struct FoldMulti {
  int n_tables;
  es_fold_table t[kFoldMax];
  __nv_bfloat16* out[kFoldMax];
};

template <bool DGRAD>
__global__ void __launch_bounds__(256)
fold_multi_kernel(const float* __restrict__ w, long sw, int N, int C, int KHW, const __grid_constant__ FoldMulti fm) {
  const int slot = blockIdx.y;
  const long total = (long)N * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int n = DGRAD ? (int)(i % N) : (int)(i / C), c = DGRAD ? (int)(i / N) : (int)(i % C);
    const float* src = w + slot * sw + ((size_t)n * C + c) * KHW;
    float v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = k < KHW ? __ldg(src + k) : 0.f;
    for (int j = 0; j < fm.n_tables; ++j) {
      const int T = fm.t[j].n_taps;
      __nv_bfloat16* o = fm.out[j] + (size_t)slot * total * T;
      for (int tt = 0; tt < T; ++tt) {
        const uint32_t m = fm.t[j].mask[tt];
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if ((m >> k) & 1u) acc += v[k];         // ascending tap order, like the single-table kernel
        if (DGRAD) o[((size_t)c * T + tt) * N + n] = f2bf(acc);
        else o[((size_t)n * T + tt) * C + c] = f2bf(acc);
      }
    }
  }
}

// thread = (n, c): reads of the folded gradient are coalesced over c, each thread writes its KH*KW contiguous outputs
__global__ void unfold_up2_kernel(const float* __restrict__ dwf, int N, int C, int KHW, es_fold_table t,
                                  float* __restrict__ dw, long sw) {
  const int slot = blockIdx.y;
  const long total = (long)N * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C), n = (int)(i / C);
    const float* src = dwf + ((size_t)slot * N + n) * t.n_taps * C + c;
    float* dst = dw + slot * sw + (size_t)i * KHW;
    for (int tt = 0; tt < t.n_taps; ++tt) {
      const float v = src[(size_t)tt * C];
      for (uint32_t m = t.mask[tt]; m; m &= m - 1) dst[__ffs(m) - 1] += v;
    }
  }
}

// all classes' folded weight gradients -> the reference layout in one read-modify-write of dw
struct UnfoldMulti {
  int n_tables;
  es_fold_table t[kFoldMax];
  const float* in[kFoldMax];
};
__global__ void __launch_bounds__(256)
unfold_multi_kernel(int N, int C, int KHW, const __grid_constant__ UnfoldMulti um, float* __restrict__ dw, long sw) {
  const int slot = blockIdx.y;
  const long total = (long)N * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C), n = (int)(i / C);
    float acc[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = 0.f;
    for (int j = 0; j < um.n_tables; ++j) {
      const int T = um.t[j].n_taps;
      const float* src = um.in[j] + ((size_t)slot * N + n) * T * C + c;
      for (int tt = 0; tt < T; ++tt) {
        const float v = __ldg(src + (size_t)tt * C);
        const uint32_t m = um.t[j].mask[tt];
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if ((m >> k) & 1u) acc[k] += v;
      }
    }
    float* dst = dw + slot * sw + (size_t)i * KHW;
#pragma unroll
    for (int k = 0; k < 32; ++k)
      if (k < KHW) dst[k] += acc[k];
  }
}

// dst[row, a*Wo + b, :] = src[row, (a*my + oy)*Wf + b*mx + ox, :]   (one output phase of an NHWC bf16 map, 16-byte chunks)
__global__ void pick_pixels_kernel(const __nv_bfloat16* __restrict__ src, int Pf, int Wf, int C, int my, int oy, int mx, int ox,
                                   int Ho, int Wo, __nv_bfloat16* __restrict__ dst) {
  const int row = blockIdx.x, c8 = C / 8, n = Ho * Wo * c8;
  const uint4* s4 = reinterpret_cast<const uint4*>(src + (size_t)row * Pf * C);
  uint4* d4 = reinterpret_cast<uint4*>(dst + (size_t)row * Ho * Wo * C);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int pix = i / c8, ch = i - pix * c8, a = pix / Wo, b = pix - a * Wo;
    d4[i] = s4[(size_t)((a * my + oy) * Wf + b * mx + ox) * c8 + ch];
  }
}
}  // namespace es

extern "C" int es_fold_up2_weights(const float* w, long slot_stride, int slots, int N, int C, int KH, int KW,
                                   const es_fold_table* t, void* w_fwd, void* w_dgrad, void* stream) {
  ES_REQUIRE(w && t && (w_fwd || w_dgrad) && slots >= 1 && N > 0 && C > 0 && t->n_taps >= 1 && t->n_taps <= 32, "bad arguments");
  ES_REQUIRE(KH * KW <= 32, "at most 32 original taps");
  es::fold_up2_kernel<<<dim3(512, slots), 256, 0, es::as_stream(stream)>>>(w, slot_stride, N, C, KH * KW, *t,
                                                                            (__nv_bfloat16*)w_fwd, (__nv_bfloat16*)w_dgrad);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_fold_weights_multi(const float* w, long slot_stride, int slots, int N, int C, int KH, int KW,
                                     const es_fold_table* tables, int n_tables, void* const* outs, int dgrad_layout,
                                     void* stream) {
  ES_REQUIRE(w && tables && outs && slots >= 1 && N > 0 && C > 0 && KH * KW <= 32, "bad arguments");
  ES_REQUIRE(n_tables >= 1 && n_tables <= es::kFoldMax, "1..13 tables per call");
  es::FoldMulti fm{};
  fm.n_tables = n_tables;
  for (int j = 0; j < n_tables; ++j) {
    ES_REQUIRE(tables[j].n_taps >= 1 && tables[j].n_taps <= 32 && outs[j], "bad table / output");
    fm.t[j] = tables[j];
    fm.out[j] = (__nv_bfloat16*)outs[j];
  }
  const dim3 grid((unsigned)min(es::ceil_div_l((long)N * C, 256), 148L * 4), slots);
  if (dgrad_layout) es::fold_multi_kernel<true><<<grid, 256, 0, es::as_stream(stream)>>>(w, slot_stride, N, C, KH * KW, fm);
  else es::fold_multi_kernel<false><<<grid, 256, 0, es::as_stream(stream)>>>(w, slot_stride, N, C, KH * KW, fm);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_unfold_up2_wgrad(const float* dw_folded, int slots, int N, int C, int KH, int KW,
                                   const es_fold_table* t, float* dw_ref, long slot_stride, void* stream) {
  ES_REQUIRE(dw_folded && t && dw_ref && slots >= 1 && t->n_taps >= 1 && t->n_taps <= 32 && KH * KW <= 32, "bad arguments");
  es::unfold_up2_kernel<<<dim3(512, slots), 256, 0, es::as_stream(stream)>>>(dw_folded, N, C, KH * KW, *t, dw_ref, slot_stride);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_unfold_wgrad_multi(const float* const* dw_folded, int slots, int N, int C, int KH, int KW,
                                     const es_fold_table* tables, int n_tables, float* dw_ref, long slot_stride,
                                     void* stream) {
  ES_REQUIRE(dw_folded && tables && dw_ref && slots >= 1 && N > 0 && C > 0 && KH * KW <= 32, "bad arguments");
  ES_REQUIRE(n_tables >= 1 && n_tables <= es::kFoldMax, "1..13 tables per call");
  es::UnfoldMulti um{};
  um.n_tables = n_tables;
  for (int j = 0; j < n_tables; ++j) {
    ES_REQUIRE(tables[j].n_taps >= 1 && tables[j].n_taps <= 32 && dw_folded[j], "bad table / input");
    um.t[j] = tables[j];
    um.in[j] = dw_folded[j];
  }
  const dim3 grid((unsigned)min(es::ceil_div_l((long)N * C, 256), 148L * 4), slots);
  es::unfold_multi_kernel<<<grid, 256, 0, es::as_stream(stream)>>>(N, C, KH * KW, um, dw_ref, slot_stride);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_pick_pixels(const void* src, int Ho_full, int Wo_full, int C, int my, int oy, int mx, int ox, int Ho, int Wo,
                              int total_rows, void* dst, void* stream) {
  ES_REQUIRE(src && dst && C % 8 == 0 && total_rows > 0 && Ho > 0 && Wo > 0, "bad arguments");
  ES_REQUIRE((Ho - 1) * my + oy < Ho_full && (Wo - 1) * mx + ox < Wo_full, "pixel selection out of range");
  es::pick_pixels_kernel<<<total_rows, 256, 0, es::as_stream(stream)>>>((const __nv_bfloat16*)src, Ho_full * Wo_full, Wo_full, C, my,
                                                                        oy, mx, ox, Ho, Wo, (__nv_bfloat16*)dst);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_gen_fc1_fwd(const float* z1, const float* z2, const float* cond, const float* w, const float* b,
                              const float* gamma, const float* beta, long slot_stride_w, long slot_stride_v,
                              const es_group* grp_gen, int E, int total_rows, int two_pass, float* x0, float* lin,
                              void* h, void* stream) {
  ES_REQUIRE(z1 && cond && w && b && grp_gen && x0 && h && (gamma ? (beta && lin) : true), "null pointer");
  ES_REQUIRE(!two_pass || z2, "two-pass batch needs z2");
  ES_REQUIRE(E >= 1 && E <= kMaxGroups && total_rows > 0, "bad sizes");
  gen_fc1_fwd_kernel<<<ceil_div(total_rows, 8), 256, 0, as_stream(stream)>>>(
      z1, z2, cond, w, b, gamma, beta, slot_stride_w, slot_stride_v, grp_gen, E, total_rows, two_pass, x0, lin,
      (__nv_bfloat16*)h);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_gen_fc1_bwd(const float* dh, const float* x0, const float* lin, const float* gamma,
                              const float* beta, long slot_stride_w, long slot_stride_v, const es_group* grp_gen,
                              int E, int total_rows, float* dw, float* db, float* dgamma, float* dbeta, void* stream) {
  ES_REQUIRE(dh && x0 && grp_gen && dw && db && (gamma ? (lin && beta && dgamma && dbeta) : true), "null pointer");
  ES_REQUIRE(E >= 1 && E <= kMaxGroups && total_rows > 0, "bad sizes");
  const int chunks = max(1, min(32, ceil_div(total_rows, 8 * 4 * E)));
  gen_fc1_bwd_kernel<<<E * chunks, 256, 0, as_stream(stream)>>>(dh, x0, lin, gamma, beta, slot_stride_w,
                                                                slot_stride_v, grp_gen, chunks, dw, db, dgamma, dbeta);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_ln_lrelu_fwd(const void* x, const float* gamma, const float* beta, long slot_stride, int F,
                               const es_group* grp, int n_groups, int total_rows, void* y, float* stats, void* stream) {
  ES_REQUIRE(x && gamma && beta && grp && y && stats, "null pointer");
  ES_REQUIRE(F > 0 && F % 8 == 0 && total_rows > 0 && n_groups >= 1 && n_groups <= kMaxGroups, "bad sizes");
  ln_lrelu_fwd_kernel<<<total_rows, 512, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, gamma, beta, slot_stride, F,
                                                                 grp, n_groups, (__nv_bfloat16*)y, stats);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_ln_lrelu_bwd(const void* dy_up, int Hs, int Ws, int Hu, int Wu, int C, const void* x,
                               const float* stats, const float* gamma, const float* beta, long slot_stride,
                               const es_group* grp, int n_groups, int total_rows, void* dx, void* stream) {
  ES_REQUIRE(dy_up && x && stats && gamma && beta && grp && dx, "null pointer");
  ES_REQUIRE(C % 8 == 0 && Hs <= 64 && Ws <= 64 && Hu <= 64 && Wu <= 64 && Hu >= Hs && Wu >= Ws, "bad geometry");
  ES_REQUIRE(total_rows > 0 && n_groups >= 1 && n_groups <= kMaxGroups, "bad sizes");
  ES_REQUIRE(slot_stride % 4 == 0 && (((uintptr_t)gamma | (uintptr_t)beta) & 15) == 0, "gamma/beta must be 16-byte aligned");
  if (Hu == Hs && Wu == Ws)
    ln_lrelu_bwd_kernel<false><<<total_rows, 512, 0, as_stream(stream)>>>((const __nv_bfloat16*)dy_up, Hs, Ws, Hu, Wu, C,
                                                                          (const __nv_bfloat16*)x, stats, gamma, beta,
                                                                          slot_stride, grp, n_groups, (__nv_bfloat16*)dx);
  else
    ln_lrelu_bwd_kernel<true><<<total_rows, 512, 0, as_stream(stream)>>>((const __nv_bfloat16*)dy_up, Hs, Ws, Hu, Wu, C,
                                                                         (const __nv_bfloat16*)x, stats, gamma, beta,
                                                                         slot_stride, grp, n_groups, (__nv_bfloat16*)dx);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_ln_affine_bwd(const void* dy_up, int Hs, int Ws, int Hu, int Wu, int C, const void* x,
                                const void* dx, const float* stats, const float* gamma, const float* beta,
                                long slot_stride, const es_group* grp, int n_groups, int total_rows,
                                const int32_t* row_map, long out_slot_stride, float* dgamma, float* dbeta, float* dbias_lin, void* stream) {
  ES_REQUIRE(dy_up && x && dx && stats && gamma && beta && grp && dgamma && dbeta && dbias_lin, "null pointer");
  ES_REQUIRE(C % 8 == 0 && Hs <= 64 && Ws <= 64 && Hu <= 64 && Wu <= 64, "bad geometry");
  ES_REQUIRE(total_rows > 0 && n_groups >= 1 && n_groups <= kMaxGroups, "bad sizes");
  const int n4 = Hs * Ws * C / 8;
  if (Hu == Hs && Wu == Ws)
    ln_affine_bwd_kernel<false><<<dim3(ceil_div(n4, 128), n_groups), 128, 0, as_stream(stream)>>>(
        (const __nv_bfloat16*)dy_up, Hs, Ws, Hu, Wu, C, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dx, stats, gamma,
        beta, slot_stride, grp, row_map, out_slot_stride, dgamma, dbeta, dbias_lin);
  else
    ln_affine_bwd_kernel<true><<<dim3(ceil_div(n4, 128), n_groups), 128, 0, as_stream(stream)>>>(
        (const __nv_bfloat16*)dy_up, Hs, Ws, Hu, Wu, C, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dx, stats, gamma,
        beta, slot_stride, grp, row_map, out_slot_stride, dgamma, dbeta, dbias_lin);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

namespace es {
int gn_cluster_fwd(const void* x, const float* gamma, const float* beta, long slot_stride, int P, int C, int groups,
                   const es_group* grp, int n_groups, int total_rows, void* y, float* stats, void* stream, int OWs, int OWu);
int gn_cluster_bwd(const void* dy_up, int Hs, int Ws, int Hu, int Wu, const void* x, const float* stats,
                   const float* gamma, const float* beta, long slot_stride, int C, int groups, const es_group* grp,
                   int n_groups, int total_rows, void* dx, float* dgamma, float* dbeta, float* dbias, void* stream);
// ES_GN_LEGACY=1 selects the one-CTA-per-sample kernel (kept for A/B measurements)
static bool gn_legacy() {
  static const bool v = [] { const char* e = getenv("ES_GN_LEGACY"); return e && e[0] == '1'; }();
  return v;
}
// ES_GN_CLUSTER_BWD=1 selects the cluster kernel for the backward too (A/B measurements: it loses to the batched loads)
static bool gn_cluster_backward() {
  static const bool v = [] { const char* e = getenv("ES_GN_CLUSTER_BWD"); return e && e[0] == '1'; }();
  return v;
}
}  // namespace es

static int gn_lrelu_fwd_impl(const void* x, const float* gamma, const float* beta, long slot_stride, int P, int C,
                             int groups, const es_group* grp, int n_groups, int total_rows, void* y, float* stats,
                             void* stream, int OWs, int OWu) {
  ES_REQUIRE(x && gamma && beta && grp && y && stats, "null pointer");
  ES_REQUIRE(C % 8 == 0 && C <= 256 && 256 % (C / 8) == 0 && groups <= 64 && C % groups == 0, "bad channels/groups");
  ES_REQUIRE(total_rows > 0 && n_groups >= 1 && n_groups <= kMaxGroups && P > 0, "bad sizes");
  if (!gn_legacy()) {   // cluster kernel (gn_cluster.cu): the sample's slab stays in shared memory; 1 = does not fit
    const int rc = gn_cluster_fwd(x, gamma, beta, slot_stride, P, C, groups, grp, n_groups, total_rows, y, stats, stream, OWs, OWu);
    if (rc != 1) return rc;
  }
  gn_lrelu_kernel<false, false><<<total_rows, 256, 0, as_stream(stream)>>>(
      (const __nv_bfloat16*)x, nullptr, P, 1, P, 1, gamma, beta, slot_stride, C, groups, grp, n_groups,
      (__nv_bfloat16*)y, stats, nullptr, nullptr, nullptr, OWs, OWu);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_gn_lrelu_fwd(const void* x, const float* gamma, const float* beta, long slot_stride, int P, int C,
                               int groups, const es_group* grp, int n_groups, int total_rows, void* y, float* stats,
                               void* stream) {
  return gn_lrelu_fwd_impl(x, gamma, beta, slot_stride, P, C, groups, grp, n_groups, total_rows, y, stats, stream, 0, 0);
}

extern "C" int es_gn_lrelu_fwd_upx(const void* x, const float* gamma, const float* beta, long slot_stride, int Hs, int Ws,
                                   int Wu, int C, int groups, const es_group* grp, int n_groups, int total_rows, void* y,
                                   float* stats, void* stream) {
  ES_REQUIRE(Hs > 0 && Ws > 0 && Ws <= 64 && Wu >= Ws && Wu <= 64 && Wu <= 2 * Ws, "bad x-upsample geometry");
  return gn_lrelu_fwd_impl(x, gamma, beta, slot_stride, Hs * Ws, C, groups, grp, n_groups, total_rows, y, stats, stream, Ws, Wu);
}

extern "C" int es_gn_lrelu_apply_fwd(const void* x, const float* pair_sums, const float* gamma, const float* beta, long slot_stride,
                                     int Hs, int Ws, int Wu, int C, int groups, const es_group* grp, int n_groups, int total_rows,
                                     void* y, float* stats, void* stream) {
  ES_REQUIRE(x && pair_sums && gamma && beta && grp && y && stats, "null pointer");
  ES_REQUIRE(C % 8 == 0 && C <= 256 && 256 % (C / 8) == 0 && groups <= 64 && C % groups == 0 && (C / groups) % 2 == 0,
             "bad channels/groups (channel pairs must not straddle groups)");
  ES_REQUIRE(total_rows > 0 && n_groups >= 1 && n_groups <= kMaxGroups && Hs > 0 && Ws > 0, "bad sizes");
  ES_REQUIRE(Wu == Ws || (Ws <= 64 && Wu > Ws && Wu <= 64 && Wu <= 2 * Ws), "bad x-upsample geometry");
  const long bytes = (long)Hs * Ws * C * 2;
  static const long chunk_bytes = [] { const char* e = getenv("ES_GN_APPLY_KB"); return (long)(e ? atoi(e) : 512) * 1024; }();
  int chunks = (int)((bytes + chunk_bytes - 1) / chunk_bytes);
  chunks = chunks < 1 ? 1 : (chunks > 64 ? 64 : chunks);
  static const int u_sel = [] { const char* e = getenv("ES_GN_APPLY_U"); return e ? atoi(e) : 8; }();
  if (u_sel == 4)
    gn_apply_fwd_kernel<4><<<dim3(total_rows, chunks), 256, 0, as_stream(stream)>>>(
        (const __nv_bfloat16*)x, pair_sums, gamma, beta, slot_stride, Hs, Ws, Wu, C, groups, grp, n_groups, (__nv_bfloat16*)y, stats);
  else
    gn_apply_fwd_kernel<8><<<dim3(total_rows, chunks), 256, 0, as_stream(stream)>>>(
        (const __nv_bfloat16*)x, pair_sums, gamma, beta, slot_stride, Hs, Ws, Wu, C, groups, grp, n_groups, (__nv_bfloat16*)y, stats);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_gn_lrelu_bwd(const void* dy_up, int Hs, int Ws, int Hu, int Wu, const void* x, const float* stats,
                               const float* gamma, const float* beta, long slot_stride, int C, int groups,
                               const es_group* grp, int n_groups, int total_rows, void* dx, float* dgamma,
                               float* dbeta, float* dbias_conv, void* stream) {
  ES_REQUIRE(dy_up && x && stats && gamma && beta && grp && dx && dgamma && dbeta, "null pointer");
  ES_REQUIRE(C % 8 == 0 && C <= 256 && 256 % (C / 8) == 0 && groups <= 64 && C % groups == 0, "bad channels/groups");
  ES_REQUIRE(Hs <= 64 && Ws <= 64 && Hu <= 64 && Wu <= 64 && Hu >= Hs && Wu >= Ws, "bad geometry");
  ES_REQUIRE(total_rows > 0 && n_groups >= 1 && n_groups <= kMaxGroups, "bad sizes");
  ES_REQUIRE(Hu <= 2 * Hs && Wu <= 2 * Ws, "fan-in of more than 2 per axis");
  if (gn_cluster_backward()) {   // measured slower than the batched-load kernel below (r01 bench_gn): opt-in only
    const int rc = gn_cluster_bwd(dy_up, Hs, Ws, Hu, Wu, x, stats, gamma, beta, slot_stride, C, groups, grp, n_groups,
                                  total_rows, dx, dgamma, dbeta, dbias_conv, stream);
    if (rc != 1) return rc;
  }
  if (Hu == Hs && Wu == Ws)
    gn_lrelu_kernel<true, false><<<total_rows, 256, 0, as_stream(stream)>>>(
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy_up, Hs, Ws, Hu, Wu, gamma, beta, slot_stride, C, groups, grp,
        n_groups, (__nv_bfloat16*)dx, const_cast<float*>(stats), dgamma, dbeta, dbias_conv);
  else
    gn_lrelu_kernel<true, true><<<total_rows, 256, 0, as_stream(stream)>>>(
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy_up, Hs, Ws, Hu, Wu, gamma, beta, slot_stride, C, groups, grp,
        n_groups, (__nv_bfloat16*)dx, const_cast<float*>(stats), dgamma, dbeta, dbias_conv);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_gen_out_fwd(const void* x, const float* w, const float* b, long slot_stride_w, long slot_stride_b, int Hs, int Ws, int C,
                              int KH, int KW, int pad, const es_group* grp_gen, int E, int total_rows, int two_pass,
                              float* img1, float* img2, void* stream) {
  ES_REQUIRE(x && w && b && grp_gen && img1 && (img2 || !two_pass), "null pointer");
  ES_REQUIRE(C % 8 == 0 && KH * KW * C <= 8192 && total_rows > 0 && E >= 1 && E <= kMaxGroups, "bad sizes");
  if (C == 64 && KH * KW <= 4) {
    gen_out_fwd64_kernel<<<total_rows, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, w, b, slot_stride_w, slot_stride_b,
                                                                    Hs, Ws, KH, KW, pad, grp_gen, E, two_pass, img1, img2);
    ES_LAUNCH_CHECK();
    return ES_OK;
  }
  gen_out_fwd_kernel<<<total_rows, 256, KH * KW * C * sizeof(float), as_stream(stream)>>>(
      (const __nv_bfloat16*)x, w, b, slot_stride_w, slot_stride_b, Hs, Ws, C, KH, KW, pad, grp_gen, E, two_pass, img1, img2);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_gen_out_bwd(const void* x, const float* w, long slot_stride_w, long slot_stride_b, int Hs, int Ws, int C, int KH, int KW,
                              int pad, const float* img1, const float* img2, const float* dimg1, const float* dimg2,
                              const es_group* grp_gen, int E, int total_rows, int two_pass, void* dx, float* dw,
                              float* db, void* stream) {
  ES_REQUIRE(x && w && img1 && dimg1 && grp_gen && dx && dw && db, "null pointer");
  ES_REQUIRE(!two_pass || (img2 && dimg2), "two-pass batch needs img2/dimg2");
  ES_REQUIRE(C % 8 == 0 && 256 % (C / 8) == 0 && KH * KW <= 4 && total_rows > 0 && E >= 1 && E <= kMaxGroups, "bad sizes");
  const int Ho = Hs + 2 * pad - KH + 1, Wo = Ws + 2 * pad - KW + 1;
  const size_t smem = (2 * KH * KW * C + Ho * Wo) * sizeof(float);
  if (KH == 2 && KW == 2)
    gen_out_bwd_kernel<2><<<total_rows, 256, smem, as_stream(stream)>>>((const __nv_bfloat16*)x, w, slot_stride_w, Hs, Ws, C,
        KH, KW, pad, img1, img2, dimg1, dimg2, grp_gen, E, two_pass, (__nv_bfloat16*)dx, dw, db, slot_stride_b);
  else
  gen_out_bwd_kernel<0><<<total_rows, 256, smem, as_stream(stream)>>>((const __nv_bfloat16*)x, w, slot_stride_w, Hs, Ws, C,
                                                                   KH, KW, pad, img1, img2, dimg1, dimg2, grp_gen, E,
                                                                   two_pass, (__nv_bfloat16*)dx, dw, db, slot_stride_b);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_pack_conv_weight(const float* w, long slot_stride, int slots, int N, int C, int KH, int KW,
                                   void* w_fwd, void* w_dgrad, void* stream) {
  ES_REQUIRE(w && (w_fwd || w_dgrad) && slots >= 1 && N > 0 && C > 0 && KH > 0 && KW > 0, "bad arguments");
  const long total = (long)N * C * KH * KW;
  pack_conv_kernel<<<dim3((unsigned)min(ceil_div_l(total, 256), 2048L), slots), 256, 0, as_stream(stream)>>>(
      w, slot_stride, N, C, KH, KW, (__nv_bfloat16*)w_fwd, (__nv_bfloat16*)w_dgrad);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_pack_dense_weight(const float* w, long slot_stride, int slots, int N, int K, const int32_t* row_map,
                                    void* w_packed, void* stream) {
  ES_REQUIRE(w && w_packed && slots >= 1 && N > 0 && K > 0, "bad arguments");
  const long total = (long)N * K;
  if (K % 8 == 0 && slot_stride % 4 == 0 && (((uintptr_t)w | (uintptr_t)w_packed) & 15) == 0) {
    pack_dense_vec8_kernel<<<dim3((unsigned)min(ceil_div_l(total / 8, 256), 148L * 16), slots), 256, 0, as_stream(stream)>>>(
        w, slot_stride, N, K / 8, row_map, (__nv_bfloat16*)w_packed);
    ES_LAUNCH_CHECK();
    return ES_OK;
  }
  pack_dense_kernel<<<dim3((unsigned)min(ceil_div_l(total, 256), 4096L), slots), 256, 0, as_stream(stream)>>>(
      w, slot_stride, N, K, row_map, (__nv_bfloat16*)w_packed);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_unpack_conv_wgrad(const float* dw_packed, int slots, int N, int C, int KH, int KW, float* dw_ref,
                                    long slot_stride, void* stream) {
  ES_REQUIRE(dw_packed && dw_ref && slots >= 1, "bad arguments");
  const long total = (long)N * C * KH * KW;
  unpack_conv_wgrad_kernel<<<dim3((unsigned)min(ceil_div_l(total, 256), 2048L), slots), 256, 0, as_stream(stream)>>>(
      dw_packed, N, C, KH, KW, dw_ref, slot_stride);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_permute_features(const float* in, long in_stride, const int32_t* row_map, int slots, int F, float* out,
                                   long out_stride, int inverse, void* stream) {
  ES_REQUIRE(in && row_map && out && slots >= 1 && F > 0, "bad arguments");
  permute_features_kernel<<<dim3(min(ceil_div(F, 256), 1024), slots), 256, 0, as_stream(stream)>>>(
      in, in_stride, row_map, F, out, out_stride, inverse);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_igemm_fwd_simt(const void* x, const void* w, const float* bias, long bias_slot_stride, void* y,
                                 const es_conv_geom* g, const es_group* grp, int n_groups, int total_rows, void* stream) {
  ES_REQUIRE(x && w && y && g && grp && total_rows > 0, "bad arguments");
  const long total = (long)total_rows * g->Ho * g->Wo * g->N;
  igemm_fwd_simt_kernel<<<(unsigned)ceil_div_l(total, 256), 256, 0, as_stream(stream)>>>(
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, bias, bias_slot_stride, (__nv_bfloat16*)y, *g, grp, n_groups, total);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_igemm_wgrad_simt(const void* x, const void* dy, float* dw, const es_conv_geom* g,
                                   const es_group* grp, int n_groups, int total_rows, void* stream) {
  ES_REQUIRE(x && dy && dw && g && grp && total_rows > 0, "bad arguments");
  const long total = (long)g->N * g->KH * g->KW * g->C;
  igemm_wgrad_simt_kernel<<<dim3((unsigned)ceil_div_l(total, 256), n_groups), 256, 0, as_stream(stream)>>>(
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, dw, *g, grp, n_groups);
  ES_LAUNCH_CHECK();
  return ES_OK;
}
