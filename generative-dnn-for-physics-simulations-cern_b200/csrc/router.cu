// K1 — router gating, gumbel softmax, argmax, expert histogram and the sort-free STABLE token->expert permutation,
// plus the router-loss backward.  Replaces RouterNetwork.forward (expertsim/models/routers/router.py:21-26) and the
// routing block of MoEWrapper.train_step (expertsim/models/moe.py:76-77,97-103,123-126,407-442).
//
// Layout: one CTA = 256 consecutive samples (8 warps x 32 samples); the 11.7k router weights live in shared memory.
// The MLP is accumulated in fp64 and rounded to fp32 per layer so the argmax agrees with the reference's fp32
// CPU/GPU result except on (measure-zero) exact ties; ties resolve to the first maximal index as torch.argmax.
#include "common.cuh"

namespace es {

constexpr int kRouterBlock = 256;
constexpr int kH1 = 128, kH2 = 64, kH3 = 32, kCond = 9;
constexpr int kMaxE = 16;

struct RouterSmem {
  // padded strides (in + 1) keep the per-lane row reads conflict-free
  static constexpr int W0 = 0;                          // [128][9]
  static constexpr int B0 = W0 + kH1 * kCond;           // [128]
  static constexpr int W2 = B0 + kH1;                   // [64][129]
  static constexpr int B2 = W2 + kH2 * (kH1 + 1);       // [64]
  static constexpr int W4 = B2 + kH2;                   // [32][65]
  static constexpr int B4 = W4 + kH3 * (kH2 + 1);       // [32]
  static constexpr int W6 = B4 + kH3;                   // [16][33]
  static constexpr int B6 = W6 + kMaxE * (kH3 + 1);     // [16]
  static constexpr int HB = B6 + kMaxE;                 // per-warp scratch [8][9+128+64+32]
  static constexpr int HB_PER_WARP = 16 + kH1 + kH2 + kH3;
  static constexpr int TOTAL = HB + 8 * HB_PER_WARP;
};

__global__ void __launch_bounds__(kRouterBlock)
router_fwd_kernel(const float* __restrict__ cond, int B, int E,
                  const float* __restrict__ w0, const float* __restrict__ b0, const float* __restrict__ w2,
                  const float* __restrict__ b2, const float* __restrict__ w4, const float* __restrict__ b4,
                  const float* __restrict__ w6, const float* __restrict__ b6,
                  const float* __restrict__ gumbel, float tau,
                  float* __restrict__ logits, float* __restrict__ gates, int64_t* __restrict__ idx,
                  float* __restrict__ h1o, float* __restrict__ h2o, float* __restrict__ h3o,
                  int32_t* __restrict__ blk_hist) {
  extern __shared__ float sm[];
  __shared__ int s_hist[kMaxE];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kH1 * kCond; i += kRouterBlock) sm[RouterSmem::W0 + i] = w0[i];
  for (int i = tid; i < kH1; i += kRouterBlock) sm[RouterSmem::B0 + i] = b0[i];
  for (int i = tid; i < kH2 * kH1; i += kRouterBlock) sm[RouterSmem::W2 + (i / kH1) * (kH1 + 1) + (i % kH1)] = w2[i];
  for (int i = tid; i < kH2; i += kRouterBlock) sm[RouterSmem::B2 + i] = b2[i];
  for (int i = tid; i < kH3 * kH2; i += kRouterBlock) sm[RouterSmem::W4 + (i / kH2) * (kH2 + 1) + (i % kH2)] = w4[i];
  for (int i = tid; i < kH3; i += kRouterBlock) sm[RouterSmem::B4 + i] = b4[i];
  for (int i = tid; i < E * kH3; i += kRouterBlock) sm[RouterSmem::W6 + (i / kH3) * (kH3 + 1) + (i % kH3)] = w6[i];
  for (int i = tid; i < E; i += kRouterBlock) sm[RouterSmem::B6 + i] = b6[i];
  if (tid < kMaxE) s_hist[tid] = 0;
  __syncthreads();

  float* hb = sm + RouterSmem::HB + warp * RouterSmem::HB_PER_WARP;
  float* xs = hb;            // [9] (padded to 16)
  float* h1 = hb + 16;       // [128]
  float* h2 = h1 + kH1;      // [64]
  float* h3 = h2 + kH2;      // [32]
  int my_idx = -1;           // lane i keeps the expert of the warp's i-th sample

  for (int i = 0; i < 32; ++i) {
    const int b = blockIdx.x * kRouterBlock + warp * 32 + i;
    if (b >= B) break;   // warp-uniform
    if (lane < kCond) xs[lane] = cond[(size_t)b * kCond + lane];
    __syncwarp();
    // layer 0: 128 outputs, 4 per lane
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int o = lane + 32 * r;
      double acc = sm[RouterSmem::B0 + o];
      const float* wr = sm + RouterSmem::W0 + o * kCond;
#pragma unroll
      for (int k = 0; k < kCond; ++k) acc += (double)wr[k] * (double)xs[k];
      h1[o] = lrelu((float)acc);
    }
    __syncwarp();
    // layer 2: 64 outputs, 2 per lane
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int o = lane + 32 * r;
      double acc = sm[RouterSmem::B2 + o];
      const float* wr = sm + RouterSmem::W2 + o * (kH1 + 1);
#pragma unroll 8
      for (int k = 0; k < kH1; ++k) acc += (double)wr[k] * (double)h1[k];
      h2[o] = lrelu((float)acc);
    }
    __syncwarp();
    {
      double acc = sm[RouterSmem::B4 + lane];
      const float* wr = sm + RouterSmem::W4 + lane * (kH2 + 1);
#pragma unroll 8
      for (int k = 0; k < kH2; ++k) acc += (double)wr[k] * (double)h2[k];
      h3[lane] = lrelu((float)acc);
    }
    __syncwarp();
    float logit = 0.f, y = -INFINITY;
    if (lane < E) {
      double acc = sm[RouterSmem::B6 + lane];
      const float* wr = sm + RouterSmem::W6 + lane * (kH3 + 1);
#pragma unroll 8
      for (int k = 0; k < kH3; ++k) acc += (double)wr[k] * (double)h3[k];
      logit = (float)acc;
      y = (logit + gumbel[(size_t)b * E + lane]) / tau;   // F.gumbel_softmax: (logits + gumbels) / tau
    }
    // softmax over the E live lanes (torch: exp(x - max) / sum)
    const float mx = warp_max(y);
    const float ex = lane < E ? expf(y - mx) : 0.f;
    const float den = warp_sum(ex);
    const float gate = ex / den;
    // argmax over gates, first maximal index
    float best = lane < E ? gate : -1.f;
    int besti = lane < E ? lane : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    if (lane < E) {
      logits[(size_t)b * E + lane] = logit;
      gates[(size_t)b * E + lane] = gate;
    }
    if (lane == 0) idx[b] = besti;
    if (lane == i) my_idx = besti;
    // keep the hidden activations for the backward pass
#pragma unroll
    for (int r = 0; r < 4; ++r) h1o[(size_t)b * kH1 + lane + 32 * r] = h1[lane + 32 * r];
#pragma unroll
    for (int r = 0; r < 2; ++r) h2o[(size_t)b * kH2 + lane + 32 * r] = h2[lane + 32 * r];
    h3o[(size_t)b * kH3 + lane] = h3[lane];
    __syncwarp();
  }
  for (int e = 0; e < E; ++e) {
    const unsigned m = __ballot_sync(0xffffffffu, my_idx == e);
    if (lane == 0 && m) atomicAdd(&s_hist[e], __popc(m));
  }
  __syncthreads();
  if (tid < E) blk_hist[blockIdx.x * E + tid] = s_hist[tid];
}

// Single CTA: per expert, exclusive scan of the per-block histograms; totals -> counts, offsets, group tables.
__global__ void __launch_bounds__(512)
router_scan_kernel(const int32_t* __restrict__ blk_hist, int nblk, int E, int min_rows,
                   int32_t* __restrict__ counts, int32_t* __restrict__ offsets,
                   es_group* __restrict__ grp_half, es_group* __restrict__ grp_gen, int32_t* __restrict__ blk_base) {
  __shared__ int s_tot[kMaxE];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp < E) {
    const int e = warp;
    int running = 0;
    for (int c = 0; c < nblk; c += 32) {
      const int blk = c + lane;
      const int v = blk < nblk ? blk_hist[blk * E + e] : 0;
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (blk < nblk) blk_base[blk * E + e] = running + incl - v;
      running += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) s_tot[e] = running;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int off = 0;
    for (int e = 0; e < E; ++e) {
      const int c = s_tot[e];
      counts[e] = c;
      offsets[e] = off;
      const int act = c >= min_rows ? c : 0;
      if (grp_half) grp_half[e] = es_group{off, act, e, act};
      if (grp_gen) grp_gen[e] = es_group{2 * off, 2 * act, e, act};
      off += c;
    }
    offsets[E] = off;
  }
}

__global__ void __launch_bounds__(kRouterBlock)
router_scatter_kernel(const int64_t* __restrict__ idx, int B, int E, const int32_t* __restrict__ offsets,
                      const int32_t* __restrict__ blk_base, int32_t* __restrict__ perm) {
  __shared__ int s_wcnt[8][kMaxE];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 8 * kMaxE) (&s_wcnt[0][0])[tid] = 0;
  __syncthreads();
  const int b = blockIdx.x * kRouterBlock + tid;
  const int e = b < B ? (int)idx[b] : -1 - lane;   // invalid lanes get unique keys
  const unsigned peers = __match_any_sync(0xffffffffu, e);
  const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
  if (e >= 0 && rank_in_warp == 0) s_wcnt[warp][e] = __popc(peers);
  __syncthreads();
  if (e >= 0) {
    int rank = rank_in_warp;
    for (int w = 0; w < warp; ++w) rank += s_wcnt[w][e];
    perm[offsets[e] + blk_base[blockIdx.x * E + e] + rank] = b;
  }
}

__global__ void gather_rows_kernel(const float* __restrict__ in, const int32_t* __restrict__ perm, int width,
                                   float* __restrict__ out, int scatter) {
  const int j = blockIdx.x;
  const int src = scatter ? j : perm[j], dst = scatter ? perm[j] : j;
  const float* s = in + (size_t)src * width;
  float* d = out + (size_t)dst * width;
  if ((width & 3) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(s);
    float4* d4 = reinterpret_cast<float4*>(d);
    for (int c = threadIdx.x; c < width / 4; c += blockDim.x) d4[c] = s4[c];
  } else {
    for (int c = threadIdx.x; c < width; c += blockDim.x) d[c] = s[c];
  }
}

__global__ void gate_sums_kernel(const float* __restrict__ gates, int B, int E, float* __restrict__ sums) {
  __shared__ float red[32];
  float acc[kMaxE];
#pragma unroll
  for (int e = 0; e < kMaxE; ++e) acc[e] = 0.f;
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x)
    for (int e = 0; e < E; ++e) acc[e] += gates[(size_t)b * E + e];
  for (int e = 0; e < E; ++e) {
    const float t = block_sum(acc[e], red);
    if (threadIdx.x == 0) atomicAdd(&sums[e], t);
  }
}

// ---------------------------------------------------------------------------------------------------------
// router backward: dL/dgates[b,e] = c_e (+ extra[b,e]);  softmax/tau backward;  MLP backward with the weight
// gradients accumulated in shared memory per CTA and flushed once with global atomics.
// ---------------------------------------------------------------------------------------------------------
struct RouterBwdSmem {
  static constexpr int W2 = 0;                       // [64][128]
  static constexpr int W4 = W2 + kH2 * kH1;          // [32][64]
  static constexpr int W6 = W4 + kH3 * kH2;          // [16][32]
  static constexpr int DW0 = W6 + kMaxE * kH3;       // [128][9]
  static constexpr int DB0 = DW0 + kH1 * kCond;
  static constexpr int DW2 = DB0 + kH1;
  static constexpr int DB2 = DW2 + kH2 * kH1;
  static constexpr int DW4 = DB2 + kH2;
  static constexpr int DB4 = DW4 + kH3 * kH2;
  static constexpr int DW6 = DB4 + kH3;
  static constexpr int DB6 = DW6 + kMaxE * kH3;
  static constexpr int SCR = DB6 + kMaxE;            // per-warp [8][16 + 32 + 64 + 128]
  static constexpr int SCR_PER_WARP = 16 + kH3 + kH2 + kH1;
  static constexpr int TOTAL = SCR + 8 * SCR_PER_WARP;
};

__global__ void __launch_bounds__(256)
router_bwd_kernel(const float* __restrict__ cond, int B, int E, int B_global,
                  const float* __restrict__ w2, const float* __restrict__ w4, const float* __restrict__ w6,
                  const float* __restrict__ gates, const float* __restrict__ h1g, const float* __restrict__ h2g,
                  const float* __restrict__ h3g, const float* __restrict__ gate_sums, float tau,
                  float alb_strength, float alb_weight, float util_strength, const float* __restrict__ extra,
                  float* dw0, float* db0, float* dw2, float* db2, float* dw4, float* db4, float* dw6, float* db6,
                  float* __restrict__ losses_out) {
  extern __shared__ float sm[];
  __shared__ float s_c[kMaxE];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kH2 * kH1; i += 256) sm[RouterBwdSmem::W2 + i] = w2[i];
  for (int i = tid; i < kH3 * kH2; i += 256) sm[RouterBwdSmem::W4 + i] = w4[i];
  for (int i = tid; i < E * kH3; i += 256) sm[RouterBwdSmem::W6 + i] = w6[i];
  for (int i = RouterBwdSmem::DW0 + tid; i < RouterBwdSmem::SCR; i += 256) sm[i] = 0.f;
  if (tid < E) {
    // d/dS_e [ alb_w * alb_s * mean_e exp(1/(S_e+eps)) ]  +  d/dS_e [ util_s * sum_e p log(p+1e-9) ], p = S/B_global
    const float S = gate_sums[tid];
    const float inv = 1.f / (S + 1e-6f);
    float c = alb_weight * alb_strength / (float)E * expf(inv) * (-inv * inv);
    if (util_strength != 0.f) {
      const float p = S / (float)B_global;
      c += util_strength * (logf(p + 1e-9f) + p / (p + 1e-9f)) / (float)B_global;
    }
    s_c[tid] = c;
  }
  __syncthreads();
  if (blockIdx.x == 0 && tid == 0 && losses_out) {
    float alb = 0.f, ent = 0.f;
    for (int e = 0; e < E; ++e) {
      const float S = gate_sums[e];
      alb += expf(1.f / (S + 1e-6f));
      const float p = S / (float)B_global;
      ent += p * logf(p + 1e-9f);
    }
    losses_out[0] = alb_strength * alb / (float)E;
    losses_out[1] = util_strength * ent;   // = -util_strength * H(pbar)
  }
  float* scr = sm + RouterBwdSmem::SCR + warp * RouterBwdSmem::SCR_PER_WARP;
  float* d4 = scr;            // [16] dlogits
  float* d3 = scr + 16;       // [32]
  float* d2 = d3 + kH3;       // [64]
  float* d1 = d2 + kH2;       // [128]

  for (int b = blockIdx.x * 8 + warp; b < B; b += gridDim.x * 8) {
    // softmax backward
    float g = 0.f, dg = 0.f;
    if (lane < E) {
      g = gates[(size_t)b * E + lane];
      dg = s_c[lane] + (extra ? extra[(size_t)b * E + lane] : 0.f);
    }
    const float dot = warp_sum(g * dg);
    if (lane < E) d4[lane] = g * (dg - dot) / tau;
    __syncwarp();
    const float h3v = h3g[(size_t)b * kH3 + lane];
    {  // delta3 = W6^T d4, masked by LeakyReLU'
      float acc = 0.f;
      for (int e = 0; e < E; ++e) acc += sm[RouterBwdSmem::W6 + e * kH3 + lane] * d4[e];
      d3[lane] = acc * (h3v > 0.f ? 1.f : kLReLU);
    }
    // dW6[e][i] += d4[e] * h3[i]
    for (int e = 0; e < E; ++e) atomicAdd(&sm[RouterBwdSmem::DW6 + e * kH3 + lane], d4[e] * h3v);
    if (lane < E) atomicAdd(&sm[RouterBwdSmem::DB6 + lane], d4[lane]);
    __syncwarp();
    float h2v[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = lane + 32 * r;
      h2v[r] = h2g[(size_t)b * kH2 + i];
      float acc = 0.f;
#pragma unroll 8
      for (int o = 0; o < kH3; ++o) acc += sm[RouterBwdSmem::W4 + o * kH2 + i] * d3[o];
      d2[i] = acc * (h2v[r] > 0.f ? 1.f : kLReLU);
    }
    for (int o = 0; o < kH3; ++o) {
      const float d = d3[o];
#pragma unroll
      for (int r = 0; r < 2; ++r) atomicAdd(&sm[RouterBwdSmem::DW4 + o * kH2 + lane + 32 * r], d * h2v[r]);
    }
    atomicAdd(&sm[RouterBwdSmem::DB4 + lane], d3[lane]);
    __syncwarp();
    float h1v[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = lane + 32 * r;
      h1v[r] = h1g[(size_t)b * kH1 + i];
      float acc = 0.f;
#pragma unroll 8
      for (int o = 0; o < kH2; ++o) acc += sm[RouterBwdSmem::W2 + o * kH1 + i] * d2[o];
      d1[i] = acc * (h1v[r] > 0.f ? 1.f : kLReLU);
    }
    for (int o = 0; o < kH2; ++o) {
      const float d = d2[o];
#pragma unroll
      for (int r = 0; r < 4; ++r) atomicAdd(&sm[RouterBwdSmem::DW2 + o * kH1 + lane + 32 * r], d * h1v[r]);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) atomicAdd(&sm[RouterBwdSmem::DB2 + lane + 32 * r], d2[lane + 32 * r]);
    __syncwarp();
    // dW0[o][k] += d1[o] * x[k]
    const float xk = lane < kCond ? cond[(size_t)b * kCond + lane] : 0.f;
    for (int o = 0; o < kH1; ++o) {
      if (lane < kCond) atomicAdd(&sm[RouterBwdSmem::DW0 + o * kCond + lane], d1[o] * xk);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) atomicAdd(&sm[RouterBwdSmem::DB0 + lane + 32 * r], d1[lane + 32 * r]);
    __syncwarp();
  }
  __syncthreads();
  for (int i = tid; i < kH1 * kCond; i += 256) atomicAdd(&dw0[i], sm[RouterBwdSmem::DW0 + i]);
  for (int i = tid; i < kH1; i += 256) atomicAdd(&db0[i], sm[RouterBwdSmem::DB0 + i]);
  for (int i = tid; i < kH2 * kH1; i += 256) atomicAdd(&dw2[i], sm[RouterBwdSmem::DW2 + i]);
  for (int i = tid; i < kH2; i += 256) atomicAdd(&db2[i], sm[RouterBwdSmem::DB2 + i]);
  for (int i = tid; i < kH3 * kH2; i += 256) atomicAdd(&dw4[i], sm[RouterBwdSmem::DW4 + i]);
  for (int i = tid; i < kH3; i += 256) atomicAdd(&db4[i], sm[RouterBwdSmem::DB4 + i]);
  for (int i = tid; i < E * kH3; i += 256) atomicAdd(&dw6[i], sm[RouterBwdSmem::DW6 + i]);
  for (int i = tid; i < E; i += 256) atomicAdd(&db6[i], sm[RouterBwdSmem::DB6 + i]);
}

// expert-distribution loss: sum_{b,b'} (g_b . g_b') |m_b - m_b'| / B * 0.1 * strength with one-hot forward gates.
// d/dg_soft[b,e] = 2 * 0.1 * strength / B * sum_{b' : idx_b' == e} |m_b - m_b'|   (G G^T is symmetric)
__global__ void router_ed_kernel(const int64_t* __restrict__ idx, const float* __restrict__ m, int B, int E, float strength,
                                 float* __restrict__ dgates, float* __restrict__ loss_out) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const float mb = m[b];
  const int eb = (int)idx[b];
  float acc[kMaxE];
#pragma unroll
  for (int e = 0; e < kMaxE; ++e) acc[e] = 0.f;
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    const float d = fabsf(mb - m[j]);
    const int ej = (int)idx[j];
#pragma unroll
    for (int e = 0; e < kMaxE; ++e) acc[e] += (ej == e) ? d : 0.f;
  }
  const float k = 0.1f * strength / (float)B;
  for (int e = 0; e < E; ++e) {
    const float t = block_sum(acc[e], red);
    if (threadIdx.x == 0) {
      dgates[(size_t)b * E + e] = 2.f * k * t;
      if (e == eb) atomicAdd(loss_out, k * t);
    }
  }
}

}  // namespace es

using namespace es;

extern "C" int es_router_fwd(const float* cond, int B, int E, const float* w0, const float* b0, const float* w2,
                             const float* b2, const float* w4, const float* b4, const float* w6, const float* b6,
                             const float* gumbel, float tau, float* logits, float* gates, int64_t* idx, float* h1,
                             float* h2, float* h3, int32_t* blk_hist, void* stream) {
  ES_REQUIRE(B > 0 && E >= 1 && E <= kMaxE, "need B>0 and 1<=E<=16");
  ES_REQUIRE(cond && gumbel && logits && gates && idx && h1 && h2 && h3 && blk_hist, "null pointer");
  ES_REQUIRE(tau > 0.f, "tau must be positive");
  const size_t smem = RouterSmem::TOTAL * sizeof(float);
  ES_CUDA(cudaFuncSetAttribute(router_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  router_fwd_kernel<<<ceil_div(B, kRouterBlock), kRouterBlock, smem, as_stream(stream)>>>(
      cond, B, E, w0, b0, w2, b2, w4, b4, w6, b6, gumbel, tau, logits, gates, idx, h1, h2, h3, blk_hist);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_router_partition(const int64_t* idx, int B, int E, int min_rows, const int32_t* blk_hist,
                                   int32_t* counts, int32_t* offsets, int32_t* perm, es_group* grp_half,
                                   es_group* grp_gen, int32_t* scratch, void* stream) {
  ES_REQUIRE(B > 0 && E >= 1 && E <= kMaxE, "need B>0 and 1<=E<=16");
  ES_REQUIRE(idx && blk_hist && counts && offsets && perm && scratch, "null pointer");
  const int nblk = ceil_div(B, kRouterBlock);
  router_scan_kernel<<<1, 512, 0, as_stream(stream)>>>(blk_hist, nblk, E, min_rows, counts, offsets, grp_half, grp_gen, scratch);
  ES_LAUNCH_CHECK();
  router_scatter_kernel<<<nblk, kRouterBlock, 0, as_stream(stream)>>>(idx, B, E, offsets, scratch, perm);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_gather_rows(const float* in, const int32_t* perm, int B, int width, float* out, void* stream) {
  ES_REQUIRE(in && perm && out && B > 0 && width > 0, "bad arguments");
  gather_rows_kernel<<<B, width >= 512 ? 128 : 32, 0, as_stream(stream)>>>(in, perm, width, out, 0);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_scatter_rows(const float* in, const int32_t* perm, int B, int width, float* out, void* stream) {
  ES_REQUIRE(in && perm && out && B > 0 && width > 0, "bad arguments");
  gather_rows_kernel<<<B, width >= 512 ? 128 : 32, 0, as_stream(stream)>>>(in, perm, width, out, 1);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_router_gate_sums(const float* gates, int B, int E, float* gate_sums, void* stream) {
  ES_REQUIRE(gates && gate_sums && B > 0 && E >= 1 && E <= kMaxE, "bad arguments");
  ES_CUDA(cudaMemsetAsync(gate_sums, 0, E * sizeof(float), as_stream(stream)));
  const int blocks = min(64, ceil_div(B, 256));
  gate_sums_kernel<<<blocks, 256, 0, as_stream(stream)>>>(gates, B, E, gate_sums);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_router_bwd(const float* cond, int B, int E, int B_global, const float* w2, const float* w4,
                             const float* w6, const float* gates, const float* h1, const float* h2, const float* h3,
                             const float* gate_sums, float tau, float alb_strength, float alb_weight,
                             float util_strength, const float* extra_dgates, float* dw0, float* db0, float* dw2,
                             float* db2, float* dw4, float* db4, float* dw6, float* db6, float* losses_out,
                             void* stream) {
  ES_REQUIRE(B > 0 && E >= 1 && E <= kMaxE && B_global >= B, "bad sizes");
  ES_REQUIRE(cond && w2 && w4 && w6 && gates && h1 && h2 && h3 && gate_sums, "null input");
  ES_REQUIRE(dw0 && db0 && dw2 && db2 && dw4 && db4 && dw6 && db6, "null gradient buffer");
  const size_t smem = RouterBwdSmem::TOTAL * sizeof(float);
  ES_CUDA(cudaFuncSetAttribute(router_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int blocks = min(64, ceil_div(B, 8));
  router_bwd_kernel<<<blocks, 256, smem, as_stream(stream)>>>(cond, B, E, B_global, w2, w4, w6, gates, h1, h2, h3,
                                                              gate_sums, tau, alb_strength, alb_weight, util_strength,
                                                              extra_dgates, dw0, db0, dw2, db2, dw4, db4, dw6, db6,
                                                              losses_out);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_router_ed_loss(const int64_t* idx, const float* m, int B, int E, float ed_strength, float* dgates,
                                 float* loss_out, void* stream) {
  ES_REQUIRE(idx && m && dgates && loss_out && B > 0 && E >= 1 && E <= kMaxE, "bad arguments");
  ES_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), as_stream(stream)));
  router_ed_kernel<<<B, 128, 0, as_stream(stream)>>>(idx, m, B, E, ed_strength, dgates, loss_out);
  ES_LAUNCH_CHECK();
  return ES_OK;
}
