// K2d — GroupNorm(32) + LeakyReLU of the proton generator (bf16 NHWC), forward and backward, as thread-block CLUSTERS.
//
// Reference: nn.GroupNorm(32, C) + nn.LeakyReLU(0.1) after every generator conv (expertsim/models/proton/generator.py:18-44).
//
// A sample's activation (340-408 KB) does not fit one SM's shared memory, so the one-CTA-per-sample kernel
// (gn_lrelu_kernel, gen_misc.cu) re-read it from HBM for every pass (mean, variance, apply: 3 reads + 1 write; backward
// 2 x (x, dy) + 1 write) with 16 bytes in flight per thread.  Here CL CTAs of one cluster split the sample's pixels, each
// keeps its slab in shared memory (cp.async, everything in flight at once), the 32 group statistics are exchanged through
// distributed shared memory, and every element crosses HBM exactly once per direction: forward 1 read + 1 write,
// backward 2 reads + 1 write.  Statistics stay two-pass (mean, then centred sum of squares) — free from shared memory.
// The per-channel affine / conv-bias gradients are reduced over the cluster first: one global atomic per channel per
// SAMPLE, as before.
#include <cooperative_groups.h>

#include "common.cuh"
#include "gen_common.cuh"

namespace cg = cooperative_groups;

namespace es {
namespace {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

template <bool BWD, bool FAN>
__global__ void __launch_bounds__(256)
gn_cluster_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy_up, int Hs, int Ws, int Hu,
                  int Wu, const float* __restrict__ gamma, const float* __restrict__ beta, long slot_stride, int C,
                  int groups, const es_group* __restrict__ grp, int n_groups, __nv_bfloat16* __restrict__ out,
                  float* __restrict__ stats, float* __restrict__ dgamma, float* __restrict__ dbeta,
                  float* __restrict__ dbias, int CL, int Pq, int OWs, int OWu) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ float s_c[4][256];        // per-channel accumulators of this CTA
  __shared__ float s_part[2][64];      // per-group partial sums, read by the other CTAs of the cluster
  __shared__ float s_g1[64], s_g2[64];
  __shared__ int ylo[64], yhi[64], xlo[64], xhi[64];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int r = blockIdx.x / CL;
  const int P = Hs * Ws, c4 = C / 8, cpg = C / groups;
  const int pb = rank * Pq, np = max(0, min(P, pb + Pq) - pb);
  const int tid = threadIdx.x;
  const int g = find_group(grp, n_groups, r);      // the same for every CTA of the cluster
  if (g < 0) {
    // rows of skipped experts: the gradient tensor is read by the weight-gradient GEMM through TMA boxes that may straddle
    // a group's end, where it meets zero-filled im2col rows — it must be finite there
    // (forward: the strip kernels of the following conv's weight gradient read x strips past a group's end against zero dy)
    if (BWD || OWu <= 0) {
      uint4* o4 = reinterpret_cast<uint4*>(out + ((size_t)r * P + pb) * C);
      for (int i = tid; i < np * c4; i += 256) o4[i] = make_uint4(0, 0, 0, 0);
    } else if (rank == 0) {
      const int n16 = (P / OWs) * OWu * c4;
      uint4* o4 = reinterpret_cast<uint4*>(out + (size_t)r * n16 * 8);
      for (int i = tid; i < n16; i += 256) o4[i] = make_uint4(0, 0, 0, 0);
    }
    return;
  }
  const int slot = grp[g].slot;
  const int cu = tid % c4, c8 = cu * 8, pstep = 256 / c4, pl0 = tid / c4;
  const float cnt = (float)(cpg * P);
  uint4* s_x = reinterpret_cast<uint4*>(smraw);                    // [Pq][c4] 8 x bf16
  uint4* s_dy = s_x + (size_t)Pq * c4;                             // backward, direct gradient: [Pq][c4] 8 x bf16
  float4* s_da = reinterpret_cast<float4*>(s_x + (size_t)Pq * c4); // backward, fan-in summed gradient: [Pq][c4] 8 x fp32
  {
    const uint4* x4 = reinterpret_cast<const uint4*>(x + ((size_t)r * P + pb) * C);
    for (int i = tid; i < np * c4; i += 256) cp_async16(s_x + i, x4 + i);
    if (BWD && !FAN) {
      const uint4* d4 = reinterpret_cast<const uint4*>(dy_up + ((size_t)r * P + pb) * C);
      for (int i = tid; i < np * c4; i += 256) cp_async16(s_dy + i, d4 + i);
    }
  }
  if (BWD && FAN && tid == 0) { build_fanin(Hs, Hu, ylo, yhi); build_fanin(Ws, Wu, xlo, xhi); }
  if (!BWD && OWu > 0 && tid == 0) build_fanin(OWs, OWu, xlo, xhi);     // forward with the x-upsampled output layout
  s_c[0][tid] = 0.f; s_c[1][tid] = 0.f; s_c[2][tid] = 0.f; s_c[3][tid] = 0.f;
  __syncthreads();
  if (BWD && FAN) {
    const __nv_bfloat16* dyr = dy_up + (size_t)r * Hu * Wu * C;
    for (int pl = pl0; pl < np; pl += pstep) {
      const int pix = pb + pl;
      float da[8];
      load_da8(dyr, Wu, C, c8, ylo, yhi, xlo, xhi, pix / Ws, pix % Ws, da);
      s_da[((size_t)pl * c4 + cu) * 2] = make_float4(da[0], da[1], da[2], da[3]);
      s_da[((size_t)pl * c4 + cu) * 2 + 1] = make_float4(da[4], da[5], da[6], da[7]);
    }
  }
  cp_async_wait_all();
  __syncthreads();
  float gk[8], bk[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { gk[k] = gamma[slot * slot_stride + c8 + k]; bk[k] = beta[slot * slot_stride + c8 + k]; }
  float f[8];
  if (!BWD) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int pl = pl0; pl < np; pl += pstep) {
      unpack8(s_x[(size_t)pl * c4 + cu], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += f[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&s_c[0][c8 + k], acc[k]);
    __syncthreads();
    if (tid < groups) {
      float s = 0.f;
      for (int k = 0; k < cpg; ++k) s += s_c[0][tid * cpg + k];
      s_part[0][tid] = s;
    }
    cluster.sync();
    if (tid < groups) {
      float t = 0.f;
      for (int rk = 0; rk < CL; ++rk) t += *cluster.map_shared_rank(&s_part[0][tid], rk);
      s_g1[tid] = t / cnt;
    }
    __syncthreads();
    float mu[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { mu[k] = s_g1[(c8 + k) / cpg]; acc[k] = 0.f; }
    for (int pl = pl0; pl < np; pl += pstep) {
      unpack8(s_x[(size_t)pl * c4 + cu], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += (f[k] - mu[k]) * (f[k] - mu[k]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&s_c[1][c8 + k], acc[k]);
    __syncthreads();
    if (tid < groups) {
      float s = 0.f;
      for (int k = 0; k < cpg; ++k) s += s_c[1][tid * cpg + k];
      s_part[1][tid] = s;
    }
    cluster.sync();
    if (tid < groups) {
      float t = 0.f;
      for (int rk = 0; rk < CL; ++rk) t += *cluster.map_shared_rank(&s_part[1][tid], rk);
      const float rstd = rsqrtf(t / cnt + kNormEps);
      s_g2[tid] = rstd;
      if (rank == 0) {
        stats[((size_t)r * groups + tid) * 2] = s_g1[tid];
        stats[((size_t)r * groups + tid) * 2 + 1] = rstd;
      }
    }
    __syncthreads();
    float rs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) rs[k] = s_g2[(c8 + k) / cpg];
    uint4* y4 = reinterpret_cast<uint4*>(out + ((size_t)r * P + pb) * C);
    uint4* yu4 = reinterpret_cast<uint4*>(out + (size_t)r * (OWu > 0 ? (P / OWs) * OWu : 0) * C);   // x-upsampled [P/OWs, OWu, C]
    for (int pl = pl0; pl < np; pl += pstep) {
      unpack8(s_x[(size_t)pl * c4 + cu], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = lrelu((f[k] - mu[k]) * rs[k] * gk[k] + bk[k]);
      if (OWu > 0) {
        // the activation is stored nearest-upsampled along x (what the following conv's upsample would read): the conv then
        // reads its source directly and both of its operands can come by TMA
        const int pix = pb + pl, h = pix / OWs, w = pix - h * OWs;
        const uint4 v = pack8(f);
        for (int xu = xlo[w]; xu < xhi[w]; ++xu) yu4[((size_t)h * OWu + xu) * c4 + cu] = v;
      } else {
        y4[(size_t)pl * c4 + cu] = pack8(f);
      }
    }
    cluster.sync();     // the other CTAs may still be reading this CTA's partial sums
  } else {
    float mu[8], rs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int gi = (c8 + k) / cpg;
      mu[k] = stats[((size_t)r * groups + gi) * 2];
      rs[k] = stats[((size_t)r * groups + gi) * 2 + 1];
    }
    float a1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, a2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float ag[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ab[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float da[8];
    auto load_grad = [&](int pl) {
      if (FAN) {
        const float4 u = s_da[((size_t)pl * c4 + cu) * 2], v = s_da[((size_t)pl * c4 + cu) * 2 + 1];
        da[0] = u.x; da[1] = u.y; da[2] = u.z; da[3] = u.w; da[4] = v.x; da[5] = v.y; da[6] = v.z; da[7] = v.w;
      } else {
        unpack8(s_dy[(size_t)pl * c4 + cu], da);
      }
    };
    for (int pl = pl0; pl < np; pl += pstep) {
      load_grad(pl);
      unpack8(s_x[(size_t)pl * c4 + cu], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xh = (f[k] - mu[k]) * rs[k];
        const float yv = xh * gk[k] + bk[k];
        const float d = da[k] * (yv > 0.f ? 1.f : kLReLU);
        ag[k] += d * xh;
        ab[k] += d;
      }
    }
    // gamma is constant per channel: sum(d * gamma) = gamma * sum(d)
#pragma unroll
    for (int k = 0; k < 8; ++k) { a1[k] = gk[k] * ab[k]; a2[k] = gk[k] * ag[k]; }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      atomicAdd(&s_c[0][c8 + k], a1[k]); atomicAdd(&s_c[1][c8 + k], a2[k]);
      atomicAdd(&s_c[2][c8 + k], ag[k]); atomicAdd(&s_c[3][c8 + k], ab[k]);
    }
    __syncthreads();
    if (tid < groups) {
      float s1 = 0.f, s2 = 0.f;
      for (int k = 0; k < cpg; ++k) { s1 += s_c[0][tid * cpg + k]; s2 += s_c[1][tid * cpg + k]; }
      s_part[0][tid] = s1;
      s_part[1][tid] = s2;
    }
    cluster.sync();
    if (tid < groups) {
      float t1 = 0.f, t2 = 0.f;
      for (int rk = 0; rk < CL; ++rk) {
        t1 += *cluster.map_shared_rank(&s_part[0][tid], rk);
        t2 += *cluster.map_shared_rank(&s_part[1][tid], rk);
      }
      s_g1[tid] = t1 / cnt;
      s_g2[tid] = t2 / cnt;
    }
    if (tid < C) s_c[0][tid] = 0.f;      // free again (only read locally, above): conv-bias gradient accumulator
    __syncthreads();
    float m1[8], m2[8], al[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 8; ++k) { m1[k] = s_g1[(c8 + k) / cpg]; m2[k] = s_g2[(c8 + k) / cpg]; }
    uint4* dx4 = reinterpret_cast<uint4*>(out + ((size_t)r * P + pb) * C);
    for (int pl = pl0; pl < np; pl += pstep) {
      load_grad(pl);
      unpack8(s_x[(size_t)pl * c4 + cu], f);
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xh = (f[k] - mu[k]) * rs[k];
        const float yv = xh * gk[k] + bk[k];
        const float d = da[k] * (yv > 0.f ? 1.f : kLReLU) * gk[k];
        o[k] = rs[k] * (d - m1[k] - xh * m2[k]);
        al[k] += o[k];
      }
      dx4[(size_t)pl * c4 + cu] = pack8(o);
    }
    if (dbias) {
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(&s_c[0][c8 + k], al[k]);
    }
    cluster.sync();
    if (rank == 0 && tid < C) {
      float tg = 0.f, tb = 0.f, tl = 0.f;
      for (int rk = 0; rk < CL; ++rk) {
        tg += *cluster.map_shared_rank(&s_c[2][tid], rk);
        tb += *cluster.map_shared_rank(&s_c[3][tid], rk);
        if (dbias) tl += *cluster.map_shared_rank(&s_c[0][tid], rk);
      }
      atomicAdd(&dgamma[slot * slot_stride + tid], tg);
      atomicAdd(&dbeta[slot * slot_stride + tid], tb);
      if (dbias) atomicAdd(&dbias[slot * slot_stride + tid], tl);
    }
    cluster.sync();     // rank 0 may still be reading this CTA's accumulators
  }
}

}  // namespace

// Picks the cluster size (bytes per pixel `bpp` of shared memory): the smallest power of two <= 8 whose slab lets two CTAs
// share an SM, else the smallest that fits at all.  Returns 0 when even a cluster of 8 does not fit.
static int pick_cluster(int P, int bpp, int* Pq_out) {
  static const size_t slab = [] { const char* e = getenv("ES_GN_SLAB_KB"); return e ? (size_t)atoi(e) * 1024 : (size_t)104000; }();
  int fit = 0;
  for (int cl = 1; cl <= 8; cl *= 2) {
    const int pq = ceil_div(P, cl);
    const size_t bytes = (size_t)pq * bpp;
    if (bytes <= slab) { *Pq_out = pq; return cl; }
    if (!fit && bytes <= 200 * 1024) fit = cl;
  }
  if (fit) *Pq_out = ceil_div(P, fit);
  return fit;
}

template <bool BWD, bool FAN>
static int launch_gn_cluster(const void* x, const void* dy_up, int Hs, int Ws, int Hu, int Wu, const float* gamma,
                             const float* beta, long slot_stride, int C, int groups, const es_group* grp, int n_groups,
                             int total_rows, void* out, float* stats, float* dgamma, float* dbeta, float* dbias,
                             void* stream, int OWs = 0, int OWu = 0) {
  int Pq = 0;
  const int bpp = (BWD ? (FAN ? 6 : 4) : 2) * C;
  const int CL = pick_cluster(Hs * Ws, bpp, &Pq);
  if (CL == 0) return 1;     // caller falls back to the one-CTA-per-sample kernel
  const size_t smem = (size_t)Pq * bpp;
  auto kern = gn_cluster_kernel<BWD, FAN>;
  ES_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  static const int carve = [] { const char* e = getenv("ES_GN_CARVEOUT"); return e ? atoi(e) : 100; }();
  if (carve >= 0) ES_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)total_rows * CL);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  ES_CUDA(cudaLaunchKernelEx(&cfg, kern, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy_up, Hs, Ws, Hu, Wu, gamma, beta,
                             slot_stride, C, groups, grp, n_groups, (__nv_bfloat16*)out, stats, dgamma, dbeta, dbias, CL, Pq, OWs, OWu));
  return ES_OK;
}

int gn_cluster_fwd(const void* x, const float* gamma, const float* beta, long slot_stride, int P, int C, int groups,
                   const es_group* grp, int n_groups, int total_rows, void* y, float* stats, void* stream, int OWs, int OWu) {
  return launch_gn_cluster<false, false>(x, nullptr, P, 1, P, 1, gamma, beta, slot_stride, C, groups, grp, n_groups,
                                         total_rows, y, stats, nullptr, nullptr, nullptr, stream, OWs, OWu);
}

int gn_cluster_bwd(const void* dy_up, int Hs, int Ws, int Hu, int Wu, const void* x, const float* stats,
                   const float* gamma, const float* beta, long slot_stride, int C, int groups, const es_group* grp,
                   int n_groups, int total_rows, void* dx, float* dgamma, float* dbeta, float* dbias, void* stream) {
  if (Hu == Hs && Wu == Ws)
    return launch_gn_cluster<true, false>(x, dy_up, Hs, Ws, Hu, Wu, gamma, beta, slot_stride, C, groups, grp, n_groups,
                                          total_rows, dx, const_cast<float*>(stats), dgamma, dbeta, dbias, stream);
  return launch_gn_cluster<true, true>(x, dy_up, Hs, Ws, Hu, Wu, gamma, beta, slot_stride, C, groups, grp, n_groups,
                                       total_rows, dx, const_cast<float*>(stats), dgamma, dbeta, dbias, stream);
}

}  // namespace es
