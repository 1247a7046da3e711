// BatchNorm (+ Dropout + LeakyReLU) of the neutron networks, forward and backward, grouped by expert.
// Reference: GeneratorNeutron (expertsim/models/neutron/generator.py:11-40: Linear/Conv -> BatchNorm -> Dropout(0.2) ->
// LeakyReLU) and AuxRegNeutron's FeatureExtractor (expertsim/models/neutron/aux_reg.py:11-49: Conv -> BatchNorm2d ->
// LeakyReLU -> Dropout(0.2)).  torch semantics: training normalises with the biased batch variance and feeds the unbiased
// one into running_var (momentum 0.1, eps 1e-5); eval uses the running statistics.
//
// A "stat group" is what one reference forward call normalises over: (expert slot, pass).  The generator's training batch
// holds G(z1) and G(z2) of an expert back to back (two passes = two reference calls = two sets of batch statistics and
// two running-stat updates); sg = 2*slot + pass.  Per stat group and channel the kernels exchange only (sum, sum of
// squares) / (sum g, sum g*xhat) in fp64 — exactly what data parallelism has to all-reduce (SyncBN), done by the caller
// between the reduce and the apply kernels.
//
// Generator tensors are bf16 NHWC [rows, P, C]; the aux regressor's are fp32 NCHW [rows, C, P].  `chmap` (nullable)
// maps a stat channel to the reference's parameter index (fc2's channels-last feature order -> NCHW flattening).
// Dropout keep decisions come from an injected mask (parity runs; laid out like the reference's NCHW activation,
// index (row*C + c)*P + p) or from a counter-based hash of (seed, element index) that backward re-evaluates.
#include "common.cuh"
#include "gen_common.cuh"

namespace es {

namespace {

__device__ __forceinline__ uint32_t mix32(uint64_t x) {   // splitmix64 finaliser
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (uint32_t)(x >> 32);
}
// keep-scale of one element: 0 (dropped) or 1/(1-p)
__device__ __forceinline__ float keep_scale(const float* __restrict__ mask, uint64_t seed, size_t idx, float p) {
  if (p <= 0.f) return 1.f;
  bool keep;
  if (mask) keep = mask[idx] != 0.f;
  else keep = (float)(mix32(seed + idx * 0x9E3779B97F4A7C15ull) >> 8) * (1.f / 16777216.f) >= p;
  return keep ? 1.f / (1.f - p) : 0.f;
}

struct SgRow { int g, slot, pass, sg; };
__device__ __forceinline__ SgRow sg_of_row(const es_group* grp, int E, int r, int two_pass) {
  SgRow o{-1, 0, 0, 0};
  o.g = find_group(grp, E, r);
  if (o.g < 0) return o;
  const es_group G = grp[o.g];
  o.slot = G.slot;
  o.pass = (two_pass && (r - G.row_start) >= G.pass_rows) ? 1 : 0;
  o.sg = 2 * G.slot + o.pass;
  return o;
}

// ------------------------------------------------------------------------------------------------ NHWC bf16 (generator)
// Iteration geometry = statistics geometry: a row is [P_it pixels][C_it stat channels].  BatchNorm2d: P_it = Hs*Ws,
// C_it = C.  BatchNorm1d over a flattened map (fc2): P_it = 1, C_it = Hs*Ws*C — the flat NHWC offset is the same, only
// the statistics are per feature.  The spatial view (Hs, Ws, Csp) is what the upsample fan-in and the reference's NCHW
// dropout-mask index need.  CTA = chunk of `rpc` rows x one block of 8*OPB stat channels; threads = OPB octets x
// (256/OPB) pixel lanes; partial sums stay in registers across the rows of a chunk and are flushed per stat group.
template <bool BWD>
__global__ void __launch_bounds__(256)
bn_reduce_nhwc_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy_up, int Hs, int Ws, int Hu,
                      int Wu, int Csp, int P_it, int C_it, int OPB, int rpc, const float* __restrict__ stats,
                      const float* __restrict__ gamma, const float* __restrict__ beta, long slot_stride,
                      const int* __restrict__ chmap, const float* __restrict__ mask, unsigned long long seed, float p_drop,
                      const es_group* __restrict__ grp, int E, int total_rows, int two_pass, double* __restrict__ sums) {
  __shared__ float s_a[256], s_b[256];
  __shared__ int ylo[64], yhi[64], xlo[64], xhi[64];
  const int P = Hs * Ws, c4 = C_it / 8;
  const int tid = threadIdx.x, ol = tid % OPB, pl = tid / OPB, lanes = 256 / OPB;
  const int oct = blockIdx.y * OPB + ol, c8 = oct * 8;
  if (BWD && tid == 0) { build_fanin(Hs, Hu, ylo, yhi); build_fanin(Ws, Wu, xlo, xhi); }
  s_a[tid] = 0.f; s_b[tid] = 0.f;
  __syncthreads();
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float f[8], da[8];
  int cur = -1;
  auto flush = [&]() {      // called by ALL threads of the CTA (cur is uniform)
#pragma unroll
    for (int k = 0; k < 8; ++k) { atomicAdd(&s_a[ol * 8 + k], a[k]); atomicAdd(&s_b[ol * 8 + k], b[k]); a[k] = 0.f; b[k] = 0.f; }
    __syncthreads();
    if (tid < OPB * 8) {
      double* d = sums + ((size_t)cur * C_it + blockIdx.y * OPB * 8 + tid) * 2;
      atomicAdd(d, (double)s_a[tid]);
      atomicAdd(d + 1, (double)s_b[tid]);
      s_a[tid] = 0.f; s_b[tid] = 0.f;
    }
    __syncthreads();
  };
  // one (row, pixel) of this thread's octet: accumulate into a[], b[]
  auto visit = [&](int r, int pix, int sg, int slot) {
    const uint4* x4 = reinterpret_cast<const uint4*>(x + (size_t)r * P * Csp);
    const __nv_bfloat16* dyr = BWD ? dy_up + (size_t)r * Hu * Wu * Csp : nullptr;
    const int flat = pix * C_it + c8;                 // NHWC offset inside the row
    unpack8(__ldg(x4 + flat / 8), f);
    if (!BWD) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { a[k] += f[k]; b[k] += f[k] * f[k]; }
    } else {
      const int spix = flat / Csp, sc8 = flat - spix * Csp;
      load_da8(dyr, Wu, Csp, sc8, ylo, yhi, xlo, xhi, spix / Ws, spix % Ws, da);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int sc = c8 + k;
        const int pc = chmap ? chmap[sc] : sc;
        const float mu = stats[((size_t)sg * C_it + sc) * 2], rs = stats[((size_t)sg * C_it + sc) * 2 + 1];
        const float xh = (f[k] - mu) * rs;
        const float ks = keep_scale(mask, seed, ((size_t)r * Csp + sc8 + k) * P + spix, p_drop);
        const float pre = (xh * gamma[slot * slot_stride + pc] + beta[slot * slot_stride + pc]) * ks;   // dropout, then LeakyReLU
        const float g2 = da[k] * (pre > 0.f ? 1.f : kLReLU) * ks;
        a[k] += g2;
        b[k] += g2 * xh;
      }
    }
  };
  if (P_it == 1) {
    // BatchNorm1d over a flattened map (one "pixel" per row): the pixel lanes would idle (r01 launch list: ONE launch of
    // 2.98 ms, 16 of 256 threads active), so the lanes walk ROWS instead.  blockIdx.x enumerates chunks of `rpc` rows that
    // lie inside ONE statistics group (expert, pass) — decoded from the device-side group table — which keeps the flush
    // uniform: one per CTA.
    int xq = blockIdx.x, sg = -1, slot = 0, rlo = 0, rhi = 0;
    for (int gi = 0; gi < E && sg < 0; ++gi) {
      const es_group G = grp[gi];
      if (G.rows == 0) continue;
      for (int ps = 0; ps < (two_pass ? 2 : 1); ++ps) {
        const int lo = G.row_start + (ps ? G.pass_rows : 0);
        const int n = two_pass ? (ps ? G.rows - G.pass_rows : G.pass_rows) : G.rows;
        const int nc = ceil_div(n, rpc);
        if (xq < nc) { sg = 2 * G.slot + ps; slot = G.slot; rlo = lo + xq * rpc; rhi = min(lo + n, rlo + rpc); break; }
        xq -= nc;
      }
    }
    if (sg < 0) return;       // uniform over the CTA
    for (int r = rlo + pl; r < rhi; r += lanes) visit(r, 0, sg, slot);
    cur = sg;
    flush();
    return;
  }
  const int r0 = blockIdx.x * rpc, r1 = min(total_rows, r0 + rpc);
  for (int r = r0; r < r1; ++r) {
    const SgRow q = sg_of_row(grp, E, r, two_pass);
    const int sg = q.g < 0 ? -1 : q.sg;
    if (sg != cur) {
      if (cur >= 0) flush();
      cur = sg;
    }
    if (sg < 0) continue;
    for (int pix = pl; pix < P_it; pix += lanes) visit(r, pix, sg, q.slot);
  }
  if (cur >= 0) flush();
}

// forward apply: y = lrelu(dropout(bn(x)));  backward apply: dx = gamma*rstd*(g2 - mean(g2) - xhat*mean(g2*xhat))
template <bool BWD>
__global__ void __launch_bounds__(256)
bn_apply_nhwc_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy_up, int Hs, int Ws, int Hu,
                     int Wu, int Csp, int C_it, const float* __restrict__ stats, const double* __restrict__ sums2,
                     const float* __restrict__ n_sg, const float* __restrict__ gamma, const float* __restrict__ beta,
                     long slot_stride, const int* __restrict__ chmap, const float* __restrict__ mask,
                     unsigned long long seed, float p_drop, const es_group* __restrict__ grp, int E, int two_pass,
                     __nv_bfloat16* __restrict__ out) {
  __shared__ int ylo[64], yhi[64], xlo[64], xhi[64];
  const int r = blockIdx.x;
  const SgRow q = sg_of_row(grp, E, r, two_pass);
  const int P = Hs * Ws, n4 = P * Csp / 8;
  uint4* o4 = reinterpret_cast<uint4*>(out + (size_t)r * P * Csp);
  if (q.g < 0) {
    for (int i = threadIdx.x; i < n4; i += blockDim.x) o4[i] = make_uint4(0, 0, 0, 0);   // finite rows for skipped experts: see gn_lrelu_kernel
    return;
  }
  if (BWD && threadIdx.x == 0) { build_fanin(Hs, Hu, ylo, yhi); build_fanin(Ws, Wu, xlo, xhi); }
  __syncthreads();
  const uint4* x4 = reinterpret_cast<const uint4*>(x + (size_t)r * P * Csp);
  const __nv_bfloat16* dyr = BWD ? dy_up + (size_t)r * Hu * Wu * Csp : nullptr;
  const float inv_n = BWD ? 1.f / n_sg[q.sg] : 0.f;
  float f[8], da[8], o[8];
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    const int flat = i * 8;
    const int spix = flat / Csp, sc8 = flat - spix * Csp;
    const int st8 = flat % C_it;                      // stat channel of the octet's first element
    unpack8(__ldg(x4 + i), f);
    if (BWD) load_da8(dyr, Wu, Csp, sc8, ylo, yhi, xlo, xhi, spix / Ws, spix % Ws, da);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int sc = st8 + k;
      const int pc = chmap ? chmap[sc] : sc;
      const size_t si = ((size_t)q.sg * C_it + sc) * 2;
      const float mu = stats[si], rs = stats[si + 1];
      const float ga = gamma[q.slot * slot_stride + pc], be = beta[q.slot * slot_stride + pc];
      const float xh = (f[k] - mu) * rs;
      const float ks = keep_scale(mask, seed, ((size_t)r * Csp + sc8 + k) * P + spix, p_drop);
      const float pre = (xh * ga + be) * ks;
      if (!BWD) {
        o[k] = lrelu(pre);
      } else {
        const float g2 = da[k] * (pre > 0.f ? 1.f : kLReLU) * ks;
        const float m1 = (float)sums2[si] * inv_n, m2 = (float)sums2[si + 1] * inv_n;
        o[k] = ga * rs * (g2 - m1 - xh * m2);
      }
    }
    o4[i] = pack8(o);
  }
}

// ------------------------------------------------------------------------------------------------ NHWC bf16, BatchNorm2d fast path
// The generic kernels above re-load the channel's statistics / affine parameters and hash one dropout decision per ELEMENT
// (r01b launch list, neutron: 25 ms of a 58 ms step at 0.2-0.6 TB/s).  For BatchNorm2d with C <= 256 a thread owns one
// channel octet for the whole row: the per-channel constants live in registers, U pixels are fetched before any is
// consumed (see gn_lrelu_kernel), and the keep decisions of an octet come from two 64-bit hashes (16 bits per element)
// instead of eight.  The hashed dropout pattern only has to agree between the forward and the two backward kernels of
// the same step; injected masks (parity runs) keep the reference's NCHW index.
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}
__device__ __forceinline__ void keep8(const float* __restrict__ mask, uint64_t seed, size_t oct, size_t nchw_base, int P,
                                      float p, float* ks) {
  if (p <= 0.f) {
#pragma unroll
    for (int k = 0; k < 8; ++k) ks[k] = 1.f;
    return;
  }
  const float inv = 1.f / (1.f - p);
  if (mask) {
#pragma unroll
    for (int k = 0; k < 8; ++k) ks[k] = mask[nchw_base + (size_t)k * P] != 0.f ? inv : 0.f;
    return;
  }
  const uint64_t h1 = mix64(seed + oct * 0x9E3779B97F4A7C15ull), h2 = mix64(h1 + 0xD1B54A32D192ED03ull);
  const uint32_t thr = (uint32_t)(p * 65536.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    ks[k] = ((uint32_t)(h1 >> (16 * k)) & 0xFFFFu) >= thr ? inv : 0.f;
    ks[4 + k] = ((uint32_t)(h2 >> (16 * k)) & 0xFFFFu) >= thr ? inv : 0.f;
  }
}

// U pixels of one channel octet: x (always) and the gradient (direct, or summed over the <= 2 x 2 upsample fan-out)
template <bool BWD, bool FAN, int U>
struct BnFetch {
  uint4 qx[U];
  uint4 qd[U][FAN ? 4 : 1];
  int nq[U];
  __device__ __forceinline__ void load(const uint4* __restrict__ x4, const uint4* __restrict__ d4, int pix0, int pstep, int P,
                                       int c4, int cu, int Ws, int Wu, const int* ylo, const int* yhi, const int* xlo,
                                       const int* xhi) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = pix0 + u * pstep;
      nq[u] = 0;
      if (pix < P) {
        qx[u] = __ldg(x4 + (size_t)pix * c4 + cu);
        if (BWD) {
          if (FAN) {
            const int sy = pix / Ws, sx = pix - sy * Ws;
            const int y0 = ylo[sy], ny = yhi[sy] - y0, x0 = xlo[sx], nx = xhi[sx] - x0;
            nq[u] = ny * nx;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (j < ny * nx) qd[u][j] = __ldg(d4 + ((size_t)(y0 + j / nx) * Wu + x0 + j % nx) * c4 + cu);
          } else {
            qd[u][0] = __ldg(d4 + (size_t)pix * c4 + cu);
          }
        }
      }
    }
  }
  __device__ __forceinline__ void grad(int u, float* da) const {
    unpack8(qd[u][0], da);
    if (FAN) {
      float t[8];
#pragma unroll
      for (int j = 1; j < 4; ++j)
        if (j < nq[u]) {
          unpack8(qd[u][j], t);
#pragma unroll
          for (int k = 0; k < 8; ++k) da[k] += t[k];
        }
    }
  }
};

template <bool BWD, bool FAN>
__global__ void __launch_bounds__(256)
bn2d_reduce_fast_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy_up, int Hs, int Ws, int Hu,
                        int Wu, int C, int rpc, const float* __restrict__ stats, const float* __restrict__ gamma,
                        const float* __restrict__ beta, long slot_stride, const float* __restrict__ mask,
                        unsigned long long seed, float p_drop, const es_group* __restrict__ grp, int E, int total_rows,
                        int two_pass, double* __restrict__ sums) {
  __shared__ float s_a[256], s_b[256];
  __shared__ int ylo[64], yhi[64], xlo[64], xhi[64];
  const int P = Hs * Ws, c4 = C / 8;
  const int tid = threadIdx.x, cu = tid % c4, c8 = cu * 8, pstep = 256 / c4, p0 = tid / c4;
  if (BWD && FAN && tid == 0) { build_fanin(Hs, Hu, ylo, yhi); build_fanin(Ws, Wu, xlo, xhi); }
  s_a[tid] = 0.f; s_b[tid] = 0.f;
  __syncthreads();
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float mu[8], rs[8], ga[8], be[8];
  int cur = -1;
  auto flush = [&]() {      // called by ALL threads of the CTA (cur is uniform)
#pragma unroll
    for (int k = 0; k < 8; ++k) { atomicAdd(&s_a[c8 + k], a[k]); atomicAdd(&s_b[c8 + k], b[k]); a[k] = 0.f; b[k] = 0.f; }
    __syncthreads();
    if (tid < C) {
      double* d = sums + ((size_t)cur * C + tid) * 2;
      atomicAdd(d, (double)s_a[tid]);
      atomicAdd(d + 1, (double)s_b[tid]);
      s_a[tid] = 0.f; s_b[tid] = 0.f;
    }
    __syncthreads();
  };
  constexpr int U = FAN ? 2 : 4;
  BnFetch<BWD, FAN, U> q;
  const int r0 = blockIdx.x * rpc, r1 = min(total_rows, r0 + rpc);
  for (int r = r0; r < r1; ++r) {
    const SgRow sr = sg_of_row(grp, E, r, two_pass);
    const int sg = sr.g < 0 ? -1 : sr.sg;
    if (sg != cur) {
      if (cur >= 0) flush();
      cur = sg;
      if (BWD && sg >= 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          mu[k] = stats[((size_t)sg * C + c8 + k) * 2]; rs[k] = stats[((size_t)sg * C + c8 + k) * 2 + 1];
          ga[k] = gamma[sr.slot * slot_stride + c8 + k]; be[k] = beta[sr.slot * slot_stride + c8 + k];
        }
      }
    }
    if (sg < 0) continue;
    const uint4* x4 = reinterpret_cast<const uint4*>(x + (size_t)r * P * C);
    const uint4* d4 = BWD ? reinterpret_cast<const uint4*>(dy_up + (size_t)r * Hu * Wu * C) : nullptr;
    float f[8], da[8], ks[8];
    for (int pix0 = p0; pix0 < P; pix0 += U * pstep) {
      q.load(x4, d4, pix0, pstep, P, c4, cu, Ws, Wu, ylo, yhi, xlo, xhi);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pix = pix0 + u * pstep;
        if (pix >= P) continue;
        unpack8(q.qx[u], f);
        if (!BWD) {
#pragma unroll
          for (int k = 0; k < 8; ++k) { a[k] += f[k]; b[k] += f[k] * f[k]; }
        } else {
          q.grad(u, da);
          keep8(mask, seed, ((size_t)r * P + pix) * c4 + cu, ((size_t)r * C + c8) * P + pix, P, p_drop, ks);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float xh = (f[k] - mu[k]) * rs[k];
            const float pre = (xh * ga[k] + be[k]) * ks[k];            // dropout, then LeakyReLU
            const float g2 = da[k] * (pre > 0.f ? 1.f : kLReLU) * ks[k];
            a[k] += g2;
            b[k] += g2 * xh;
          }
        }
      }
    }
  }
  if (cur >= 0) flush();
}

template <bool BWD, bool FAN>
__global__ void __launch_bounds__(256)
bn2d_apply_fast_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy_up, int Hs, int Ws, int Hu,
                       int Wu, int C, const float* __restrict__ stats, const double* __restrict__ sums2,
                       const float* __restrict__ n_sg, const float* __restrict__ gamma, const float* __restrict__ beta,
                       long slot_stride, const float* __restrict__ mask, unsigned long long seed, float p_drop,
                       const es_group* __restrict__ grp, int E, int two_pass, __nv_bfloat16* __restrict__ out) {
  __shared__ int ylo[64], yhi[64], xlo[64], xhi[64];
  const int r = blockIdx.x;
  const SgRow sr = sg_of_row(grp, E, r, two_pass);
  const int P = Hs * Ws, c4 = C / 8;
  uint4* o4 = reinterpret_cast<uint4*>(out + (size_t)r * P * C);
  if (sr.g < 0) {
    for (int i = threadIdx.x; i < P * c4; i += blockDim.x) o4[i] = make_uint4(0, 0, 0, 0);   // finite rows for skipped experts: see gn_lrelu_kernel
    return;
  }
  if (BWD && FAN) {
    if (threadIdx.x == 0) { build_fanin(Hs, Hu, ylo, yhi); build_fanin(Ws, Wu, xlo, xhi); }
    __syncthreads();
  }
  const int tid = threadIdx.x, cu = tid % c4, c8 = cu * 8, pstep = 256 / c4, p0 = tid / c4;
  float mu[8], rs[8], ga[8], be[8], m1[8], m2[8];
  const float inv_n = BWD ? 1.f / n_sg[sr.sg] : 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const size_t si = ((size_t)sr.sg * C + c8 + k) * 2;
    mu[k] = stats[si]; rs[k] = stats[si + 1];
    ga[k] = gamma[sr.slot * slot_stride + c8 + k]; be[k] = beta[sr.slot * slot_stride + c8 + k];
    m1[k] = BWD ? (float)sums2[si] * inv_n : 0.f;
    m2[k] = BWD ? (float)sums2[si + 1] * inv_n : 0.f;
  }
  const uint4* x4 = reinterpret_cast<const uint4*>(x + (size_t)r * P * C);
  const uint4* d4 = BWD ? reinterpret_cast<const uint4*>(dy_up + (size_t)r * Hu * Wu * C) : nullptr;
  constexpr int U = FAN ? 2 : 4;
  BnFetch<BWD, FAN, U> q;
  float f[8], da[8], ks[8], o[8];
  for (int pix0 = p0; pix0 < P; pix0 += U * pstep) {
    q.load(x4, d4, pix0, pstep, P, c4, cu, Ws, Wu, ylo, yhi, xlo, xhi);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = pix0 + u * pstep;
      if (pix >= P) continue;
      unpack8(q.qx[u], f);
      if (BWD) q.grad(u, da);
      keep8(mask, seed, ((size_t)r * P + pix) * c4 + cu, ((size_t)r * C + c8) * P + pix, P, p_drop, ks);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xh = (f[k] - mu[k]) * rs[k];
        const float pre = (xh * ga[k] + be[k]) * ks[k];
        if (!BWD) {
          o[k] = lrelu(pre);
        } else {
          const float g2 = da[k] * (pre > 0.f ? 1.f : kLReLU) * ks[k];
          o[k] = ga[k] * rs[k] * (g2 - m1[k] - xh * m2[k]);
        }
      }
      o4[(size_t)pix * c4 + cu] = pack8(o);
    }
  }
}

// ------------------------------------------------------------------------------------------------ NCHW fp32 (aux regressor)
// order in the reference: BatchNorm -> LeakyReLU -> Dropout.   CTA = (channel, chunk of rows).
template <bool BWD>
__global__ void __launch_bounds__(256)
bn_reduce_nchw_kernel(const float* __restrict__ x, const float* __restrict__ dy, int C, int P, int per,
                      const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                      long slot_stride, const float* __restrict__ mask, unsigned long long seed, float p_drop,
                      const es_group* __restrict__ grp, int E, int total_rows, double* __restrict__ sums) {
  __shared__ float red[32];
  const int c = blockIdx.x, r0 = blockIdx.y * per, r1 = min(total_rows, r0 + per);
  int cur = -1;
  float a = 0.f, b = 0.f;
  for (int r = r0; r < r1; ++r) {
    const SgRow q = sg_of_row(grp, E, r, 0);     // uniform across the CTA
    if (q.sg != cur || q.g < 0) {
      if (cur >= 0) {
        a = block_sum(a, red); b = block_sum(b, red);
        if (threadIdx.x == 0) { atomicAdd(sums + ((size_t)cur * C + c) * 2, (double)a); atomicAdd(sums + ((size_t)cur * C + c) * 2 + 1, (double)b); }
      }
      a = 0.f; b = 0.f;
      cur = q.g < 0 ? -1 : q.sg;
    }
    if (q.g < 0) continue;
    const size_t base = ((size_t)r * C + c) * P;
    float mu = 0.f, rs = 0.f, ga = 0.f, be = 0.f;
    if (BWD) {
      mu = stats[((size_t)q.sg * C + c) * 2]; rs = stats[((size_t)q.sg * C + c) * 2 + 1];
      ga = gamma[q.slot * slot_stride + c]; be = beta[q.slot * slot_stride + c];
    }
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
      const float xv = x[base + p];
      if (!BWD) { a += xv; b += xv * xv; }
      else {
        const float xh = (xv - mu) * rs;
        const float pre = xh * ga + be;
        const float g2 = dy[base + p] * keep_scale(mask, seed, base + p, p_drop) * (pre > 0.f ? 1.f : kLReLU);
        a += g2; b += g2 * xh;
      }
    }
  }
  if (cur >= 0) {
    a = block_sum(a, red); b = block_sum(b, red);
    if (threadIdx.x == 0) { atomicAdd(sums + ((size_t)cur * C + c) * 2, (double)a); atomicAdd(sums + ((size_t)cur * C + c) * 2 + 1, (double)b); }
  }
}

template <bool BWD>
__global__ void __launch_bounds__(256)
bn_apply_nchw_kernel(const float* __restrict__ x, const float* __restrict__ dy, int C, int P,
                     const float* __restrict__ stats, const double* __restrict__ sums2, const float* __restrict__ n_sg,
                     const float* __restrict__ gamma, const float* __restrict__ beta, long slot_stride,
                     const float* __restrict__ mask, unsigned long long seed, float p_drop,
                     const es_group* __restrict__ grp, int E, float* __restrict__ out) {
  const int r = blockIdx.x;
  const SgRow q = sg_of_row(grp, E, r, 0);
  if (q.g < 0) return;
  const float inv_n = BWD ? 1.f / n_sg[q.sg] : 0.f;
  for (int i = threadIdx.x; i < C * P; i += blockDim.x) {
    const int c = i / P;
    const size_t si = ((size_t)q.sg * C + c) * 2, idx = (size_t)r * C * P + i;
    const float mu = stats[si], rs = stats[si + 1];
    const float ga = gamma[q.slot * slot_stride + c], be = beta[q.slot * slot_stride + c];
    const float xh = (x[idx] - mu) * rs;
    const float pre = xh * ga + be;
    const float ks = keep_scale(mask, seed, idx, p_drop);
    if (!BWD) out[idx] = lrelu(pre) * ks;
    else {
      const float g2 = dy[idx] * ks * (pre > 0.f ? 1.f : kLReLU);
      out[idx] = ga * rs * (g2 - (float)sums2[si] * inv_n - xh * (float)sums2[si + 1] * inv_n);
    }
  }
}

// ------------------------------------------------------------------------------------------------ finalise
// sums -> (mean, rstd); training additionally advances running_mean / running_var (pass 0 then pass 1, like two
// consecutive reference forward calls) and num_batches_tracked.  eval: stats come from the running buffers.
__global__ void bn_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ n_sg, int CS, int passes,
                                   int training, float momentum, const int* __restrict__ chmap, float* __restrict__ rmean,
                                   float* __restrict__ rvar, long buf_stride, long long* __restrict__ nbt, long nbt_stride,
                                   const es_group* __restrict__ grp, int slots, float* __restrict__ stats) {
  const int slot = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (grp && grp[slot].rows == 0) return;
  if (c == 0 && training && nbt) nbt[slot * nbt_stride] += passes;
  if (c >= CS) return;
  const int pc = chmap ? chmap[c] : c;
  for (int ps = 0; ps < passes; ++ps) {
    const int sg = 2 * slot + ps;
    const size_t si = ((size_t)sg * CS + c) * 2;
    if (training) {
      const double n = (double)n_sg[sg];
      const double mean = sums[si] / n;
      double var = sums[si + 1] / n - mean * mean;
      var = var > 0.0 ? var : 0.0;
      stats[si] = (float)mean;
      stats[si + 1] = (float)(1.0 / sqrt(var + (double)kNormEps));
      if (rmean) {
        const double unb = n > 1.0 ? var * n / (n - 1.0) : var;
        float* rm = rmean + slot * buf_stride + pc;
        float* rv = rvar + slot * buf_stride + pc;
        *rm = (1.f - momentum) * *rm + momentum * (float)mean;
        *rv = (1.f - momentum) * *rv + momentum * (float)unb;
      }
    } else {
      stats[si] = rmean[slot * buf_stride + pc];
      stats[si + 1] = rsqrtf(rvar[slot * buf_stride + pc] + kNormEps);
    }
  }
}

// dgamma[slot][c] += scale * sum_pass sum(g2*xhat), dbeta += scale * sum_pass sum(g2)
__global__ void bn_affine_grads_kernel(const double* __restrict__ sums2, int CS, int passes, float scale,
                                       const int* __restrict__ chmap, const es_group* __restrict__ grp,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, long slot_stride) {
  const int slot = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= CS || (grp && grp[slot].rows == 0)) return;
  const int pc = chmap ? chmap[c] : c;
  double g = 0.0, b = 0.0;
  for (int ps = 0; ps < passes; ++ps) {
    const size_t si = ((size_t)(2 * slot + ps) * CS + c) * 2;
    b += sums2[si];
    g += sums2[si + 1];
  }
  dgamma[slot * slot_stride + pc] += scale * (float)g;
  dbeta[slot * slot_stride + pc] += scale * (float)b;
}

int pick_opb(int C) {
  int opb = 32;
  while (opb > 1 && (C / 8) % opb != 0) opb >>= 1;
  return opb;
}

}  // namespace
}  // namespace es

using namespace es;

#define BN_GEOM_OK(Hs, Ws, Hu, Wu, C) ((C) > 0 && (C) % 8 == 0 && (Hs) > 0 && (Ws) > 0 && (Hs) <= 64 && (Ws) <= 64 && (Hu) <= 64 && (Wu) <= 64 && (Hu) >= (Hs) && (Wu) >= (Ws))

// BatchNorm2d fast path: per-channel statistics, C/8 octets divide the 256 threads, no channel map, fan-in <= 2 per axis
static bool bn2d_fast(int feat_stats, int C, const int32_t* chmap, int Hs, int Ws, int Hu, int Wu) {
  static const bool on = [] { const char* e = getenv("ES_BN_LEGACY"); return !(e && e[0] == '1'); }();
  return on && !feat_stats && !chmap && C % 8 == 0 && C <= 256 && 256 % (C / 8) == 0 && Hu <= 2 * Hs && Wu <= 2 * Ws;
}

static int rows_per_cta(int total_rows, int P_it) {
  int rpc = 4096 / (P_it > 0 ? P_it : 1);
  if (rpc < 1) rpc = 1;
  if (rpc > 64) rpc = 64;
  return rpc;
}

/* geometry arguments shared by the NHWC entry points: spatial view (Hs, Ws, C) of a row; feat_stats != 0 selects
 * BatchNorm1d over the flattened map (statistics per feature), else BatchNorm2d (statistics per channel) */
extern "C" int es_bn_stats_nhwc(const void* x, int Hs, int Ws, int C, int feat_stats, const es_group* grp, int E,
                                int total_rows, int two_pass, double* sums, void* stream) {
  ES_REQUIRE(x && grp && sums && Hs > 0 && Ws > 0 && C > 0 && C % 8 == 0 && total_rows > 0 && E >= 1 && E <= kMaxGroups, "bad arguments");
  const int P_it = feat_stats ? 1 : Hs * Ws, C_it = feat_stats ? Hs * Ws * C : C;
  const int opb = pick_opb(C_it), rpc = rows_per_cta(total_rows, P_it);
  if (bn2d_fast(feat_stats, C, nullptr, Hs, Ws, Hs, Ws)) {
    bn2d_reduce_fast_kernel<false, false><<<ceil_div(total_rows, rpc), 256, 0, as_stream(stream)>>>(
        (const __nv_bfloat16*)x, nullptr, Hs, Ws, Hs, Ws, C, rpc, nullptr, nullptr, nullptr, 0, nullptr, 0ull, 0.f, grp, E,
        total_rows, two_pass, sums);
    ES_LAUNCH_CHECK();
    return ES_OK;
  }
  bn_reduce_nhwc_kernel<false><<<dim3(ceil_div(total_rows, rpc) + (P_it == 1 ? 2 * E : 0), C_it / 8 / opb), 256, 0, as_stream(stream)>>>(
      (const __nv_bfloat16*)x, nullptr, Hs, Ws, Hs, Ws, C, P_it, C_it, opb, rpc, nullptr, nullptr, nullptr, 0, nullptr, nullptr,
      0ull, 0.f, grp, E, total_rows, two_pass, sums);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_bn_finalize(const double* sums, const float* n_sg, int CS, int passes, int training, float momentum,
                              const int32_t* chmap, float* running_mean, float* running_var, long buf_slot_stride,
                              int64_t* num_batches_tracked, long nbt_slot_stride, const es_group* grp, int slots,
                              float* stats, void* stream) {
  ES_REQUIRE(stats && CS > 0 && slots >= 1 && (passes == 1 || passes == 2), "bad arguments");
  ES_REQUIRE(training ? (sums && n_sg) : (running_mean && running_var), "missing statistics source");
  bn_finalize_kernel<<<dim3(ceil_div(CS, 256), slots), 256, 0, as_stream(stream)>>>(
      sums, n_sg, CS, passes, training, momentum, chmap, running_mean, running_var, buf_slot_stride,
      (long long*)num_batches_tracked, nbt_slot_stride, grp, slots, stats);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_bn_apply_fwd_nhwc(const void* x, int Hs, int Ws, int C, int feat_stats, const float* stats,
                                    const float* gamma, const float* beta, long slot_stride, const int32_t* chmap,
                                    const float* keep_mask, unsigned long long seed, float p_drop, const es_group* grp, int E,
                                    int total_rows, int two_pass, void* y, void* stream) {
  ES_REQUIRE(x && stats && gamma && beta && grp && y && Hs > 0 && Ws > 0 && C > 0 && C % 8 == 0 && total_rows > 0, "bad arguments");
  if (bn2d_fast(feat_stats, C, chmap, Hs, Ws, Hs, Ws)) {
    bn2d_apply_fast_kernel<false, false><<<total_rows, 256, 0, as_stream(stream)>>>(
        (const __nv_bfloat16*)x, nullptr, Hs, Ws, Hs, Ws, C, stats, nullptr, nullptr, gamma, beta, slot_stride, keep_mask, seed,
        p_drop, grp, E, two_pass, (__nv_bfloat16*)y);
    ES_LAUNCH_CHECK();
    return ES_OK;
  }
  bn_apply_nhwc_kernel<false><<<total_rows, 256, 0, as_stream(stream)>>>(
      (const __nv_bfloat16*)x, nullptr, Hs, Ws, Hs, Ws, C, feat_stats ? Hs * Ws * C : C, stats, nullptr, nullptr, gamma, beta,
      slot_stride, chmap, keep_mask, seed, p_drop, grp, E, two_pass, (__nv_bfloat16*)y);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_bn_bwd_reduce_nhwc(const void* dy_up, int Hs, int Ws, int Hu, int Wu, int C, int feat_stats, const void* x,
                                     const float* stats, const float* gamma, const float* beta, long slot_stride,
                                     const int32_t* chmap, const float* keep_mask, unsigned long long seed, float p_drop,
                                     const es_group* grp, int E, int total_rows, int two_pass, double* sums2, void* stream) {
  ES_REQUIRE(dy_up && x && stats && gamma && beta && grp && sums2 && total_rows > 0, "bad arguments");
  ES_REQUIRE(BN_GEOM_OK(Hs, Ws, Hu, Wu, C), "bad geometry");
  const int P_it = feat_stats ? 1 : Hs * Ws, C_it = feat_stats ? Hs * Ws * C : C;
  const int opb = pick_opb(C_it), rpc = rows_per_cta(total_rows, P_it);
  if (bn2d_fast(feat_stats, C, chmap, Hs, Ws, Hu, Wu)) {
    const bool fan = Hu != Hs || Wu != Ws;
    const dim3 grid(ceil_div(total_rows, rpc));
#define ES_BN_RED(FANV)                                                                                                \
    bn2d_reduce_fast_kernel<true, FANV><<<grid, 256, 0, as_stream(stream)>>>(                                          \
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy_up, Hs, Ws, Hu, Wu, C, rpc, stats, gamma, beta, slot_stride,  \
        keep_mask, seed, p_drop, grp, E, total_rows, two_pass, sums2);
    if (fan) { ES_BN_RED(true) } else { ES_BN_RED(false) }
#undef ES_BN_RED
    ES_LAUNCH_CHECK();
    return ES_OK;
  }
  bn_reduce_nhwc_kernel<true><<<dim3(ceil_div(total_rows, rpc) + (P_it == 1 ? 2 * E : 0), C_it / 8 / opb), 256, 0, as_stream(stream)>>>(
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy_up, Hs, Ws, Hu, Wu, C, P_it, C_it, opb, rpc, stats, gamma, beta,
      slot_stride, chmap, keep_mask, seed, p_drop, grp, E, total_rows, two_pass, sums2);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_bn_bwd_apply_nhwc(const void* dy_up, int Hs, int Ws, int Hu, int Wu, int C, int feat_stats, const void* x,
                                    const float* stats, const double* sums2, const float* n_sg, const float* gamma,
                                    const float* beta, long slot_stride, const int32_t* chmap, const float* keep_mask,
                                    unsigned long long seed, float p_drop, const es_group* grp, int E, int total_rows,
                                    int two_pass, void* dx, void* stream) {
  ES_REQUIRE(dy_up && x && stats && sums2 && n_sg && gamma && beta && grp && dx && total_rows > 0, "bad arguments");
  ES_REQUIRE(BN_GEOM_OK(Hs, Ws, Hu, Wu, C), "bad geometry");
  if (bn2d_fast(feat_stats, C, chmap, Hs, Ws, Hu, Wu)) {
    const bool fan = Hu != Hs || Wu != Ws;
#define ES_BN_APP(FANV)                                                                                                \
    bn2d_apply_fast_kernel<true, FANV><<<total_rows, 256, 0, as_stream(stream)>>>(                                     \
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy_up, Hs, Ws, Hu, Wu, C, stats, sums2, n_sg, gamma, beta,       \
        slot_stride, keep_mask, seed, p_drop, grp, E, two_pass, (__nv_bfloat16*)dx);
    if (fan) { ES_BN_APP(true) } else { ES_BN_APP(false) }
#undef ES_BN_APP
    ES_LAUNCH_CHECK();
    return ES_OK;
  }
  bn_apply_nhwc_kernel<true><<<total_rows, 256, 0, as_stream(stream)>>>(
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy_up, Hs, Ws, Hu, Wu, C, feat_stats ? Hs * Ws * C : C, stats, sums2, n_sg,
      gamma, beta, slot_stride, chmap, keep_mask, seed, p_drop, grp, E, two_pass, (__nv_bfloat16*)dx);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_bn_affine_grads(const double* sums2, int CS, int passes, float scale, const int32_t* chmap,
                                  const es_group* grp, int slots, float* dgamma, float* dbeta, long slot_stride,
                                  void* stream) {
  ES_REQUIRE(sums2 && dgamma && dbeta && CS > 0 && slots >= 1 && (passes == 1 || passes == 2), "bad arguments");
  bn_affine_grads_kernel<<<dim3(ceil_div(CS, 256), slots), 256, 0, as_stream(stream)>>>(sums2, CS, passes, scale, chmap, grp,
                                                                                       dgamma, dbeta, slot_stride);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_bn2d_stats(const float* x, int C, int P, const es_group* grp, int E, int total_rows, double* sums,
                             void* stream) {
  ES_REQUIRE(x && grp && sums && C > 0 && P > 0 && total_rows > 0 && E >= 1 && E <= kMaxGroups, "bad arguments");
  const int per = max(1, ceil_div(total_rows * C, 8 * 148));
  bn_reduce_nchw_kernel<false><<<dim3(C, ceil_div(total_rows, per)), 256, 0, as_stream(stream)>>>(
      x, nullptr, C, P, per, nullptr, nullptr, nullptr, 0, nullptr, 0ull, 0.f, grp, E, total_rows, sums);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_bn2d_apply_fwd(const float* x, int C, int P, const float* stats, const float* gamma, const float* beta,
                                 long slot_stride, const float* keep_mask, unsigned long long seed, float p_drop,
                                 const es_group* grp, int E, int total_rows, float* y, void* stream) {
  ES_REQUIRE(x && stats && gamma && beta && grp && y && C > 0 && P > 0 && total_rows > 0, "bad arguments");
  bn_apply_nchw_kernel<false><<<total_rows, 256, 0, as_stream(stream)>>>(x, nullptr, C, P, stats, nullptr, nullptr, gamma, beta,
                                                                        slot_stride, keep_mask, seed, p_drop, grp, E, y);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_bn2d_bwd_reduce(const float* dy, const float* x, int C, int P, const float* stats, const float* gamma,
                                  const float* beta, long slot_stride, const float* keep_mask, unsigned long long seed,
                                  float p_drop, const es_group* grp, int E, int total_rows, double* sums2, void* stream) {
  ES_REQUIRE(dy && x && stats && gamma && beta && grp && sums2 && C > 0 && P > 0 && total_rows > 0, "bad arguments");
  const int per = max(1, ceil_div(total_rows * C, 8 * 148));
  bn_reduce_nchw_kernel<true><<<dim3(C, ceil_div(total_rows, per)), 256, 0, as_stream(stream)>>>(
      x, dy, C, P, per, stats, gamma, beta, slot_stride, keep_mask, seed, p_drop, grp, E, total_rows, sums2);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_bn2d_bwd_apply(const float* dy, const float* x, int C, int P, const float* stats, const double* sums2,
                                 const float* n_sg, const float* gamma, const float* beta, long slot_stride,
                                 const float* keep_mask, unsigned long long seed, float p_drop, const es_group* grp, int E,
                                 int total_rows, float* dx, void* stream) {
  ES_REQUIRE(dy && x && stats && sums2 && n_sg && gamma && beta && grp && dx && total_rows > 0, "bad arguments");
  bn_apply_nchw_kernel<true><<<total_rows, 256, 0, as_stream(stream)>>>(x, dy, C, P, stats, sums2, n_sg, gamma, beta,
                                                                       slot_stride, keep_mask, seed, p_drop, grp, E, dx);
  ES_LAUNCH_CHECK();
  return ES_OK;
}
