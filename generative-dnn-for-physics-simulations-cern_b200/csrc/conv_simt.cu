// K3 convolutions as fp32 SIMT implicit GEMMs (discriminator + auxiliary regressor, forward / data gradient / weight
// gradient), grouped by expert.  The channel counts of these networks (1..256, mostly 32/64) and their fp32 parity bar
// keep them off the tensor cores; the first direct kernels (one thread = one pixel x 8 channels, a whole sample staged in
// shared memory) ran at 0.5-3 TFLOP/s (r01 launch list: 27 ms of an 89 ms step).  Here every conv is a register-tiled
// GEMM: CTA tile 64 (pixels, sample-major inside one expert group) x BN (channels) x 16 (reduction), 256 threads, thread
// tile 4 x BN/16, operands gathered into shared memory with lanes along the pixel axis (coalesced NCHW reads/writes),
// reduction index -> (channel, ky, kx) through a per-CTA lookup table.
//   FWD       y[m, co]  = sum_k xcol[m, k]  w[co, k]              k = (ci, ky, kx)
//   BWD_DATA  dx[m, ci] = sum_k dycol[m, k] w[co, ci, ky, kx]     k = (co, ky, kx), m = input pixel
//   BWD_W     dw[co, k] += sum_m dy[m, co] xcol[m, k]             split over m, fp32 atomics
#include <stdlib.h>

#include "common.cuh"

namespace es {
int conv2d_bwd_weight_fewtaps(const float* x, const float* dy, const es_conv2d* g, const es_group* grp, int n_groups,
                              int total_rows, float* dw, float* db, long slot_stride_w, long slot_stride_b, void* stream);
int conv2d_bwd_data_ci1(const float* dy, const float* w, long slot_stride_w, const es_conv2d* g, const es_group* grp,
                        int n_groups, int total_rows, float* dx, int accumulate, void* stream);
namespace {

constexpr int kCM = 64, kCK = 16;
constexpr int kPA = kCM + 8;          // pitch of the [reduction][64] operand tile: 72 -> mma fragment loads hit 32 distinct banks

// ---- 3xTF32 on the (legacy, warp-level) tensor path.  The discriminator / aux-regressor convolutions keep an fp32 parity bar
// (2e-3 on gradients that cancel heavily), which plain TF32 (10 mantissa bits) misses and bf16 tcgen05 misses by far.  Split
// every fp32 operand x = big + small with big = tf32(x), small = tf32(x - big): a*b ~= big_a*big_b + big_a*small_b +
// small_a*big_b drops only the small*small term (2^-22 relative) and accumulates in fp32 — fp32-class accuracy at one
// third of the TF32 rate, still several times the FFMA rate of the register-tiled loop it replaces (which ran at
// 10-12 TFLOP/s: 2 shared-memory loads per 8 FMAs).
__device__ __forceinline__ void split_tf32(float x, uint32_t& big, uint32_t& small) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(big) : "f"(x));
  const float r = x - __uint_as_float(big);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(small) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// One 16-deep reduction chunk of the CTA tile D[64 x BN] += A^T B with A = As[16][kPA] (reduction-major, 64 rows of D
// contiguous), B = Bs[16][BN + 8].  8 warps: warp w owns rows 16*(w&3).. and columns (w>>2)*(BN/2)..; c[j][4] = its BN/16 n8 tiles.
template <int BN>
__device__ __forceinline__ void chunk_mma_tf32x3(const float* __restrict__ As, const float* __restrict__ Bs, float (*c)[4]) {
  constexpr int NT = BN / 16, PB = BN + 8;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int r0 = (w & 3) * 16 + g, n0 = (w >> 2) * (BN / 2) + g;
#pragma unroll
  for (int k8 = 0; k8 < kCK; k8 += 8) {
    uint32_t ab[4], as[4];
    split_tf32(As[(k8 + t) * kPA + r0], ab[0], as[0]);
    split_tf32(As[(k8 + t) * kPA + r0 + 8], ab[1], as[1]);
    split_tf32(As[(k8 + t + 4) * kPA + r0], ab[2], as[2]);
    split_tf32(As[(k8 + t + 4) * kPA + r0 + 8], ab[3], as[3]);
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      uint32_t bb[2], bs[2];
      split_tf32(Bs[(k8 + t) * PB + n0 + 8 * j], bb[0], bs[0]);
      split_tf32(Bs[(k8 + t + 4) * PB + n0 + 8 * j], bb[1], bs[1]);
      mma_tf32(c[j], as, bb);      // small terms first: they are added to the accumulator before the large product
      mma_tf32(c[j], ab, bs);
      mma_tf32(c[j], ab, bb);
    }
  }
}
// accumulator fragments -> the [BN][64 + 4] staging tile the SIMT epilogue indexing reads (row of D contiguous)
template <int BN>
__device__ __forceinline__ void stage_mma_acc(float* __restrict__ Cs, const float (*c)[4]) {
  constexpr int NT = BN / 16;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int r0 = (w & 3) * 16 + g, n0 = (w >> 2) * (BN / 2) + 2 * t;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    Cs[(n0 + 8 * j) * (kCM + 4) + r0] = c[j][0];
    Cs[(n0 + 8 * j + 1) * (kCM + 4) + r0] = c[j][1];
    Cs[(n0 + 8 * j) * (kCM + 4) + r0 + 8] = c[j][2];
    Cs[(n0 + 8 * j + 1) * (kCM + 4) + r0 + 8] = c[j][3];
  }
}

struct KEnt { int off; short ky, kx; };   // source offset of reduction index k relative to the window origin

__device__ __forceinline__ bool tile_of(const es_group* grp, int n_groups, int px_per_row, int tile_px, int t, int& g,
                                        int& m0, int& mtot) {
  for (int i = 0; i < n_groups; ++i) {
    const int tot = grp[i].rows * px_per_row;
    const int nt = ceil_div(tot, tile_px);
    if (t < nt) { g = i; m0 = t * tile_px; mtot = tot; return true; }
    t -= nt;
  }
  return false;
}

// MODE 0: forward.  MODE 1: data gradient (src = dy, "output" = dx over input pixels).
template <int MODE, int BN, bool TC>
__global__ void __launch_bounds__(256)
conv_gemm_kernel(const float* __restrict__ src, const float* __restrict__ w, const float* __restrict__ bias, long sw,
                 long sb, es_conv2d g, const es_group* __restrict__ grp, int n_groups, float* __restrict__ dst,
                 int accumulate) {
  extern __shared__ float sm[];
  constexpr int TN = BN / 16;
  const int KHW = g.KH * g.KW;
  const int Nn = MODE == 0 ? g.Co : g.Ci;                        // GEMM N extent
  const int Hf = MODE == 0 ? g.Ho : g.Hi, Wf = MODE == 0 ? g.Wo : g.Wi;      // full pixel grid of the destination
  const int Hs = MODE == 0 ? g.Hi : g.Ho, Ws = MODE == 0 ? g.Wi : g.Wo;      // pixel grid of the source
  const int Cs = MODE == 0 ? g.Ci : g.Co;
  // Data gradient of a strided conv: an input pixel (iy, ix) only meets the taps with ky = (iy + pad) mod stride (same in
  // x) — 6.25 of 25 for k5/s2.  blockIdx.z walks the stride x stride parity classes; each is a dense GEMM over its own
  // pixel sub-grid (iy = S*py + cy) and tap subset (ky = ky0 + S*i), so no reduction step multiplies by a structural zero.
  const int S = MODE == 1 ? g.stride : 1;
  const int cy = MODE == 1 ? (int)blockIdx.z / S : 0, cx = MODE == 1 ? (int)blockIdx.z % S : 0;
  const int Hm = (Hf - cy + S - 1) / S, Wm = (Wf - cx + S - 1) / S;          // pixel grid of the M axis (this class)
  const int ky0 = MODE == 1 ? (cy + g.pad) % S : 0, kx0 = MODE == 1 ? (cx + g.pad) % S : 0;
  const int nky = (g.KH - ky0 + S - 1) / S, nkx = (g.KW - kx0 + S - 1) / S;
  const int KHWc = nky * nkx;
  const int K = Cs * KHWc;                                       // reduction length
  const int PXm = Hm * Wm, PXs = Hs * Ws, PXf = Hf * Wf;
  if (PXm <= 0) return;   // (a class without taps, e.g. 1x1 stride 2, still writes its zeros)
  KEnt* ktab = reinterpret_cast<KEnt*>(sm);                       // [K]
  float* As = sm + 2 * ((K + 1) & ~1);                            // [kCK][kPA]
  float* Bs = As + kCK * kPA;                                     // [kCK][BN + 8]
  float* Cst = Bs + kCK * (BN + 8);                                // TC: [BN][kCM + 4] accumulator staging for the epilogue
  int gi, m0, mtot;
  if (!tile_of(grp, n_groups, PXm, kCM, blockIdx.x, gi, m0, mtot)) return;
  const int slot = grp[gi].slot, row_start = grp[gi].row_start;
  const int n0 = blockIdx.y * BN;
  const int tid = threadIdx.x;
  for (int k = tid; k < K; k += 256) {
    const int c = k / KHWc, t = k - c * KHWc, ky = ky0 + (t / nkx) * S, kx = kx0 + (t % nkx) * S;
    KEnt e;
    e.ky = (short)ky; e.kx = (short)kx;
    e.off = c * PXs;
    ktab[k] = e;
  }
  // this thread's gather row (fixed): m = m0 + (tid & 63)
  const int ml = tid & 63, kl = tid >> 6;                         // kl in 0..3: reduction rows kl, kl+4, kl+8, kl+12
  const int m = m0 + ml;
  const bool mv = m < mtot;
  const int smp = mv ? m / PXm : 0, pm = mv ? m - smp * PXm : 0;
  const int py = pm / Wm, px = pm - py * Wm;
  const float* sbase = src + (size_t)(row_start + smp) * Cs * PXs;
  const float* wslot = w + slot * sw;
  const int tm = tid & 15, tn = tid >> 4;
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float cfr[BN / 16][4];
#pragma unroll
  for (int j = 0; j < BN / 16; ++j) { cfr[j][0] = 0.f; cfr[j][1] = 0.f; cfr[j][2] = 0.f; cfr[j][3] = 0.f; }
  __syncthreads();
  // Software pipeline: the operands of reduction step k0 + 16 are fetched into registers while step k0 is multiplied out of
  // shared memory, so the global-load latency of the gather overlaps the FMA loop instead of preceding it.
  constexpr int NB = kCK * BN / 256;       // weight elements per thread and step
  float areg[4], breg[NB];
  auto fetch = [&](int k0) {
    // ---- A tile: gathered source values
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kk = kl + 4 * j, k = k0 + kk;
      float v = 0.f;
      if (mv && k < K) {
        const KEnt e = ktab[k];
        if (MODE == 0) {
          const int iy = py * g.stride + e.ky - g.pad, ix = px * g.stride + e.kx - g.pad;
          if (iy >= 0 && iy < Hs && ix >= 0 && ix < Ws) v = __ldg(sbase + e.off + iy * Ws + ix);
        } else {
          const int ty = py * S + cy + g.pad - e.ky, tx = px * S + cx + g.pad - e.kx;     // multiples of S by construction
          if (ty >= 0 && tx >= 0) {
            const int oy = ty / S, ox = tx / S;
            if (oy < Hs && ox < Ws) v = __ldg(sbase + e.off + oy * Ws + ox);
          }
        }
      }
      areg[j] = v;
    }
    // ---- B tile: weights
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const int i = tid + 256 * q;
      const int kk = i % kCK, n = i / kCK, k = k0 + kk;
      float v = 0.f;
      if (k < K && n0 + n < Nn) {
        if (MODE == 0) v = __ldg(wslot + (size_t)(n0 + n) * K + k);
        else {
          const KEnt e = ktab[k];
          v = __ldg(wslot + ((size_t)(k / KHWc) * g.Ci + n0 + n) * KHW + e.ky * g.KW + e.kx);
        }
      }
      breg[q] = v;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += kCK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) As[(kl + 4 * j) * kPA + ml] = areg[j];
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const int i = tid + 256 * q;
      Bs[(i % kCK) * (BN + 8) + i / kCK] = breg[q];
    }
    __syncthreads();
    if (k0 + kCK < K) fetch(k0 + kCK);
    if (TC) {
      chunk_mma_tf32x3<BN>(As, Bs, cfr);
    } else {
#pragma unroll
      for (int kk = 0; kk < kCK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(As + kk * kPA + tm * 4);
        float b[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[kk * (BN + 8) + tn * TN + j];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          acc[0][j] = fmaf(a.x, b[j], acc[0][j]);
          acc[1][j] = fmaf(a.y, b[j], acc[1][j]);
          acc[2][j] = fmaf(a.z, b[j], acc[2][j]);
          acc[3][j] = fmaf(a.w, b[j], acc[3][j]);
        }
      }
    }
    __syncthreads();
  }
  if (TC) {     // fragments -> shared memory -> the same (4 pixels x TN channels) register tile the epilogue below stores
    stage_mma_acc<BN>(Cst, cfr);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = Cst[(tn * TN + j) * (kCM + 4) + tm * 4 + i];
  }
  // ---- epilogue: NCHW store, lanes along pixels
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int mm = m0 + tm * 4 + i;
    if (mm >= mtot) continue;
    const int s2 = mm / PXm, p2 = mm - s2 * PXm;
    const int pf = MODE == 0 ? p2 : ((p2 / Wm) * S + cy) * Wf + (p2 % Wm) * S + cx;
    float* o = dst + (size_t)(row_start + s2) * Nn * PXf + pf;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tn * TN + j;
      if (n >= Nn) continue;
      float v = acc[i][j];
      if (MODE == 0 && bias) v += bias[slot * sb + n];
      if (accumulate) o[(size_t)n * PXf] += v;
      else o[(size_t)n * PXf] = v;
    }
  }
}

// weight gradient: CTA = (64-wide k tile, chunk of `mper` group pixels); D[k, co] += sum_m xcol[m, k] dy[m, co].
// The gather walks pixels with an incremental (sample, oy, ox) counter — no integer division in the reduction loop.
template <int BN, bool TC>
__global__ void __launch_bounds__(256)
conv_wgrad_gemm_kernel(const float* __restrict__ x, const float* __restrict__ dy, es_conv2d g,
                       const es_group* __restrict__ grp, int n_groups, int mper, float* __restrict__ dw,
                       float* __restrict__ db, long sw, long sb) {
  constexpr int kWK = 64;                                  // k tile
  __shared__ __align__(16) float As[kCK][kPA];         // [m chunk][k tile]
  __shared__ __align__(16) float Bs[kCK][BN + 8];      // [m chunk][co]
  __shared__ float Cs[TC ? BN * (kCM + 4) : 1];        // TC: accumulator staging for the epilogue
  constexpr int TN = BN / 16;
  const int KHW = g.KH * g.KW, K = g.Ci * KHW, PXo = g.Ho * g.Wo, PXi = g.Hi * g.Wi;
  int gi, mbeg, mtot;
  if (!tile_of(grp, n_groups, PXo, mper, blockIdx.y, gi, mbeg, mtot)) return;
  const int mend = min(mtot, mbeg + mper);
  const int slot = grp[gi].slot, row_start = grp[gi].row_start;
  const int k0 = blockIdx.x * kWK, n0 = blockIdx.z * BN;
  const int tid = threadIdx.x;
  // A gather: thread owns reduction column ka = k0 + (tid >> 2) and the 4 consecutive pixels ma .. ma+3 of every chunk
  const int ka = k0 + (tid >> 2), ma = (tid & 3) * 4;
  const bool kv = ka < K;
  const int ci = kv ? ka / KHW : 0, t = kv ? ka - ci * KHW : 0, ky = t / g.KW, kx = t - ky * g.KW;
  int a_s, a_oy, a_ox;                                     // (sample, oy, ox) of pixel mbeg + ma
  {
    const int mm = mbeg + ma;
    a_s = mm / PXo;
    const int p2 = mm - a_s * PXo;
    a_oy = p2 / g.Wo;
    a_ox = p2 - a_oy * g.Wo;
  }
  // B gather: thread owns channel nb = tid >> 4 (+16, ...) and pixel mb = tid & 15 of every chunk
  const int mb = tid & 15;
  int b_s, b_p;
  {
    const int mm = mbeg + mb;
    b_s = mm / PXo;
    b_p = mm - b_s * PXo;
  }
  const int tm = tid & 15, tn = tid >> 4;
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;
  float cfr[BN / 16][4];
#pragma unroll
  for (int j = 0; j < BN / 16; ++j) { cfr[j][0] = 0.f; cfr[j][1] = 0.f; cfr[j][2] = 0.f; cfr[j][3] = 0.f; }
  // Software pipeline (see conv_gemm_kernel): the operands of chunk mc + 16 are fetched into registers while chunk mc is
  // multiplied out of shared memory.
  constexpr int NBW = BN / 16;
  float areg[4], breg[NBW];
  auto fetch = [&](int mc) {
    // A tile: xcol[m, k] for 16 consecutive m, 64 k
    {
      int s2 = a_s, oy = a_oy, ox = a_ox;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = 0.f;
        if (kv && mc + ma + j < mend) {
          const int iy = oy * g.stride + ky - g.pad, ix = ox * g.stride + kx - g.pad;
          if (iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi)
            v = __ldg(x + ((size_t)(row_start + s2) * g.Ci + ci) * PXi + iy * g.Wi + ix);
        }
        areg[j] = v;
        if (++ox == g.Wo) { ox = 0; if (++oy == g.Ho) { oy = 0; ++s2; } }
      }
      // advance the base counter by one chunk (16 pixels)
      a_ox += kCK;
      while (a_ox >= g.Wo) { a_ox -= g.Wo; if (++a_oy == g.Ho) { a_oy = 0; ++a_s; } }
    }
    // B tile: dy[m, co]: lanes along m
#pragma unroll
    for (int q = 0; q < NBW; ++q) {
      const int nn = (tid >> 4) + 16 * q;
      float v = 0.f;
      if (mc + mb < mend && n0 + nn < g.Co) v = __ldg(dy + ((size_t)(row_start + b_s) * g.Co + n0 + nn) * PXo + b_p);
      breg[q] = v;
    }
    b_p += kCK;
    while (b_p >= PXo) { b_p -= PXo; ++b_s; }
  };
  if (mbeg < mend) fetch(mbeg);
  for (int mc = mbeg; mc < mend; mc += kCK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) As[ma + j][tid >> 2] = areg[j];
#pragma unroll
    for (int q = 0; q < NBW; ++q) Bs[mb][(tid >> 4) + 16 * q] = breg[q];
    __syncthreads();
    if (mc + kCK < mend) fetch(mc + kCK);
    if (db && blockIdx.x == 0 && tid < BN) {
#pragma unroll
      for (int mm = 0; mm < kCK; ++mm) bsum += Bs[mm][tid];
    }
    if (TC) {
      chunk_mma_tf32x3<BN>(&As[0][0], &Bs[0][0], cfr);
    } else {
#pragma unroll
      for (int mm = 0; mm < kCK; ++mm) {
        const float4 a = *reinterpret_cast<const float4*>(&As[mm][tm * 4]);
        float b[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[mm][tn * TN + j];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          acc[0][j] = fmaf(a.x, b[j], acc[0][j]);
          acc[1][j] = fmaf(a.y, b[j], acc[1][j]);
          acc[2][j] = fmaf(a.z, b[j], acc[2][j]);
          acc[3][j] = fmaf(a.w, b[j], acc[3][j]);
        }
      }
    }
    __syncthreads();
  }
  if (TC) {
    stage_mma_acc<BN>(Cs, cfr);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = Cs[(tn * TN + j) * (kCM + 4) + tm * 4 + i];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + tm * 4 + i;
    if (k >= K) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tn * TN + j;
      if (n < g.Co) atomicAdd(&dw[slot * sw + (size_t)n * K + k], acc[i][j]);
    }
  }
  if (db && blockIdx.x == 0 && tid < BN && n0 + tid < g.Co) atomicAdd(&db[slot * sb + n0 + tid], bsum);
}

bool geom_ok(const es_conv2d* g) {
  return g && g->Ci > 0 && g->Co > 0 && g->KH > 0 && g->KW > 0 && g->stride > 0 && g->pad >= 0 &&
         g->Ho == (g->Hi + 2 * g->pad - g->KH) / g->stride + 1 && g->Wo == (g->Wi + 2 * g->pad - g->KW) / g->stride + 1 &&
         g->Ho > 0 && g->Wo > 0;
}

template <int MODE>
int launch_conv_gemm(const float* src, const float* w, const float* b, long sw, long sb, const es_conv2d* g,
                     const es_group* grp, int n_groups, int total_rows, float* dst, int accumulate, cudaStream_t st) {
  const int KHW = g->KH * g->KW;
  const int K = (MODE == 0 ? g->Ci : g->Co) * KHW, Nn = MODE == 0 ? g->Co : g->Ci;
  const int S = MODE == 0 ? 1 : g->stride;
  const int PXm = MODE == 0 ? g->Ho * g->Wo : ceil_div(g->Hi, S) * ceil_div(g->Wi, S);   // largest parity class
  const long tiles = ceil_div_l((long)total_rows * PXm, kCM) + n_groups;
  if (tiles >= 2147483647L) { set_error("conv gemm: too many tiles"); return ES_ERR_INVALID; }
  const int BN = Nn > 32 ? 64 : (Nn > 16 ? 32 : 16);
  // 3xTF32 is OPT-IN (ES_CONV_TF32X3=1).  Measured on B200 (batch 1024, E = 8): proton 29.10k vs 28.84k samples/s, neutron
  // 27.16k vs 27.29k — these kernels are bound by their gather (per-element address arithmetic), not by the FFMA loop — while
  // the per-kernel error grows from ~1e-7 to 0.2-1.6e-5 and the step-level aux-regressor gradients leave the 2e-3 bar.
  static const bool tc = [] { const char* e = getenv("ES_CONV_TF32X3"); return e && e[0] == '1'; }();
  const size_t smem = (2 * (size_t)((K + 1) & ~1) + kCK * kPA + kCK * (BN + 8) + (tc ? BN * (kCM + 4) : 0)) * sizeof(float);
  if (smem > 200 * 1024) { set_error("conv gemm: reduction table does not fit in shared memory"); return ES_ERR_INVALID; }
  const dim3 grid((unsigned)tiles, ceil_div(Nn, BN), S * S);
#define ES_LAUNCH_CG(BNV)                                                                                              \
  {                                                                                                                    \
    if (tc) {                                                                                                          \
      if (smem > 48 * 1024) cudaFuncSetAttribute(conv_gemm_kernel<MODE, BNV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
      conv_gemm_kernel<MODE, BNV, true><<<grid, 256, smem, st>>>(src, w, b, sw, sb, *g, grp, n_groups, dst, accumulate); \
    } else {                                                                                                           \
      if (smem > 48 * 1024) cudaFuncSetAttribute(conv_gemm_kernel<MODE, BNV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
      conv_gemm_kernel<MODE, BNV, false><<<grid, 256, smem, st>>>(src, w, b, sw, sb, *g, grp, n_groups, dst, accumulate); \
    }                                                                                                                  \
  }
  if (BN == 64) ES_LAUNCH_CG(64) else if (BN == 32) ES_LAUNCH_CG(32) else ES_LAUNCH_CG(16)
#undef ES_LAUNCH_CG
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error(std::string("conv gemm launch: ") + cudaGetErrorString(e)); return ES_ERR_CUDA; }
  return ES_OK;
}

}  // namespace
}  // namespace es

using namespace es;

extern "C" int es_conv2d_fwd(const float* x, const float* w, const float* b, long slot_stride_w, long slot_stride_b,
                             const es_conv2d* g, const es_group* grp, int n_groups, int total_rows, float* y,
                             void* stream) {
  ES_REQUIRE(x && w && grp && y, "null pointer");
  ES_REQUIRE(geom_ok(g), "bad conv geometry");
  ES_REQUIRE(n_groups >= 1 && n_groups <= kMaxGroups && total_rows > 0, "bad sizes");
  return launch_conv_gemm<0>(x, w, b, slot_stride_w, slot_stride_b, g, grp, n_groups, total_rows, y, 0, as_stream(stream));
}

extern "C" int es_conv2d_bwd_data(const float* dy, const float* w, long slot_stride_w, const es_conv2d* g,
                                  const es_group* grp, int n_groups, int total_rows, float* dx, int accumulate,
                                  void* stream) {
  ES_REQUIRE(dy && w && grp && dx, "null pointer");
  ES_REQUIRE(geom_ok(g), "bad conv geometry");
  ES_REQUIRE(n_groups >= 1 && n_groups <= kMaxGroups && total_rows > 0, "bad sizes");
  if (g->Ci == 1)    // gradient w.r.t. the 1-channel image: a GEMM with N = 1 wastes the tile; direct kernel instead
    return conv2d_bwd_data_ci1(dy, w, slot_stride_w, g, grp, n_groups, total_rows, dx, accumulate, stream);
  return launch_conv_gemm<1>(dy, w, nullptr, slot_stride_w, 0, g, grp, n_groups, total_rows, dx, accumulate, as_stream(stream));
}

extern "C" int es_conv2d_bwd_weight(const float* x, const float* dy, const es_conv2d* g, const es_group* grp,
                                    int n_groups, int total_rows, float* dw, float* db, long slot_stride_w,
                                    long slot_stride_b, void* stream) {
  ES_REQUIRE(x && dy && grp && dw, "null pointer");
  ES_REQUIRE(geom_ok(g), "bad conv geometry");
  ES_REQUIRE(n_groups >= 1 && n_groups <= kMaxGroups && total_rows > 0, "bad sizes");
  const int K = g->Ci * g->KH * g->KW, PXo = g->Ho * g->Wo;
  if (K <= 9 && g->Co % 8 == 0)   // a 64-wide k tile would be 86 % padding: pixel-parallel register kernel instead
    return conv2d_bwd_weight_fewtaps(x, dy, g, grp, n_groups, total_rows, dw, db, slot_stride_w, slot_stride_b, stream);
  const int ktiles = ceil_div(K, 64);
  const int BN = g->Co > 32 ? 64 : (g->Co > 16 ? 32 : 16);
  const int ntiles = ceil_div(g->Co, BN);
  // split the pixel reduction so that ~4 waves of CTAs exist, but keep >= 256 pixels per CTA
  const long total_px = (long)total_rows * PXo;
  long chunks = ceil_div_l(4L * 148, (long)ktiles * ntiles);
  long mper = ceil_div_l(total_px, chunks);
  if (mper < 256) mper = 256;
  mper = ceil_div_l(mper, kCK) * kCK;
  const long ychunks = ceil_div_l(total_px, mper) + n_groups;
  ES_REQUIRE(ychunks < 65535 && mper < 2147483647L, "too many reduction chunks");
  const dim3 grid(ktiles, (unsigned)ychunks, ntiles);
  cudaStream_t st = as_stream(stream);
  static const bool tc = [] { const char* e = getenv("ES_CONV_TF32X3"); return e && e[0] == '1'; }();
#define ES_LAUNCH_WG(BNV, TCV) conv_wgrad_gemm_kernel<BNV, TCV><<<grid, 256, 0, st>>>(x, dy, *g, grp, n_groups, (int)mper, dw, db, slot_stride_w, slot_stride_b)
  if (tc) { if (BN == 64) ES_LAUNCH_WG(64, true); else if (BN == 32) ES_LAUNCH_WG(32, true); else ES_LAUNCH_WG(16, true); }
  else { if (BN == 64) ES_LAUNCH_WG(64, false); else if (BN == 32) ES_LAUNCH_WG(32, false); else ES_LAUNCH_WG(16, false); }
#undef ES_LAUNCH_WG
  ES_LAUNCH_CHECK();
  return ES_OK;
}
