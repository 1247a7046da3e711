// K2 — grouped (per-expert) bf16 implicit-GEMM engine on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// One kernel template, four problem kinds, all D[128 x BN] (fp32, TMEM) += A[128 x 64] * B[BN x 64]^T per pipeline stage:
//   FWD          y[pix, n]     = sum_k  im2col(x)[pix, k] * w[n, k]          A: gathered rows (K-major)  B: weights (K-major)
//                (forward convs with the nearest-upsample folded into the gather, their data gradients, fc2 forward)
//   WGRAD_CONV   dw[n, kk]    += sum_pix im2col(x)[pix, kk] * dy[pix, n]      A: gathered (MN-major)      B: dy (MN-major)
//   DENSE_DGRAD  dx[row, k]   += sum_n  dy[row, n] * w[n, k]                  A: dy rows (K-major)        B: w (MN-major)
//   DENSE_WGRAD  dw[n, k]      = sum_row dy[row, n] * x[row, k]               A: dy (MN-major)            B: x (MN-major)
//
// Shared-memory operand tiles are arrays of 128-byte rows in the canonical SWIZZLE_128B layout (16-byte chunk c of row r
// is stored at chunk c ^ (r & 7) inside its 1024-byte 8-row group), which serves both K-major tiles (row = m or n index,
// 64 consecutive k) and MN-major tiles (row = k index, 64 consecutive m/n; one 64-row block per 64 m/n).  Rows are
// filled by 128 loader threads with 16-byte cp.async (zero-fill for padding / ragged tails), which is what lets the
// gather express conv windows, zero padding, nearest upsampling and per-expert row ranges without materialising im2col.
// Warp roles: warps 0-3 load, then run the epilogue (TMEM lane quarter = warp id); warp 4 owns TMEM and issues the MMAs.
// Pipeline: kStages smem stages, full[] barriers (128 loader arrivals after cp.async.wait_group + fence.proxy.async),
// empty[] barriers (tcgen05.commit), one tmem_full barrier for the epilogue.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace es {

enum IgemmMode { FWD = 0, WGRAD_CONV = 1, DENSE_DGRAD = 2, DENSE_WGRAD = 3 };

constexpr int kStages = 4;
constexpr int kLag = 2;
constexpr int kLoaderThreads = 128;
constexpr int kIgemmThreads = 160;
constexpr int kStageABytes = kBM * 128;          // 16 KB
constexpr int kMaxBN = 256;
constexpr int kStageBBytes = kMaxBN * 128;       // 32 KB
constexpr int kStageBytes = kStageABytes + kStageBBytes;
constexpr size_t kIgemmSmem = (size_t)kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;

struct IgemmParams {
  const es_group* grp;
  int n_groups;
  // geometry (conv view); dense layers use Hs=Ws=Hu=Wu=Ho=Wo=KH=KW=1, pad=0
  int Hs, Ws, C, Hu, Wu, Ho, Wo, KH, KW, pad, P;
  int Nout;   // FWD: output channels.  WGRAD_CONV: dy channels.  DENSE_*: N (features)
  int BN;     // N tile (multiple of 32, <= 256)
  int KK;     // reduction length of FWD (KH*KW*C) / M extent of WGRAD_CONV / K (256) of the dense kinds
  int splits;
  unsigned char ymap[64], xmap[64];
  const __nv_bfloat16* a_src;
  const __nv_bfloat16* b_src;
  long b_slot_stride;
  const float* bias;
  long bias_slot_stride;
  void* out;
  long out_slot_stride;
  const int* row_map;
  int* err_flag;
};

// write one 128-byte row (8 x 16 B) of a swizzled tile; `row` is the row index inside its 64/128/256-row block
__device__ __forceinline__ void load_row128(uint32_t block_base, int row, const __nv_bfloat16* src, bool valid) {
  const uint32_t rbase = block_base + (uint32_t)row * 128u;
  const int sw = row & 7;
#pragma unroll
  for (int c = 0; c < 8; ++c) cp_async16(rbase + (uint32_t)((c ^ sw) << 4), src + c * 8, valid);
}

// ------------------------------------------------------------------------------------------------ the kernel
template <int MODE>
__global__ void __launch_bounds__(kIgemmThreads, 1) igemm_kernel(const __grid_constant__ IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---------------- work assignment (uniform per CTA; exits happen before any barrier / TMEM allocation)
  int g = -1, m0 = 0, n0 = 0, kb_begin = 0, kb_end = 0;
  int grp_row_start = 0, grp_rows = 0, slot = 0;
  if (MODE == FWD || MODE == DENSE_DGRAD) {
    int t = blockIdx.x;
    for (int i = 0; i < p.n_groups; ++i) {
      const int tiles = ceil_div(p.grp[i].rows * p.P, kBM);
      if (t < tiles) { g = i; break; }
      t -= tiles;
    }
    if (g < 0) return;
    m0 = t * kBM;
    if (MODE == FWD) {
      n0 = blockIdx.y * p.BN;
      kb_begin = 0;
      kb_end = p.KK / kBK;
    } else {
      const int nkb = p.Nout / kBK, per = ceil_div(nkb, p.splits);
      kb_begin = blockIdx.y * per;
      kb_end = min(nkb, kb_begin + per);
    }
  } else {
    g = (MODE == WGRAD_CONV) ? (int)blockIdx.y / p.splits : (int)blockIdx.y;
    m0 = blockIdx.x * kBM;
    const int ktot = p.grp[g].rows * p.P;
    const int nkb = ceil_div(ktot, kBK);
    if (MODE == WGRAD_CONV) {
      const int per = ceil_div(nkb, p.splits), sp = blockIdx.y % p.splits;
      kb_begin = sp * per;
      kb_end = min(nkb, kb_begin + per);
    } else {
      kb_begin = 0;
      kb_end = nkb;
    }
  }
  if (kb_end <= kb_begin) return;
  grp_row_start = p.grp[g].row_start;
  grp_rows = p.grp[g].rows;
  slot = p.grp[g].slot;
  const int nkb = kb_end - kb_begin;
  const int BN = p.BN;

  // ---------------- shared memory carve-up
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_base = base + kStages * kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * kStages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < BN) tmem_cols <<= 1;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), kLoaderThreads);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    // =========================================================================== LOADERS (128 threads)
    // ---- A-operand per-thread state
    const __nv_bfloat16* a_ptr = p.a_src;
    bool a_valid = false;
    int oy = 0, ox = 0, ky = 0, kx = 0, c0 = 0;       // FWD / DENSE_DGRAD gather cursor
    long a_row_base = 0;
    int a_seg = 0, a_row = 0;                          // MN-major units
    int tap_y = 0, tap_x = 0, a_c0 = 0;                // WGRAD_CONV fixed (tap, channel block) of this thread's segment
    if (MODE == FWD || MODE == DENSE_DGRAD) {
      const int m = m0 + tid;
      a_valid = m < grp_rows * p.P;
      const int sample = a_valid ? m / p.P : 0;
      const int pix = a_valid ? m % p.P : 0;
      oy = pix / p.Wo;
      ox = pix % p.Wo;
      a_row_base = (long)(grp_row_start + sample) * p.Hs * p.Ws;
      // reduction cursor starts at k = kb_begin * 64
      const int k0 = kb_begin * kBK;
      const int tap = k0 / p.C;
      c0 = k0 % p.C;
      ky = tap / p.KW;
      kx = tap % p.KW;
    } else if (MODE == WGRAD_CONV) {
      a_seg = tid >> 6;
      a_row = tid & 63;
      const int kk = m0 + a_seg * 64;
      const int tap = kk / p.C;
      a_c0 = kk % p.C;
      tap_y = tap / p.KW;
      tap_x = tap % p.KW;
    } else {  // DENSE_WGRAD: A = dy[row][Nout], MN tile m0..m0+127
      a_seg = tid >> 6;
      a_row = tid & 63;
    }
    const __nv_bfloat16* wslot = p.b_src + (MODE == FWD || MODE == DENSE_DGRAD ? (long)slot * p.b_slot_stride : 0L);
    const int ktot = grp_rows * p.P;   // MN-major kinds: number of k-rows in the group

    int signalled = 0;
    for (int it = 0; it < nkb; ++it) {
      const int kb = kb_begin + it;
      const int s = it % kStages;
      if (it >= kStages) mbar_wait(empty_bar(s), ((it / kStages) - 1) & 1, p.err_flag, 1);
      const uint32_t sa = base + s * kStageBytes;
      const uint32_t sb = sa + kStageABytes;
      // ------------------------------------------------ A tile
      if (MODE == FWD || MODE == DENSE_DGRAD) {
        const int uy = oy + ky - p.pad, ux = ox + kx - p.pad;
        const bool inb = a_valid && uy >= 0 && uy < p.Hu && ux >= 0 && ux < p.Wu;
        const int sy = inb ? p.ymap[uy] : 0, sx = inb ? p.xmap[ux] : 0;
        const __nv_bfloat16* src = a_ptr + ((a_row_base + (long)sy * p.Ws + sx) * p.C + c0);
        load_row128(sa, tid, inb ? src : a_ptr, inb);
        c0 += kBK;
        if (c0 >= p.C) { c0 = 0; if (++kx == p.KW) { kx = 0; ++ky; } }
      } else if (MODE == WGRAD_CONV) {
        const int pidx = kb * kBK + a_row;                 // pixel index inside the group
        bool inb = pidx < ktot;
        const int sample = inb ? pidx / p.P : 0;
        const int pix = inb ? pidx % p.P : 0;
        const int uy = pix / p.Wo + tap_y - p.pad, ux = pix % p.Wo + tap_x - p.pad;
        inb = inb && uy >= 0 && uy < p.Hu && ux >= 0 && ux < p.Wu;
        const int sy = inb ? p.ymap[uy] : 0, sx = inb ? p.xmap[ux] : 0;
        const __nv_bfloat16* src = a_ptr + (((long)(grp_row_start + sample) * p.Hs * p.Ws + (long)sy * p.Ws + sx) * p.C + a_c0);
        load_row128(sa + a_seg * (64 * 128), a_row, inb ? src : a_ptr, inb);
      } else {  // DENSE_WGRAD
        const int r = kb * kBK + a_row;
        const bool inb = r < ktot;
        const __nv_bfloat16* src = a_ptr + ((long)(grp_row_start + (inb ? r : 0)) * p.Nout + m0 + a_seg * 64);
        load_row128(sa + a_seg * (64 * 128), a_row, src, inb);
      }
      // ------------------------------------------------ B tile
      if (MODE == FWD) {
        for (int r = tid; r < BN; r += kLoaderThreads) {
          const __nv_bfloat16* src = wslot + ((long)(n0 + r) * p.KK + (long)kb * kBK);
          load_row128(sb, r, src, true);
        }
      } else {
        // MN-major dense source S[krow][ld]; tile = BN columns starting at column 0 (BN == full N extent of B)
        const int nseg = BN >> 6;
        for (int u = tid; u < nseg * 64; u += kLoaderThreads) {
          const int seg = u >> 6, row = u & 63;
          const int kr = kb * kBK + row;
          const __nv_bfloat16* src;
          bool inb;
          if (MODE == DENSE_DGRAD) {        // B = w[slot][n][k]: k-row = feature n
            inb = kr < p.Nout;
            src = wslot + ((long)(inb ? kr : 0) * p.KK + seg * 64);
          } else if (MODE == WGRAD_CONV) {  // B = dy[pixel][Nout]
            inb = kr < ktot;
            src = p.b_src + (((long)grp_row_start * p.P + (inb ? kr : 0)) * p.Nout + seg * 64);
          } else {                           // DENSE_WGRAD: B = x[row][KK]
            inb = kr < ktot;
            src = p.b_src + ((long)(grp_row_start + (inb ? kr : 0)) * p.KK + seg * 64);
          }
          load_row128(sb + seg * (64 * 128), row, src, inb);
        }
      }
      cp_async_commit();
      if (it - signalled >= kLag) {
        cp_async_wait<kLag>();
        fence_proxy_async();
        mbar_arrive(full_bar(signalled % kStages));
        ++signalled;
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    while (signalled < nkb) {
      mbar_arrive(full_bar(signalled % kStages));
      ++signalled;
    }

    // =========================================================================== EPILOGUE (same 4 warps)
    mbar_wait(tmem_full_bar, 0, p.err_flag, 3);
    tc_fence_after();
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int m = m0 + tid;   // accumulator row owned by this thread
    uint32_t r[32];
    if (MODE == FWD) {
      const bool ok = m < grp_rows * p.P;
      __nv_bfloat16* yrow = reinterpret_cast<__nv_bfloat16*>(p.out) + (((long)grp_row_start * p.P + m) * p.Nout + n0);
      const float* bias = p.bias ? p.bias + (long)slot * p.bias_slot_stride + n0 : nullptr;
      for (int c = 0; c < BN; c += 32) {
        tmem_ld32(t_lane + c, r);
        if (ok) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(r[j]) + (bias ? __ldg(bias + c + j) : 0.f);
          uint4* dst = reinterpret_cast<uint4*>(yrow + c);
#pragma unroll
          for (int q = 0; q < 4; ++q) dst[q] = pack8(f + 8 * q);
        }
      }
    } else if (MODE == WGRAD_CONV) {
      // accumulator row = packed weight column kk = m0 + tid, accumulator column = output channel n
      float* dw = reinterpret_cast<float*>(p.out) + (long)slot * p.out_slot_stride;
      for (int c = 0; c < BN; c += 32) {
        tmem_ld32(t_lane + c, r);
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(dw + (long)(c + j) * p.KK + m, __uint_as_float(r[j]));
      }
    } else if (MODE == DENSE_DGRAD) {
      const bool ok = m < grp_rows;
      float* dx = reinterpret_cast<float*>(p.out) + (long)(grp_row_start + m) * p.KK;
      for (int c = 0; c < BN; c += 32) {
        tmem_ld32(t_lane + c, r);
        if (ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(dx + c + j, __uint_as_float(r[j]));
        }
      }
    } else {  // DENSE_WGRAD: row = feature n (packed order) -> reference row row_map[n]
      const int n = m;
      const int nref = p.row_map ? p.row_map[n] : n;
      float* dw = reinterpret_cast<float*>(p.out) + (long)slot * p.out_slot_stride + (long)nref * p.KK;
      for (int c = 0; c < BN; c += 32) {
        tmem_ld32(t_lane + c, r);
        float4* dst = reinterpret_cast<float4*>(dw + c);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          dst[q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                               __uint_as_float(r[4 * q + 3]));
      }
    }
    tc_fence_before();
  } else {
    // =========================================================================== MMA ISSUER (warp 4)
    constexpr bool A_MN = (MODE == WGRAD_CONV || MODE == DENSE_WGRAD);
    constexpr bool B_MN = (MODE != FWD);
    const uint32_t idesc = make_idesc(BN, A_MN, B_MN);
    for (int it = 0; it < nkb; ++it) {
      const int s = it % kStages;
      mbar_wait(full_bar(s), (it / kStages) & 1, p.err_flag, 2);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = base + s * kStageBytes;
        const uint32_t sb = sa + kStageABytes;
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          // K-major: advance 16 k = 32 B inside the swizzle atom.  MN-major: advance 16 k-rows = 2048 B.
          const uint64_t ad = A_MN ? make_desc(sa + k * 2048, 64 * 128, 1024) : make_desc(sa + k * 32, 16, 1024);
          const uint64_t bd = B_MN ? make_desc(sb + k * 2048, 64 * 128, 1024) : make_desc(sb + k * 32, 16, 1024);
          umma_bf16(tmem_base, ad, bd, idesc, (it | k) ? 1u : 0u);
        }
        umma_commit(empty_bar(s));
        if (it == nkb - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------ host side
static int* err_flag_ptr() { return pipeline_err_flag(); }

void fill_maps(IgemmParams& p) {
  // torch 'nearest': src = min(floor(dst * (in/out as float)), in-1)
  const float sy = (float)p.Hs / (float)p.Hu, sx = (float)p.Ws / (float)p.Wu;
  for (int i = 0; i < 64; ++i) {
    int y = (int)floorf((float)i * sy), x = (int)floorf((float)i * sx);
    p.ymap[i] = (unsigned char)(y < p.Hs - 1 ? y : p.Hs - 1);
    p.xmap[i] = (unsigned char)(x < p.Ws - 1 ? x : p.Ws - 1);
  }
}

int check_geom(const es_conv_geom* g) {
  if (!g) return 0;
  if (g->C <= 0 || g->C % 64 != 0) return 0;
  if (g->Hu > 64 || g->Wu > 64 || g->Hu < g->Hs || g->Wu < g->Ws) return 0;
  if (g->Ho != g->Hu + 2 * g->pad - g->KH + 1 || g->Wo != g->Wu + 2 * g->pad - g->KW + 1) return 0;
  if (g->Ho <= 0 || g->Wo <= 0) return 0;
  return 1;
}

void geom_to_params(const es_conv_geom* g, IgemmParams& p) {
  p.Hs = g->Hs; p.Ws = g->Ws; p.C = g->C; p.Hu = g->Hu; p.Wu = g->Wu; p.Ho = g->Ho; p.Wo = g->Wo;
  p.KH = g->KH; p.KW = g->KW; p.pad = g->pad; p.P = g->Ho * g->Wo;
  p.KK = g->KH * g->KW * g->C;
  fill_maps(p);
}

template <int MODE>
int launch_igemm(const IgemmParams& p, dim3 grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(igemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kIgemmSmem);
    if (e != cudaSuccess) { set_error(std::string("igemm smem attribute: ") + cudaGetErrorString(e)); return ES_ERR_CUDA; }
    attr_set = true;
  }
  igemm_kernel<MODE><<<grid, kIgemmThreads, kIgemmSmem, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error(std::string("igemm launch: ") + cudaGetErrorString(e)); return ES_ERR_CUDA; }
  return ES_OK;
}


}  // namespace es

using namespace es;

namespace es {
int dense_wgrad_tma(const void* dy, const void* x, void* xpad, float* dw, long dw_slot_stride, int N, int K, const int32_t* row_map,
                    const es_group* grp, int n_groups, int total_rows, cudaStream_t st);
int dense_dgrad_tma(const void* dy, const void* w, float* dx, int N, int K, const es_group* grp, int n_groups, int total_rows,
                    cudaStream_t st);
}

extern "C" int es_dense_dgrad(const void* dy, const void* w, float* dx, int N, int K, const es_group* grp,
                              int n_groups, int total_rows, void* stream) {
  ES_REQUIRE(dy && w && dx && grp, "null pointer");
  ES_REQUIRE(N % 64 == 0 && K % 64 == 0 && K <= 256, "need N % 64 == 0 and K in {64,128,192,256}");
  ES_REQUIRE(n_groups >= 1 && n_groups <= kMaxGroups && total_rows > 0, "bad group count / rows");
  {   // TMA-fed kernel (dense_tma.cu) for the fc2 shape; 1 = shape not covered -> the gather kernel below
    const int rc = dense_dgrad_tma(dy, w, dx, N, K, grp, n_groups, total_rows, as_stream(stream));
    if (rc != 1) { if (rc != ES_OK) set_error("es_dense_dgrad: TMA kernel launch failed"); return rc; }
  }
  IgemmParams p{};
  p.Hs = p.Ws = p.Hu = p.Wu = p.Ho = p.Wo = p.KH = p.KW = 1; p.pad = 0; p.P = 1;
  p.C = N;   // the gather walks the reduction dimension n in 64-wide blocks
  fill_maps(p);
  p.grp = grp; p.n_groups = n_groups;
  p.Nout = N; p.KK = K; p.BN = K;
  const int mt = ceil_div(total_rows, kBM) + n_groups;
  int splits = ceil_div(2 * 148, mt);
  if (splits > N / kBK) splits = N / kBK;
  if (splits < 1) splits = 1;
  p.splits = splits;
  p.a_src = (const __nv_bfloat16*)dy; p.b_src = (const __nv_bfloat16*)w; p.b_slot_stride = (long)N * K;
  p.out = dx; p.err_flag = err_flag_ptr();
  return launch_igemm<DENSE_DGRAD>(p, dim3(mt, splits), as_stream(stream));
}

extern "C" int es_dense_wgrad(const void* dy, const void* x, float* dw, long dw_slot_stride, int N, int K,
                              const int32_t* row_map, const es_group* grp, int n_groups, int total_rows, void* scratch,
                              void* stream) {
  ES_REQUIRE(dy && x && dw && grp, "null pointer");
  ES_REQUIRE(N % kBM == 0 && K % 64 == 0 && K <= 256, "need N % 128 == 0 and K in {64,...,256}");
  ES_REQUIRE(n_groups >= 1 && n_groups <= kMaxGroups && total_rows > 0, "bad group count / rows");
  {   // TMA-fed kernel (dense_tma.cu); needs the caller's scratch for the zero-padded per-group copy of x
    const int rc = dense_wgrad_tma(dy, x, scratch, dw, dw_slot_stride, N, K, row_map, grp, n_groups, total_rows, as_stream(stream));
    if (rc != 1) { if (rc != ES_OK) set_error("es_dense_wgrad: TMA kernel launch failed"); return rc; }
  }
  IgemmParams p{};
  p.Hs = p.Ws = p.Hu = p.Wu = p.Ho = p.Wo = p.KH = p.KW = 1; p.pad = 0; p.P = 1; p.C = 64;
  fill_maps(p);
  p.grp = grp; p.n_groups = n_groups;
  p.Nout = N; p.KK = K; p.BN = K; p.splits = 1;
  p.a_src = (const __nv_bfloat16*)dy; p.b_src = (const __nv_bfloat16*)x;
  p.out = dw; p.out_slot_stride = dw_slot_stride; p.row_map = row_map; p.err_flag = err_flag_ptr();
  return launch_igemm<DENSE_WGRAD>(p, dim3(N / kBM, n_groups), as_stream(stream));
}
