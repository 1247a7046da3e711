// fc2 backward at the HBM roof: TMA-fed tcgen05 kernels for the two dense products behind the generator's big linear layer
// (proton: 92 160 x 256 per expert, 47 MB of bf16 weights and a 94 MB fp32 weight gradient per expert).  Both are pure
// streaming problems — 0.75 / 1.13 GB of compulsory HBM traffic at batch 1024 — that the round-1 kernels (per-thread
// cp.async for both operands, igemm_tc.cu) ran at 0.25 of the measured copy rate.
//   MODE 0  weight gradient  dw[slot][row_map[f]][k] = sum_r dy[r][f] * x[r][k]        D tile = 128 f x 256 k, reduction over the
//           group's rows in blocks of 64.  dy tiles are MN-major TMA boxes {64 f, 64 rows}; x comes from a zero-padded
//           per-group copy (xpad[g][RP][K], rows past the group's end are zero) so that the dy rows a box drags in from the
//           NEXT group multiply zeros — TMA cannot zero-fill at a group boundary.  Epilogue: direct fp32 stores, one 1 KB row
//           per TMEM lane (no atomics: every (f, k) has exactly one producer).
//   MODE 1  data gradient    dx[r][k] += sum_f dy[r][f] * w[slot][f][k]                 D tile = 128 rows x 256 k, reduction over f
//           split across CTAs; dy tiles are K-major boxes {64 f, 128 rows}, w tiles MN-major boxes {64 k, 64 f}; rows past
//           the group's end are computed with the wrong expert's weights and masked in the epilogue (RED into dx).
// 6 warps: TMA producer, MMA issuer / TMEM owner, 4 epilogue warps; 4 stages of 48 KB; accumulators double-buffered in TMEM.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace es {
namespace {

constexpr int kDStages = 4;
constexpr int kDStageA = 128 * 128;              // 16 KB: 128 f x 64 rows (MODE 0) or 128 rows x 64 f (MODE 1)
constexpr int kDStageB = 256 * 128;              // 32 KB: 256 k x 64 reduction rows
constexpr int kDStage = kDStageA + kDStageB;
constexpr int kDStagingPitch = 36;                // floats per staged row: 32 + 4 (16-byte accesses of a quarter-warp stay conflict-free)
constexpr size_t kDSmem = (size_t)kDStages * kDStage + 1024 + 2048 + 4 * 32 * kDStagingPitch * 4;

struct DenseParams {
  const es_group* grp;
  int n_groups, N, K, RP, splits, tiles_m;       // N features, K = 256 inner width, RP = padded rows per group (MODE 0)
  const int32_t* row_map;
  float* out;
  long out_slot_stride;
  int* err_flag;
};

struct DTile { int g, slot, rows, row_start, t0, kb0, kb1; };

// MODE 0: unit = (group, feature tile);  MODE 1: unit = (M tile of a group, split of the feature range)
template <int MODE>
__device__ __forceinline__ bool dense_decode(int u, const DenseParams& p, const es_group* s_grp, const int* s_tiles, DTile& t) {
  if (MODE == 0) {
    const int nft = p.N / kBM;
    t.g = u / nft;
    if (t.g >= p.n_groups) return false;
    const es_group G = s_grp[t.g];
    t.slot = G.slot; t.rows = G.rows; t.row_start = G.row_start;
    t.t0 = (u - t.g * nft) * kBM;                 // first feature
    t.kb0 = 0; t.kb1 = ceil_div(G.rows, kBK);
    return G.rows > 0;
  }
  int mt = u / p.splits;
  const int sp = u - mt * p.splits;
  for (int i = 0; i < p.n_groups; ++i) {
    const int n = s_tiles[i];
    if (mt < n) {
      const es_group G = s_grp[i];
      t.g = i; t.slot = G.slot; t.rows = G.rows; t.row_start = G.row_start;
      t.t0 = mt * kBM;                            // first row of the tile inside the group
      const int nkb = p.N / kBK, per = ceil_div(nkb, p.splits);
      t.kb0 = sp * per; t.kb1 = min(nkb, t.kb0 + per);
      return t.kb1 > t.kb0;
    }
    mt -= n;
  }
  return false;
}

template <int MODE>
__global__ void __launch_bounds__(192, 1)
dense_tma_kernel(const __grid_constant__ DenseParams p, const __grid_constant__ CUtensorMap tmap_a,
                 const __grid_constant__ CUtensorMap tmap_b) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_base = base + kDStages * kDStage;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (16 + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (18 + b); };
  uint8_t* gen = smem_raw + (bar_base - raw);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + 8 * 20);
  const uint32_t tmem_slot = bar_base + 8u * 20;
  int* s_tiles = reinterpret_cast<int*>(gen + 256);
  es_group* s_grp = reinterpret_cast<es_group*>(gen + 512);
  constexpr int BN = 256;

  if (tid < p.n_groups) {
    const es_group gq = p.grp[tid];
    s_grp[tid] = gq;
    s_tiles[tid] = ceil_div(gq.rows, kBM);
  }
  if (tid == 0) {
    for (int s = 0; s < kDStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_b); }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  int total = 0;
  if (MODE == 0) total = p.n_groups * (p.N / kBM);
  else { for (int i = 0; i < p.n_groups; ++i) total += s_tiles[i]; total *= p.splits; }

  if (warp == 0) {
    // =========================================================================== TMA PRODUCER
    if (lane == 0) {
      uint32_t it = 0;
      for (int u = blockIdx.x; u < total; u += gridDim.x) {
        DTile t;
        if (!dense_decode<MODE>(u, p, s_grp, s_tiles, t)) continue;
        for (int kb = t.kb0; kb < t.kb1; ++kb, ++it) {
          const int s = it % kDStages;
          if (it >= (uint32_t)kDStages) mbar_wait(empty_bar(s), ((it / kDStages) - 1) & 1, p.err_flag, 4);
          mbar_arrive_expect_tx(full_bar(s), (uint32_t)kDStage);
          const uint32_t sa = base + s * kDStage, sb = sa + kDStageA;
          if (MODE == 0) {
            // A: dy[rows kb*64.., features t0..t0+127] as two MN-major boxes {64 f, 64 rows}
            tma_load_2d(sa, &tmap_a, t.t0, t.row_start + kb * kBK, full_bar(s));
            tma_load_2d(sa + 8192u, &tmap_a, t.t0 + 64, t.row_start + kb * kBK, full_bar(s));
            // B: xpad[g][rows kb*64..][k] as four MN-major boxes {64 k, 64 rows}
            for (int sg = 0; sg < 4; ++sg) tma_load_2d(sb + sg * 8192u, &tmap_b, sg * 64, t.g * p.RP + kb * kBK, full_bar(s));
          } else {
            // A: dy[rows t0.., features kb*64..] as one K-major box {64 f, 128 rows}
            tma_load_2d(sa, &tmap_a, kb * kBK, t.row_start + t.t0, full_bar(s));
            // B: w[slot][features kb*64..][k] as four MN-major boxes {64 k, 64 f}
            for (int sg = 0; sg < 4; ++sg) tma_load_2d(sb + sg * 8192u, &tmap_b, sg * 64, t.slot * p.N + kb * kBK, full_bar(s));
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================================================================== MMA ISSUER
    const uint32_t idesc = make_idesc_m(BN, kBM, MODE == 0, true);
    uint32_t it = 0, tcount = 0;
    for (int u = blockIdx.x; u < total; u += gridDim.x) {
      DTile t;
      if (!dense_decode<MODE>(u, p, s_grp, s_tiles, t)) continue;
      const uint32_t buf = tcount & 1;
      if (tcount >= 2) mbar_wait(tempty_bar(buf), ((tcount >> 1) - 1) & 1, p.err_flag, 5);
      tc_fence_after();
      const uint32_t tacc = tmem_base + buf * (uint32_t)BN;
      for (int kb = t.kb0; kb < t.kb1; ++kb, ++it) {
        const int s = it % kDStages;
        mbar_wait(full_bar(s), (it / kDStages) & 1, p.err_flag, 2);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = base + s * kDStage, sb = sa + kDStageA;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // MN-major operands advance 16 reduction rows = 2048 B per step (LBO = next 64-wide block, SBO = 8 rows);
            // the K-major A of MODE 1 advances 32 B inside its 128-byte rows
            const uint64_t ad = MODE == 0 ? make_desc(sa + k * 2048, 8192, 1024) : make_desc(sa + k * 32, 16, 1024);
            const uint64_t bd = make_desc(sb + k * 2048, 8192, 1024);
            umma_bf16(tacc, ad, bd, idesc, (kb > t.kb0 || k) ? 1u : 0u);
          }
          umma_commit(empty_bar(s));
          if (kb == t.kb1 - 1) umma_commit(tfull_bar(buf));
        }
        __syncwarp();
      }
      ++tcount;
    }
    tc_fence_before();
  } else {
    // =========================================================================== EPILOGUE (warps 2-5; TMEM lane quarter = warp % 4).
    // (Eight epilogue warps, two per quarter splitting the columns, were measured: 0.507 vs 0.441 ms — not used.)
    const int q = warp & 3, c_lo = 0, c_hi = BN;
    uint32_t tcount = 0;
    for (int u = blockIdx.x; u < total; u += gridDim.x) {
      DTile t;
      if (!dense_decode<MODE>(u, p, s_grp, s_tiles, t)) continue;
      const uint32_t buf = tcount & 1;
      mbar_wait(tfull_bar(buf), (tcount >> 1) & 1, p.err_flag, 3);
      tc_fence_after();
      const uint32_t t_lane = tmem_base + buf * (uint32_t)BN + ((uint32_t)(q * 32) << 16);
      const int m = t.t0 + q * 32 + lane;          // MODE 0: feature;  MODE 1: row inside the group
      uint32_t r[32];
      if (MODE == 0) {
        const int orow = p.row_map ? p.row_map[m] : m;
        float* o = p.out + (long)t.slot * p.out_slot_stride + (long)orow * p.K;
        // A lane owns a whole output row (TMEM lane = feature), so storing straight from registers makes every STG touch 32
        // different 1 KB rows (ncu: lg_throttle 14 stall cycles per issue).  The 32 x 32 chunk is transposed through a
        // per-warp shared-memory tile instead: 8 lanes then write one row's 128 contiguous bytes, 4 rows per instruction.
        float* stg = reinterpret_cast<float*>(gen + 2048) + q * (32 * kDStagingPitch);
        const unsigned long long oaddr = reinterpret_cast<unsigned long long>(o);
        for (int c = c_lo; c < c_hi; c += 32) {
          tmem_ld32(t_lane + c, r);
          float4* w4 = reinterpret_cast<float4*>(stg + lane * kDStagingPitch);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            w4[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = i * 4 + (lane >> 3), piece = lane & 7;
            const unsigned long long ra = __shfl_sync(0xffffffffu, oaddr, row);
            const float4 v = *reinterpret_cast<const float4*>(stg + row * kDStagingPitch + piece * 4);
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(ra) + c + piece * 4) = v;
          }
          __syncwarp();
        }
      } else {
        // split-K partial sums: transposed through the same per-warp tile, then one 16-byte vector RED per 4 columns — 8 lanes
        // cover a row's 128 contiguous bytes (32 scalar REDs per lane, each to its own 1 KB row, before)
        const bool ok = m < t.rows;
        float* stg = reinterpret_cast<float*>(gen + 2048) + q * (32 * kDStagingPitch);
        const unsigned long long oaddr = ok ? reinterpret_cast<unsigned long long>(p.out + (long)(t.row_start + m) * p.K) : 0ull;
        for (int c = c_lo; c < c_hi; c += 32) {
          tmem_ld32(t_lane + c, r);
          float4* w4 = reinterpret_cast<float4*>(stg + lane * kDStagingPitch);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            w4[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = i * 4 + (lane >> 3), piece = lane & 7;
            const unsigned long long ra = __shfl_sync(0xffffffffu, oaddr, row);
            const float4 v = *reinterpret_cast<const float4*>(stg + row * kDStagingPitch + piece * 4);
            if (ra) atomicAdd(reinterpret_cast<float4*>(reinterpret_cast<float*>(ra) + c + piece * 4), v);
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));
      ++tcount;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// x [rows][K] bf16 -> xpad[g][RP][K]: the group's rows, then zeros up to the next multiple of 64 rows
__global__ void __launch_bounds__(256)
pad_group_rows_kernel(const __nv_bfloat16* __restrict__ x, const es_group* __restrict__ grp, int n_groups, int K8, int RP,
                      __nv_bfloat16* __restrict__ xpad) {
  const int g = blockIdx.y;
  const es_group G = grp[g];
  const int rpad = ceil_div(G.rows, kBK) * kBK;
  const uint4* src = reinterpret_cast<const uint4*>(x) + (size_t)G.row_start * K8;
  uint4* dst = reinterpret_cast<uint4*>(xpad) + (size_t)g * RP * K8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rpad * K8; i += gridDim.x * blockDim.x)
    dst[i] = i < G.rows * K8 ? src[i] : make_uint4(0, 0, 0, 0);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn dense_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

bool encode2d(CUtensorMap* m, const void* base, long cols, long rows, int box_cols, int box_rows) {
  EncodeTiledFn enc = dense_encode_fn();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// -> ES_OK when launched; 1 when the shape is not covered (caller falls back to the gather kernels of igemm_tc.cu)
int dense_wgrad_tma(const void* dy, const void* x, void* xpad, float* dw, long dw_slot_stride, int N, int K, const int32_t* row_map,
                    const es_group* grp, int n_groups, int total_rows, cudaStream_t st) {
  static const bool on = [] { const char* e = getenv("ES_DENSE_TMA"); return !(e && e[0] == '0'); }();
  if (!on || !xpad || K != 256 || N % kBM != 0 || n_groups > 64) return 1;
  const int RP = ceil_div(total_rows, kBK) * kBK;
  alignas(64) CUtensorMap ta, tb;
  if (!encode2d(&ta, dy, N, total_rows, 64, 64) || !encode2d(&tb, xpad, K, (long)n_groups * RP, 64, 64)) return 1;
  pad_group_rows_kernel<<<dim3(ceil_div(RP * (K / 8), 256 * 4), n_groups), 256, 0, st>>>(
      (const __nv_bfloat16*)x, grp, n_groups, K / 8, RP, (__nv_bfloat16*)xpad);
  DenseParams p{};
  p.grp = grp; p.n_groups = n_groups; p.N = N; p.K = K; p.RP = RP; p.splits = 1; p.row_map = row_map; p.out = dw;
  p.out_slot_stride = dw_slot_stride; p.err_flag = pipeline_err_flag();
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(dense_tma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDSmem) != cudaSuccess) return 1;
    if (cudaFuncSetAttribute(dense_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDSmem) != cudaSuccess) return 1;
    attr = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int units = n_groups * (N / kBM);
  dense_tma_kernel<0><<<units < sms ? units : sms, 192, kDSmem, st>>>(p, ta, tb);
  return cudaGetLastError() == cudaSuccess ? ES_OK : ES_ERR_CUDA;
}

int dense_dgrad_tma(const void* dy, const void* w, float* dx, int N, int K, const es_group* grp, int n_groups, int total_rows,
                    cudaStream_t st) {
  static const bool on = [] { const char* e = getenv("ES_DENSE_TMA"); return !(e && e[0] == '0'); }();
  if (!on || K != 256 || N % kBK != 0 || n_groups > 64) return 1;
  alignas(64) CUtensorMap ta, tb;
  if (!encode2d(&ta, dy, N, total_rows, 64, 128) || !encode2d(&tb, w, K, (long)64 * N, 64, 64)) return 1;
  DenseParams p{};
  p.grp = grp; p.n_groups = n_groups; p.N = N; p.K = K; p.out = dx; p.err_flag = pipeline_err_flag();
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int mt = ceil_div(total_rows, kBM) + n_groups;             // upper bound on M tiles
  int splits = ceil_div(2 * sms, mt);
  if (splits > N / kBK) splits = N / kBK;
  if (splits < 1) splits = 1;
  p.splits = splits;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(dense_tma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDSmem) != cudaSuccess) return 1;
    if (cudaFuncSetAttribute(dense_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDSmem) != cudaSuccess) return 1;
    attr = true;
  }
  const int units = mt * splits;
  dense_tma_kernel<1><<<units < sms ? units : sms, 192, kDSmem, st>>>(p, ta, tb);
  return cudaGetLastError() == cudaSuccess ? ES_OK : ES_ERR_CUDA;
}

}  // namespace es
