// K2 (forward kind) — persistent, warp-specialised grouped implicit GEMM on tcgen05:  y[pix, n] = sum_k im2col(x)[pix, k] w[n, k].
// Used by the generator's forward convolutions (nearest upsample folded into the gather), by their data gradients (same
// kernel, transposed + flipped packed weights) and by fc2.
//
// Why this shape (r01 ncu of the first, non-persistent kernel: L1TEX 82 % busy, L2 50 %, tensor pipe 16 %): every operand
// byte went through per-thread 16-byte cp.async, 32 different lines per warp instruction.  Here
//   * B (weights, K-major [slots*N][KK] bf16) arrives by TMA — one elected thread, 128B-swizzled box {64 k, BN n}, no
//     L1TEX traffic at all;
//   * A (im2col rows: one output pixel x 64 channels = one 128-byte line) is still a gather, because the nearest upsample
//     and the per-expert ragged row ranges are not expressible as a TMA box, but 8 consecutive lanes now copy ONE line
//     (4 full lines per warp instruction), through L1 (.ca), and the reduction runs channel-block-major / tap-minor so
//     the KH*KW shifted re-reads of a source line hit L1 instead of L2;
//   * the CTA is persistent (one per SM, static tile striding) with a double-buffered TMEM accumulator, so the epilogue
//     of tile i (TMEM -> registers -> +bias -> bf16 -> global) overlaps the main loop of tile i+1.
// Warp roles: 0-7 A gather, 8 TMA producer (B), 9 MMA issuer + TMEM owner, 10-13 epilogue (TMEM lane quarter = warp % 4).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace es {

namespace {

constexpr int kFStages = 4;
constexpr int kFLag = 2;
constexpr int kFLoaders = 128;
constexpr int kFThreads = 320;
// forward kernel: 8 gather warps (r01 ncu: tensor 21-39 %, L1TEX <= 50 %, L2 <= 27 % — the gather was bound by the latency of
// its own address-generation instruction stream per warp, not by any memory pipe), then TMA, MMA and 4 epilogue warps
constexpr int kGW = 8;
constexpr int kGLoaders = kGW * 32;
constexpr int kGThreads = (kGW + 6) * 32;
constexpr int kFStageA = kBM * 128;            // 16 KB
constexpr int kFStageB = 256 * 128;            // 32 KB (BN <= 256)
constexpr int kFStage = kFStageA + kFStageB;
constexpr int kFMaxGroups = 64;
constexpr size_t kFSmem = (size_t)kFStages * kFStage + 1024 /*align*/ + 2048 /*barriers + tables*/ + 4096 /*bias rows of the epilogue warps*/ + 4 * 32 * 80 /*their store-transposition tiles*/;

// The reduction is a TAP TABLE: tap t reads the (virtually upsampled) source at (oy*my + tdy[t], ox*mx + tdx[t]) and its C
// weights start at column tkoff[t] of the packed weight row.  A plain conv is tdy = ky - pad, tdx = kx - pad, tkoff = t*C,
// my = mx = 1.  The phase-folded x2-upsample convs and their combined data gradient use other tables (see
// es_igemm_taps_fwd).  Output pixel (a, b) of the M-space grid [Ho, Wo] is stored at ((a*o_my + o_oy)*Wo_full + b*o_mx + o_ox).
struct FwdParams {
  const es_group* grp;
  int n_groups;
  int Hs, Ws, C, Hu, Wu, Ho, Wo, P;
  int n_taps, my, mx;
  signed char tdy[32], tdx[32];
  int tkoff[32];
  int o_my, o_oy, o_mx, o_ox, Wo_full, P_full;
  int Nout, BN, KK, n_tiles_n;
  unsigned char ymap[64], xmap[64];
  const __nv_bfloat16* a_src;
  const float* bias;
  long bias_slot_stride;
  __nv_bfloat16* out;
  float* pair_sums;            // optional [total rows][Nout/2][2]: per-sample channel-pair (sum, sum of squares) of the stored outputs
  int epi_staged;              // pair kernel: bias row prefetched to shared memory + stores transposed through it (the dense product)
  int* err_flag;
};

struct TileInfo {
  int g, m0, n0, rows, row_start, slot;
};

// tile t -> (group, m0, n0).  s_tiles[i] = number of M tiles of group i (shared memory).
__device__ __forceinline__ bool decode_tile(int t, const FwdParams& p, const int* s_tiles, const es_group* s_grp, int tile_m,
                                            TileInfo& ti) {
  int mt = t / p.n_tiles_n;
  const int nt = t - mt * p.n_tiles_n;
  for (int i = 0; i < p.n_groups; ++i) {
    const int n = s_tiles[i];
    if (mt < n) {
      ti.g = i; ti.m0 = mt * tile_m; ti.n0 = nt * p.BN;
      ti.rows = s_grp[i].rows; ti.row_start = s_grp[i].row_start; ti.slot = s_grp[i].slot;
      return true;
    }
    mt -= n;
  }
  return false;
}

// MT  = 128-row M sub-tiles per CTA tile (1 or 2).  Two sub-tiles share every weight (B) tile, which halves the L2->SM
//       weight traffic per FLOP — the bound of the N = 256 layers (r01: B alone was 7.4 TB/s of L2 reads at 948 TFLOP/s).
// G4  = A rows arrive by TMA tile::gather4 (four 128-byte rows per instruction, row indices computed per tap, -1 = zero
//       padding) issued by one warp, instead of per-thread cp.async.  cp.async moves 64 B/clk/SM through L1TEX, which is
//       exactly what a 128 x 128 x 64 k-block needs at tensor peak, so the N <= 128 layers were gather-bound at ~50 %;
//       gather4 bypasses L1TEX (at the price of re-reading shifted taps from L2 instead of L1).
template <int MT, bool G4>
__global__ void __launch_bounds__(kGThreads, 1)
igemm_fwd_kernel(const __grid_constant__ FwdParams p, const __grid_constant__ CUtensorMap tmap_w,
                 const __grid_constant__ CUtensorMap tmap_a) {
  constexpr int kStages = MT == 1 ? 4 : 3;
  constexpr int kStageA = MT * kFStageA;
  constexpr int kStage = kStageA + kFStageB;
  constexpr int kTileM = MT * kBM;
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_base = base + kStages * kStage;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (4 + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (8 + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (10 + b); };
  uint8_t* gen = smem_raw + (bar_base - raw);                 // generic pointer to the barrier/table area
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + 8 * 12);
  const uint32_t tmem_slot = bar_base + 8u * 12;
  int* s_tiles = reinterpret_cast<int*>(gen + 128);            // [64]
  es_group* s_grp = reinterpret_cast<es_group*>(gen + 384);    // [64] x 16 B
  unsigned char* s_ymap = gen + 384 + 1024;                    // 64
  unsigned char* s_xmap = s_ymap + 64;                         // 64
  signed char* s_tdy = reinterpret_cast<signed char*>(s_xmap + 64);   // 32
  signed char* s_tdx = s_tdy + 32;                             // 32
  int* s_tkoff = reinterpret_cast<int*>(s_tdx + 32);           // 32 ints

  const int BN = p.BN;
  const int acc_cols = MT * BN;                                // TMEM columns of one accumulator set
  const uint32_t nbuf = 2 * acc_cols <= 512 ? 2u : 1u;         // double-buffer the accumulators when TMEM allows
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < (int)nbuf * acc_cols) tmem_cols <<= 1;

  // ---- one-time setup
  if (tid < p.n_groups) {
    const es_group gq = p.grp[tid];
    s_grp[tid] = gq;
    s_tiles[tid] = ceil_div(gq.rows * p.P, kTileM);
  }
  if (tid < 64) { s_ymap[tid] = p.ymap[tid]; s_xmap[tid] = p.xmap[tid]; }
  if (tid < 32) { s_tdy[tid] = p.tdy[tid]; s_tdx[tid] = p.tdx[tid]; s_tkoff[tid] = p.tkoff[tid]; }
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), G4 ? 2 : kGLoaders + 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kGW && lane == 0) { tma_prefetch_desc(&tmap_w); if (G4) tma_prefetch_desc(&tmap_a); }
  if (warp == kGW + 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  int total_tiles = 0;
  for (int i = 0; i < p.n_groups; ++i) total_tiles += s_tiles[i];
  total_tiles *= p.n_tiles_n;
  const int taps = p.n_taps;
  const int cblks = p.C / kBK;
  const int nkb = taps * cblks;

  if (warp < kGW) {
    if (!G4) {
      // ========================================================================= A GATHER by cp.async (128 threads)
      // lane group of 8 threads copies one 128-byte row; thread handles rows (tid>>3) + 16*j, chunk tid&7
      constexpr int RPT = MT * kBM / (kGLoaders / 8);       // rows per thread: 8 lanes per row
      constexpr int RSTEP = kGLoaders / 8;
      const int chunk = tid & 7, rsub = tid >> 3;
      uint32_t it = 0, signalled = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        TileInfo ti;
        decode_tile(tile, p, s_tiles, s_grp, kTileM, ti);
        int r_base[RPT];    // pixel index of the sample's first source pixel, or -1 for rows past the group's end
        int r_oyx[RPT];     // oy << 8 | ox
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
          const int m = ti.m0 + rsub + RSTEP * j;
          const bool valid = m < ti.rows * p.P;
          const int sample = valid ? m / p.P : 0;
          const int pix = valid ? m - sample * p.P : 0;
          const int oy = pix / p.Wo;
          r_oyx[j] = (oy << 8) | (pix - oy * p.Wo);
          r_base[j] = valid ? (ti.row_start + sample) * p.Hs * p.Ws : -1;
        }
        int cb = 0, tap = 0;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kStages;
          if (it >= kStages) mbar_wait(empty_bar(s), ((it / kStages) - 1) & 1, p.err_flag, 1);
          const uint32_t sa = base + s * kStage;
          const int c0 = cb * kBK + chunk * 8;
          const int ty = s_tdy[tap], tx = s_tdx[tap];
#pragma unroll
          for (int j = 0; j < RPT; ++j) {
            const int r = rsub + RSTEP * j;
            const int uy = (r_oyx[j] >> 8) * p.my + ty, ux = (r_oyx[j] & 255) * p.mx + tx;
            const bool inb = r_base[j] >= 0 && uy >= 0 && uy < p.Hu && ux >= 0 && ux < p.Wu;
            const int sy = inb ? s_ymap[uy] : 0, sx = inb ? s_xmap[ux] : 0;
            const __nv_bfloat16* src = p.a_src + (inb ? ((long)(r_base[j] + sy * p.Ws + sx) * p.C + c0) : 0L);
            cp_async16_ca(sa + (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4), src, inb);
          }
          cp_async_commit();
          if (++tap == taps) { tap = 0; ++cb; }
          if (it - signalled >= (uint32_t)kFLag) {
            cp_async_wait<kFLag>();
            fence_proxy_async();
            mbar_arrive(full_bar(signalled % kStages));
            ++signalled;
          }
        }
      }
      cp_async_wait<0>();
      fence_proxy_async();
      while (signalled < it) {
        mbar_arrive(full_bar(signalled % kStages));
        ++signalled;
      }
    } else if (warp == 0) {
      // ========================================================================= A GATHER by TMA gather4 (one warp)
      // lane l owns tile rows 4l .. 4l+3 of every 128-row sub-tile: one gather4 (512 B) per sub-tile per k-block
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        TileInfo ti;
        decode_tile(tile, p, s_tiles, s_grp, kTileM, ti);
        int r_base[4 * MT], r_oyx[4 * MT];
#pragma unroll
        for (int j = 0; j < 4 * MT; ++j) {
          const int m = ti.m0 + (j >> 2) * kBM + lane * 4 + (j & 3);
          const bool valid = m < ti.rows * p.P;
          const int sample = valid ? m / p.P : 0;
          const int pix = valid ? m - sample * p.P : 0;
          const int oy = pix / p.Wo;
          r_oyx[j] = (oy << 8) | (pix - oy * p.Wo);
          r_base[j] = valid ? (ti.row_start + sample) * p.Hs * p.Ws : -1;
        }
        int cb = 0, tap = 0;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kStages;
          if (it >= kStages) mbar_wait(empty_bar(s), ((it / kStages) - 1) & 1, p.err_flag, 1);
          const uint32_t sa = base + s * kStage;
          if (lane == 0) mbar_arrive_expect_tx(full_bar(s), (uint32_t)kStageA);
          __syncwarp();
          const int ty = s_tdy[tap], tx = s_tdx[tap];
          int idx[4 * MT];
#pragma unroll
          for (int j = 0; j < 4 * MT; ++j) {
            const int uy = (r_oyx[j] >> 8) * p.my + ty, ux = (r_oyx[j] & 255) * p.mx + tx;
            const bool inb = r_base[j] >= 0 && uy >= 0 && uy < p.Hu && ux >= 0 && ux < p.Wu;
            idx[j] = inb ? r_base[j] + s_ymap[uy] * p.Ws + s_xmap[ux] : -1;     // -1: out of bounds -> zero rows
          }
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            tma_gather4(sa + mt * kFStageA + lane * 512u, &tmap_a, cb * kBK, idx[4 * mt], idx[4 * mt + 1], idx[4 * mt + 2],
                        idx[4 * mt + 3], full_bar(s));
          if (++tap == taps) { tap = 0; ++cb; }
        }
      }
    }
  } else if (warp == kGW) {
    // =========================================================================== TMA PRODUCER (weights)
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        TileInfo ti;
        decode_tile(tile, p, s_tiles, s_grp, kTileM, ti);
        const int wrow = ti.slot * p.Nout + ti.n0;
        int cb = 0, tap = 0;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kStages;
          if (it >= kStages) mbar_wait(empty_bar(s), ((it / kStages) - 1) & 1, p.err_flag, 4);
          mbar_arrive_expect_tx(full_bar(s), (uint32_t)BN * 128u);
          tma_load_2d(base + s * kStage + kStageA, &tmap_w, s_tkoff[tap] + cb * kBK, wrow, full_bar(s));
          if (++tap == taps) { tap = 0; ++cb; }
        }
      }
    }
  } else if (warp == kGW + 1) {
    // =========================================================================== MMA ISSUER
    const uint32_t idesc = make_idesc(BN, false, false);
    uint32_t it = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t buf = nbuf == 2 ? (tcount & 1) : 0u;
      const uint32_t use = nbuf == 2 ? (tcount >> 1) : tcount;      // how many times this buffer was used before
      if (use >= 1) mbar_wait(tempty_bar(buf), (use - 1) & 1, p.err_flag, 5);
      tc_fence_after();
      const uint32_t tacc = tmem_base + buf * (uint32_t)acc_cols;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % kStages;
        mbar_wait(full_bar(s), (it / kStages) & 1, p.err_flag, 2);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = base + s * kStage;
          const uint32_t sb = sa + kStageA;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t bd = make_desc(sb + k * 32, 16, 1024);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
              umma_bf16(tacc + mt * BN, make_desc(sa + mt * kFStageA + k * 32, 16, 1024), bd, idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(empty_bar(s));
          if (kb == nkb - 1) umma_commit(tfull_bar(buf));
        }
        __syncwarp();
      }
    }
    tc_fence_before();
  } else {
    // =========================================================================== EPILOGUE (last 4 warps)
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      TileInfo ti;
      decode_tile(tile, p, s_tiles, s_grp, kTileM, ti);
      const uint32_t buf = nbuf == 2 ? (tcount & 1) : 0u;
      const uint32_t use = nbuf == 2 ? (tcount >> 1) : tcount;
      // The tile's bias row is fetched into a per-warp shared-memory row BEFORE waiting for the accumulator: with one N tile
      // per bias segment (fc2: 360 tiles of 256 features per expert) every tile meets cold bias addresses, and 8 x 32 dependent
      // L2 round trips inside the epilogue made it the longest stage of the pipeline (ncu: 17k cycles per tile, tensor 14 %).
      const float* bias = p.bias ? p.bias + (long)ti.slot * p.bias_slot_stride + ti.n0 : nullptr;
      float* s_bias = reinterpret_cast<float*>(gen + 2048) + q * 256;
      uint8_t* s_stage = gen + 2048 + 4096 + q * (32 * 80);
      if (bias) {
        for (int i = lane; i < BN; i += 32) s_bias[i] = __ldg(bias + i);
        __syncwarp();
      }
      mbar_wait(tfull_bar(buf), use & 1, p.err_flag, 3);
      tc_fence_after();
      uint32_t r[32];
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const uint32_t t_lane = tmem_base + buf * (uint32_t)acc_cols + mt * BN + ((uint32_t)(q * 32) << 16);
        const int m = ti.m0 + mt * kBM + q * 32 + lane;
        const bool ok = m < ti.rows * p.P;
        const int smp = ok ? m / p.P : 0, pix = ok ? m - smp * p.P : 0;
        const int oa = pix / p.Wo, ob = pix - oa * p.Wo;
        const long opix = (long)(ti.row_start + smp) * p.P_full + (oa * p.o_my + p.o_oy) * p.Wo_full + ob * p.o_mx + p.o_ox;
        __nv_bfloat16* yrow = p.out + (opix * p.Nout + ti.n0);
        const unsigned long long yaddr = ok ? reinterpret_cast<unsigned long long>(yrow) : 0ull;
        for (int c = 0; c < BN; c += 32) {
          tmem_ld32(t_lane + c, r);
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(r[j]) + (bias ? s_bias[c + j] : 0.f);
          // A lane owns one output row, so storing from registers makes every STG touch 32 rows (16 bytes each).  The
          // 32 x 64-byte chunk goes through a per-warp shared-memory tile (80-byte pitch) instead: 4 lanes then write one
          // row's 64 contiguous bytes, 8 rows per instruction.
          uint4* w4 = reinterpret_cast<uint4*>(s_stage + lane * 80);
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) w4[qq] = pack8(f + 8 * qq);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = i * 8 + (lane >> 2), piece = lane & 3;
            const unsigned long long ra = __shfl_sync(0xffffffffu, yaddr, row);
            const uint4 v = *reinterpret_cast<const uint4*>(s_stage + row * 80 + piece * 16);
            if (ra) *reinterpret_cast<uint4*>(ra + (unsigned long long)(c * 2 + piece * 16)) = v;
          }
          __syncwarp();
        }
      }
      __syncwarp();                      // every lane is done with s_bias before the next tile overwrites it
      tc_fence_before();
      mbar_arrive(tempty_bar(buf));
    }
  }
  __syncthreads();
  if (warp == kGW + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// CTA-PAIR variant (cta_group::2).  Why: at 128 output pixels per CTA every CTA streams the WHOLE weight matrix of its expert
// through L2 -> shared memory for each tile — conv1 forward moved 7.3 GB of L2 reads in 0.89 ms (32 B/clk/SM against the
// ~42 B/clk/SM the fabric gives 148 SMs pulling at once), and the N <= 128 layers need 77 B/clk/SM at the tensor rate.  Two
// CTAs of a cluster (the two SMs of a TPC) now share ONE 256-pixel tile: each gathers its own 128 im2col rows, each loads only
// HALF of the weight tile (N/2 rows of B), and the leader issues tcgen05.mma.cta_group::2 with M = 256 — the tensor cores of
// both SMs read A from their own shared memory and the B halves from both.  Weight traffic per SM and per FLOP halves (L2
// and shared-memory operand reads alike), a stage shrinks to 16 + <= 16 KB so the pipeline is 6 deep instead of 4.
// Protocol (all barriers that gate the MMA live in the LEADER, rank 0):
//   full[s]   17 arrivals: 8 gather warps of each CTA (cp.async.wait_group -> fence.proxy.async -> __syncwarp -> lane 0
//             arrives, remotely from rank 1) + the leader's arrive.expect_tx for BOTH weight halves; rank 1's TMA signals its
//             bytes on the leader's barrier (cp.async.bulk.tensor ... .cta_group::2).
//   empty[s]  tcgen05.commit multicast to BOTH CTAs (each CTA's producers wait on their own copy).
//   tfull[b]  commit multicast to both; each CTA's epilogue drains ITS 128 accumulator rows from its own TMEM.
//   tempty[b] 8 arrivals on the leader: one per epilogue warp of each CTA.
template <int kStages>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGThreads, 1)
igemm_pair_kernel(const __grid_constant__ FwdParams p, const __grid_constant__ CUtensorMap tmap_wh) {
  constexpr int kLag = 3;
  constexpr int kStageB = 128 * 128;             // half of a BN <= 256 weight tile: <= 128 rows x 128 B
  constexpr int kStage = kFStageA + kStageB;     // 32 KB
  constexpr int kTileM = 2 * kBM;
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_base = base + kStages * kStage;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (16 + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (18 + b); };
  uint8_t* gen = smem_raw + (bar_base - raw);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + 8 * 20);
  const uint32_t tmem_slot = bar_base + 8u * 20;
  int* s_tiles = reinterpret_cast<int*>(gen + 256);            // [64]
  es_group* s_grp = reinterpret_cast<es_group*>(gen + 512);    // [64] x 16 B
  unsigned char* s_ymap = gen + 512 + 1024;
  unsigned char* s_xmap = s_ymap + 64;
  signed char* s_tdy = reinterpret_cast<signed char*>(s_xmap + 64);
  signed char* s_tdx = s_tdy + 32;
  int* s_tkoff = reinterpret_cast<int*>(s_tdx + 32);

  const int BN = p.BN, BH = p.BN / 2;
  const uint32_t nbuf = 2 * BN <= 512 ? 2u : 1u;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < (int)nbuf * BN) tmem_cols <<= 1;

  if (tid < p.n_groups) {
    const es_group gq = p.grp[tid];
    s_grp[tid] = gq;
    s_tiles[tid] = ceil_div(gq.rows * p.P, kTileM);            // PAIR tiles (256 pixels) of the group
  }
  if (tid < 64) { s_ymap[tid] = p.ymap[tid]; s_xmap[tid] = p.xmap[tid]; }
  if (tid < 32) { s_tdy[tid] = p.tdy[tid]; s_tdx[tid] = p.tdx[tid]; s_tkoff[tid] = p.tkoff[tid]; }
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 2 * kGW + 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kGW && lane == 0) tma_prefetch_desc(&tmap_wh);
  if (warp == kGW + 1) tmem_alloc_pair(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // both CTAs' barriers are initialised before anyone arrives remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  int total_tiles = 0;
  for (int i = 0; i < p.n_groups; ++i) total_tiles += s_tiles[i];
  total_tiles *= p.n_tiles_n;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int taps = p.n_taps;
  const int cblks = p.C / kBK;
  const int nkb = taps * cblks;

  if (warp < kGW) {
    // ========================================================================= A GATHER (this CTA's 128 rows of the pair tile)
    constexpr int RPT = kBM / (kGLoaders / 8);
    constexpr int RSTEP = kGLoaders / 8;
    const int chunk = tid & 7, rsub = tid >> 3;
    uint32_t it = 0, signalled = 0;
    for (int tile = pair; tile < total_tiles; tile += n_pairs) {
      TileInfo ti;
      decode_tile(tile, p, s_tiles, s_grp, kTileM, ti);
      const int m_first = ti.m0 + (int)rank * kBM;
      int r_base[RPT], r_oyx[RPT];
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const int m = m_first + rsub + RSTEP * j;
        const bool valid = m < ti.rows * p.P;
        const int sample = valid ? m / p.P : 0;
        const int pix = valid ? m - sample * p.P : 0;
        const int oy = pix / p.Wo;
        r_oyx[j] = (oy << 8) | (pix - oy * p.Wo);
        r_base[j] = valid ? (ti.row_start + sample) * p.Hs * p.Ws : -1;
      }
      int cb = 0, tap = 0;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % kStages;
        if (it >= (uint32_t)kStages) mbar_wait(empty_bar(s), ((it / kStages) - 1) & 1, p.err_flag, 1);
        const uint32_t sa = base + s * kStage;
        const int c0 = cb * kBK + chunk * 8;
        const int ty = s_tdy[tap], tx = s_tdx[tap];
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
          const int r = rsub + RSTEP * j;
          const int uy = (r_oyx[j] >> 8) * p.my + ty, ux = (r_oyx[j] & 255) * p.mx + tx;
          const bool inb = r_base[j] >= 0 && uy >= 0 && uy < p.Hu && ux >= 0 && ux < p.Wu;
          const int sy = inb ? s_ymap[uy] : 0, sx = inb ? s_xmap[ux] : 0;
          const __nv_bfloat16* src = p.a_src + (inb ? ((long)(r_base[j] + sy * p.Ws + sx) * p.C + c0) : 0L);
          cp_async16_ca(sa + (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4), src, inb);
        }
        cp_async_commit();
        if (++tap == taps) { tap = 0; ++cb; }
        if (it - signalled >= (uint32_t)kLag) {
          cp_async_wait<kLag>();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_shared(full_bar(signalled % kStages), 0));
          ++signalled;
        }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    __syncwarp();
    while (signalled < it) {
      if (lane == 0) mbar_arrive_cluster(mapa_shared(full_bar(signalled % kStages), 0));
      ++signalled;
    }
  } else if (warp == kGW) {
    // =========================================================================== TMA PRODUCER (this CTA's half of B)
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs) {
        TileInfo ti;
        decode_tile(tile, p, s_tiles, s_grp, kTileM, ti);
        const int wrow = ti.slot * p.Nout + ti.n0 + (int)rank * BH;
        int cb = 0, tap = 0;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kStages;
          if (it >= (uint32_t)kStages) mbar_wait(empty_bar(s), ((it / kStages) - 1) & 1, p.err_flag, 4);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(s), (uint32_t)BN * 128u);      // both halves
          tma_load_2d_pair(base + s * kStage + kFStageA, &tmap_wh, s_tkoff[tap] + cb * kBK, wrow, mapa_shared(full_bar(s), 0));
          if (++tap == taps) { tap = 0; ++cb; }
        }
      }
    }
  } else if (warp == kGW + 1) {
    // =========================================================================== MMA ISSUER (leader CTA only)
    if (rank == 0) {
      const uint32_t idesc = make_idesc_m(BN, 2 * kBM, false, false);
      uint32_t it = 0, tcount = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs, ++tcount) {
        const uint32_t buf = nbuf == 2 ? (tcount & 1) : 0u;
        const uint32_t use = nbuf == 2 ? (tcount >> 1) : tcount;
        if (use >= 1) mbar_wait(tempty_bar(buf), (use - 1) & 1, p.err_flag, 5);
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * (uint32_t)BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kStages;
          // cta-scope wait: `.acquire.cluster` would add MEMBAR + CCTL.IVALL (an L1 invalidate) per k-block on the leader, and
          // this thread reads nothing through the generic proxy — it only issues MMAs
          mbar_wait(full_bar(s), (it / kStages) & 1, p.err_flag, 2);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = base + s * kStage;
            const uint32_t sb = sa + kFStageA;
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_bf16_pair(tacc, make_desc(sa + k * 32, 16, 1024), make_desc(sb + k * 32, 16, 1024), idesc, (kb | k) ? 1u : 0u);
            umma_commit_pair(empty_bar(s), 3);
            if (kb == nkb - 1) umma_commit_pair(tfull_bar(buf), 3);
          }
          __syncwarp();
        }
      }
      tc_fence_before();
    }
  } else {
    // =========================================================================== EPILOGUE (this CTA's 128 rows)
    const int q = warp & 3;
    uint32_t tcount = 0;
    for (int tile = pair; tile < total_tiles; tile += n_pairs, ++tcount) {
      TileInfo ti;
      decode_tile(tile, p, s_tiles, s_grp, kTileM, ti);
      const uint32_t buf = nbuf == 2 ? (tcount & 1) : 0u;
      const uint32_t use = nbuf == 2 ? (tcount >> 1) : tcount;
      mbar_wait(tfull_bar(buf), use & 1, p.err_flag, 3);
      tc_fence_after();
      const float* bias = p.bias ? p.bias + (long)ti.slot * p.bias_slot_stride + ti.n0 : nullptr;
      uint32_t r[32];
      const uint32_t t_lane = tmem_base + buf * (uint32_t)BN + ((uint32_t)(q * 32) << 16);
      const int m = ti.m0 + (int)rank * kBM + q * 32 + lane;
      const bool ok = m < ti.rows * p.P;
      const int smp = ok ? m / p.P : 0, pix = ok ? m - smp * p.P : 0;
      const int oa = pix / p.Wo, ob = pix - oa * p.Wo;
      const long opix = (long)(ti.row_start + smp) * p.P_full + (oa * p.o_my + p.o_oy) * p.Wo_full + ob * p.o_mx + p.o_ox;
      __nv_bfloat16* yrow = p.out + (opix * p.Nout + ti.n0);
      for (int c = 0; c < BN; c += 32) {
        tmem_ld32(t_lane + c, r);
        if (ok) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(r[j]) + (bias ? __ldg(bias + c + j) : 0.f);
          uint4* dst = reinterpret_cast<uint4*>(yrow + c);
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) dst[qq] = pack8(f + 8 * qq);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(tempty_bar(buf), 0));
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // the leader's MMAs read the peer's shared memory: nobody leaves before both are done
  if (warp == kGW + 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// TMA-FED CTA-PAIR variant: BOTH operands arrive by TMA.  For a tap-table conv that reads the source directly — source pixel
// (oy*my + dy_t, ox*mx + dx_t), no nearest-upsample map in between: the phase-folded x2 convs, their data gradients, every
// plain conv's data gradient — the im2col rows of a tile are exactly what TMA's im2col mode enumerates: 128 consecutive base
// pixels of a bounding box (lower corner = smallest tap offset, traversal stride = my/mx), shifted by the tap offset given in
// the instruction, zero-filled outside the image, written as 128B-swizzled 64-channel rows — the K-major layout the MMA
// reads.  One instruction per (tap, channel block) replaces 256 threads x 4 cp.async: no address arithmetic, no L1TEX
// traffic (the gather read every line through L1 and wrote it to shared memory: two passes over the 128 B/clk SRAM that
// the tensor core's operand reads also need — ncu: L1TEX 58-63 % at 27-46 % tensor), ragged expert tails simply run into the
// next sample (the epilogue masks those rows).  6 warps: TMA producer, MMA issuer / TMEM owner, 4 epilogue warps.
// Barrier protocol: as igemm_pair_kernel, except that full[s] takes ONE arrival (the leader's expect_tx for the four boxes of
// the stage — A and the B half of both CTAs; rank 1's boxes signal the leader's barrier).
// Norm fusion: the GroupNorm that follows a generator conv needs per-(sample, group) mean and variance of the conv's output —
// a reduction over the whole sample that forced the norm kernel to read its input twice or to park it in shared memory.
// The epilogue already holds every output value in registers, so it accumulates per-(sample, channel PAIR) sum and sum of
// squares of the bf16-ROUNDED values it stores (what the norm kernel will read back) and the norm becomes one streaming pass.
// f[32] = this lane's pixel, 32 consecutive channels.  The 16 pairs x 2 moments = 32 values per lane are transposed-reduced
// over the warp's 32 pixels with 31 shuffles (recursive halving: lane L ends up with the total of value L), then one
// coalesced 128-byte atomic add per warp.  Lanes of a warp almost always belong to one sample; a boundary costs a second round.
__device__ __forceinline__ void epilogue_pair_sums(const float* f, bool ok, int row, float* __restrict__ sums, int half_n,
                                                   int col0, int lane) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float a = __bfloat162float(__float2bfloat16_rn(f[2 * j])), b = __bfloat162float(__float2bfloat16_rn(f[2 * j + 1]));
    v[2 * j] = a + b;
    v[2 * j + 1] = a * a + b * b;
  }
  unsigned todo = __ballot_sync(0xffffffffu, ok);
  while (todo) {
    const int lrow = __shfl_sync(0xffffffffu, row, __ffs(todo) - 1);
    const bool mine = ok && row == lrow;
    float w[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) w[i] = mine ? v[i] : 0.f;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const bool hi = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < off; ++i) {
        const float send = hi ? w[i] : w[i + off], keep = hi ? w[i + off] : w[i];
        w[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    atomicAdd(sums + ((long)lrow * half_n + (col0 >> 1)) * 2 + lane, w[0]);
    todo &= ~__ballot_sync(0xffffffffu, mine);
  }
}

struct TmaAParams {
  int low_w, low_h;         // bounding-box lower corner = base-pixel coordinate of output (0, 0)
};

template <int kStages>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
igemm_tma_pair_kernel(const __grid_constant__ FwdParams p, const __grid_constant__ TmaAParams ta,
                      const __grid_constant__ CUtensorMap tmap_wh, const __grid_constant__ CUtensorMap tmap_x) {
  constexpr int kStageB = 128 * 128;
  constexpr int kStage = kFStageA + kStageB;     // 32 KB
  constexpr int kTileM = 2 * kBM;
  extern __shared__ uint8_t smem_raw[];
  // (shuffle: the warp index becomes warp-uniform for the compiler, so the MMA issuer's descriptors live in uniform registers)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_base = base + kStages * kStage;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (16 + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (18 + b); };
  uint8_t* gen = smem_raw + (bar_base - raw);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + 8 * 20);
  const uint32_t tmem_slot = bar_base + 8u * 20;
  int* s_tiles = reinterpret_cast<int*>(gen + 256);            // [64]
  es_group* s_grp = reinterpret_cast<es_group*>(gen + 512);    // [64] x 16 B

  const int BN = p.BN, BH = p.BN / 2;
  const uint32_t nbuf = 2 * BN <= 512 ? 2u : 1u;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < (int)nbuf * BN) tmem_cols <<= 1;

  if (tid < p.n_groups) {
    const es_group gq = p.grp[tid];
    s_grp[tid] = gq;
    s_tiles[tid] = ceil_div(gq.rows * p.P, kTileM);
  }
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap_wh); tma_prefetch_desc(&tmap_x); }
  if (warp == 1) tmem_alloc_pair(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  int total_tiles = 0;
  for (int i = 0; i < p.n_groups; ++i) total_tiles += s_tiles[i];
  total_tiles *= p.n_tiles_n;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int taps = p.n_taps;
  const int cblks = p.C / kBK;
  const int nkb = taps * cblks;

  if (warp == 0) {
    // =========================================================================== TMA PRODUCER: A (im2col) + this CTA's half of B
    if (lane == 0) {
      const uint32_t lead_full0 = mapa_shared(full_bar(0), 0);
      uint32_t it = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs) {
        TileInfo ti;
        decode_tile(tile, p, s_tiles, s_grp, kTileM, ti);
        const int wrow = ti.slot * p.Nout + ti.n0 + (int)rank * BH;
        // first im2col row of this CTA's half: output pixel (sample, oy, ox) -> base coordinate (lower corner + index * stride);
        // rows past the group's end run into the next sample (masked by the epilogue) or off the tensor (zero-filled)
        const int m = ti.m0 + (int)rank * kBM;
        const int sample = m / p.P, pix = m - sample * p.P;
        const int oy = pix / p.Wo, ox = pix - oy * p.Wo;
        const int cn = ti.row_start + sample, ch = ta.low_h + oy * p.my, cw = ta.low_w + ox * p.mx;
        int cb = 0, tap = 0;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kStages;
          if (it >= (uint32_t)kStages) mbar_wait(empty_bar(s), ((it / kStages) - 1) & 1, p.err_flag, 4);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(s), 2u * (uint32_t)kFStageA + (uint32_t)BN * 128u);
          const uint32_t lead_full = lead_full0 + 8u * s;
          const uint32_t sa = base + s * kStage;
          tma_im2col_4d_pair(sa, &tmap_x, cb * kBK, cw, ch, cn, (uint32_t)(p.tdx[tap] - ta.low_w), (uint32_t)(p.tdy[tap] - ta.low_h), lead_full);
          tma_load_2d_pair(sa + kFStageA, &tmap_wh, p.tkoff[tap] + cb * kBK, wrow, lead_full);
          if (++tap == taps) { tap = 0; ++cb; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================================================================== MMA ISSUER (leader CTA only)
    if (rank == 0) {
      const uint32_t idesc = make_idesc_m(BN, 2 * kBM, false, false);
      uint32_t it = 0, tcount = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs, ++tcount) {
        const uint32_t buf = nbuf == 2 ? (tcount & 1) : 0u;
        const uint32_t use = nbuf == 2 ? (tcount >> 1) : tcount;
        if (use >= 1) mbar_wait(tempty_bar(buf), (use - 1) & 1, p.err_flag, 5);
        tc_fence_after();
        const uint32_t tacc = __shfl_sync(0xffffffffu, tmem_base, 0) + buf * (uint32_t)BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kStages;
          mbar_wait(full_bar(s), (it / kStages) & 1, p.err_flag, 2);
          tc_fence_after();
          const uint32_t sa = base + s * kStage;          // descriptors from warp-uniform values, outside the lane-0 branch
          const uint64_t a0 = make_desc(sa, 16, 1024), b0 = make_desc(sa + kFStageA, 16, 1024);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_bf16_pair_elect(tacc, a0 + 2 * k, b0 + 2 * k, idesc, (kb | k) ? 1u : 0u);
          umma_commit_pair_elect(empty_bar(s), 3);
          if (kb == nkb - 1) umma_commit_pair_elect(tfull_bar(buf), 3);
          __syncwarp();
        }
      }
      tc_fence_before();
    }
  } else {
    // =========================================================================== EPILOGUE (warps 2-5: TMEM lane quarter = warp % 4)
    const int q = warp & 3;
    uint32_t tcount = 0;
    for (int tile = pair; tile < total_tiles; tile += n_pairs, ++tcount) {
      TileInfo ti;
      decode_tile(tile, p, s_tiles, s_grp, kTileM, ti);
      const uint32_t buf = nbuf == 2 ? (tcount & 1) : 0u;
      const uint32_t use = nbuf == 2 ? (tcount >> 1) : tcount;
      const float* bias = p.bias ? p.bias + (long)ti.slot * p.bias_slot_stride + ti.n0 : nullptr;
      // dense product (fc2: one N tile per bias segment, 184 KB between output rows): the bias row is fetched into a per-warp
      // shared-memory row BEFORE the accumulator wait and the stores are transposed through shared memory (as igemm_fwd_kernel)
      float* s_bias = reinterpret_cast<float*>(gen + 2048) + q * 256;
      uint8_t* s_stage = gen + 2048 + 4096 + q * (32 * 80);
      const bool staged = p.epi_staged != 0;
      if (staged && bias) {
        for (int i = lane; i < BN; i += 32) s_bias[i] = __ldg(bias + i);
        __syncwarp();
      }
      mbar_wait(tfull_bar(buf), use & 1, p.err_flag, 3);
      tc_fence_after();
      uint32_t r[32];
      const uint32_t t_lane = tmem_base + buf * (uint32_t)BN + ((uint32_t)(q * 32) << 16);
      const int m = ti.m0 + (int)rank * kBM + q * 32 + lane;
      const bool ok = m < ti.rows * p.P;
      const int smp = ok ? m / p.P : 0, pix = ok ? m - smp * p.P : 0;
      const int oa = pix / p.Wo, ob = pix - oa * p.Wo;
      const long opix = (long)(ti.row_start + smp) * p.P_full + (oa * p.o_my + p.o_oy) * p.Wo_full + ob * p.o_mx + p.o_ox;
      __nv_bfloat16* yrow = p.out + (opix * p.Nout + ti.n0);
      const unsigned long long yaddr = ok ? reinterpret_cast<unsigned long long>(yrow) : 0ull;
      for (int c = 0; c < BN; c += 32) {
        tmem_ld32(t_lane + c, r);
        float f[32];
        if (staged) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(r[j]) + (bias ? s_bias[c + j] : 0.f);
          uint4* w4 = reinterpret_cast<uint4*>(s_stage + lane * 80);
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) w4[qq] = pack8(f + 8 * qq);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = i * 8 + (lane >> 2), piece = lane & 3;
            const unsigned long long ra = __shfl_sync(0xffffffffu, yaddr, row);
            const uint4 v = *reinterpret_cast<const uint4*>(s_stage + row * 80 + piece * 16);
            if (ra) *reinterpret_cast<uint4*>(ra + (unsigned long long)(c * 2 + piece * 16)) = v;
          }
          __syncwarp();
          continue;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(r[j]) + (bias ? __ldg(bias + c + j) : 0.f);
        if (ok) {
          uint4* dst = reinterpret_cast<uint4*>(yrow + c);
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) dst[qq] = pack8(f + 8 * qq);
        }
        if (p.pair_sums) epilogue_pair_sums(f, ok, ti.row_start + smp, p.pair_sums, p.Nout >> 1, ti.n0 + c, lane);
      }
      __syncwarp();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(tempty_bar(buf), 0));
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// TMA-fed pair kernel with TAP-ROW STRIPS.  igemm_tma_pair_kernel is bound by the L2->SM fill (one 16 KB im2col box per tap per
// CTA against 128 x N x 64 MACs: ~41 B/clk/SM arrive, the MMA of an N = 64/128 k-block wants 4x/2x that).  When the taps of
// the table form rows (same dy, nx consecutive dx, consecutive weight columns) the nx windows of a row are ONE strip of
// 128 + nx - 1 base pixels if the M axis enumerates output pixels with the PADDED row pitch Wp = Wo + nx - 1 — which in TMA
// im2col terms is just a bounding box nx - 1 columns wider: one im2col instruction of 128 + nx - 1 pixels per (strip, channel
// block), and the nx taps are MMAs whose A descriptors start j rows (j * 128 B) further down the same 128B-swizzled strip
// (the 128B swizzle is a function of the absolute shared-memory address, so a row-shifted start stays consistent with what
// TMA wrote.  Measured on sm_100a: the descriptor's base-offset field must stay 0 for such starts — setting it to the row
// phase j gives wrong products).  Fill per MAC drops by nx on the A side; the masked columns cost Wp / Wo - 1 of the MMA
// work.  mt = 2 M sub-tiles per pair share every weight box (N <= 128: halves the weight fill as well).
constexpr int kTSThreads = 320;                  // producer warp, MMA warp, 2 x 4 epilogue warps
struct TStripParams {
  int n_strips, nx, Wp, Pp;                      // Pp = Ho * Wp
  int mt;                                        // 256-row M sub-tiles per pair tile (1 or 2): sub-tiles share every weight box
  int stages, stage_bytes, a_bytes;              // a_bytes = strip bytes rounded up to 1 KB; stage = mt * a_bytes + nx * BN/2 * 128
  int low_w, low_h;
  signed char sdy[16], sdx[16];
  int skoff[16];                                 // weight column of the strip's first tap; tap j at skoff + j*C
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTSThreads, 1)
igemm_tma_strip_kernel(const __grid_constant__ FwdParams p, const __grid_constant__ TStripParams ts,
                       const __grid_constant__ CUtensorMap tmap_wh, const __grid_constant__ CUtensorMap tmap_x) {
  extern __shared__ uint8_t smem_raw[];
  // the shuffle makes the warp index warp-UNIFORM for the compiler: role branches become uniform branches and the MMA issuer's
  // descriptors stay in uniform registers
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int mt = ts.mt, kTileM = 2 * kBM * mt;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int kStages = ts.stages, kStage = ts.stage_bytes;
  const uint32_t bar_base = base + kStages * kStage;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (16 + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (18 + b); };
  uint8_t* gen = smem_raw + (bar_base - raw);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + 8 * 20);
  const uint32_t tmem_slot = bar_base + 8u * 20;
  int* s_tiles = reinterpret_cast<int*>(gen + 256);            // [64]
  es_group* s_grp = reinterpret_cast<es_group*>(gen + 512);    // [64] x 16 B

  const int BN = p.BN, BH = p.BN / 2, nx = ts.nx;
  const uint32_t acc_cols = (uint32_t)(mt * BN);                // accumulator columns per buffer: one N tile per M sub-tile
  const uint32_t nbuf = 2 * acc_cols <= 512 ? 2u : 1u;
  uint32_t tmem_cols = 32;
  while (tmem_cols < nbuf * acc_cols) tmem_cols <<= 1;

  if (tid < p.n_groups) {
    const es_group gq = p.grp[tid];
    s_grp[tid] = gq;
    s_tiles[tid] = ceil_div(gq.rows * ts.Pp, kTileM);
  }
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 8 * mt);       // one arrival per active epilogue warp of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap_wh); tma_prefetch_desc(&tmap_x); }
  if (warp == 1) tmem_alloc_pair(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  int total_tiles = 0;
  for (int i = 0; i < p.n_groups; ++i) total_tiles += s_tiles[i];
  total_tiles *= p.n_tiles_n;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int cblks = p.C / kBK;
  const int nkb = ts.n_strips * cblks;
  const uint32_t strip_bytes = (uint32_t)(kBM + nx - 1) * 128u;

  if (warp == 0) {
    // =========================================================================== TMA PRODUCER: strip (im2col) + nx half weight boxes
    // The whole warp runs the loop on warp-uniform values (tile fields broadcast by shuffle) and every instruction is issued by
    // the elected lane inside the asm statement, so coordinates and addresses stay in uniform registers (see the MMA issuer).
    {
      const uint32_t lead_full0 = __shfl_sync(0xffffffffu, mapa_shared(full_bar(0), 0), 0);
      uint32_t it = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs) {
        TileInfo ti;
        decode_tile(tile, p, s_tiles, s_grp, kTileM, ti);
        const int t_slot = uniform_i32(ti.slot), t_n0 = uniform_i32(ti.n0), t_m0 = uniform_i32(ti.m0), t_row0 = uniform_i32(ti.row_start);
        const int wrow = t_slot * p.Nout + t_n0 + (int)rank * BH;
        // base coordinates of the two M sub-tiles (scalars, not arrays: a runtime-indexed array would live in local memory)
        const int m_a = t_m0 + (int)rank * kBM, m_b = m_a + 2 * kBM;
        const int smp_a = m_a / ts.Pp, pix_a = m_a - smp_a * ts.Pp, oy_a = pix_a / ts.Wp, ox_a = pix_a - oy_a * ts.Wp;
        const int smp_b = m_b / ts.Pp, pix_b = m_b - smp_b * ts.Pp, oy_b = pix_b / ts.Wp, ox_b = pix_b - oy_b * ts.Wp;
        const int cn_a = t_row0 + smp_a, ch_a = ts.low_h + oy_a * p.my, cw_a = ts.low_w + ox_a;
        const int cn_b = t_row0 + smp_b, ch_b = ts.low_h + oy_b * p.my, cw_b = ts.low_w + ox_b;
        int cb = 0, st = 0;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kStages;
          if (it >= (uint32_t)kStages) {        // ONE lane polls: the producer runs ahead and spins here most of the time, and 32
                                                // lanes spinning on try_wait slowed the whole CTA (conv3 fwd 0.85 vs 0.47 ms)
            if (lane == 0) mbar_wait(empty_bar(s), ((it / kStages) - 1) & 1, p.err_flag, 4);
            __syncwarp();
          }
          if (rank == 0) mbar_arrive_expect_tx_elect(full_bar(s), 2u * (uint32_t)mt * strip_bytes + (uint32_t)(nx * BN) * 128u);
          const uint32_t lead_full = lead_full0 + 8u * s;
          const uint32_t sa = base + s * kStage;
          const uint32_t offw = (uint32_t)(ts.sdx[st] - ts.low_w), offh = (uint32_t)(ts.sdy[st] - ts.low_h);
          tma_im2col_4d_pair_elect(sa, &tmap_x, cb * kBK, cw_a, ch_a, cn_a, offw, offh, lead_full);
          if (mt == 2) tma_im2col_4d_pair_elect(sa + ts.a_bytes, &tmap_x, cb * kBK, cw_b, ch_b, cn_b, offw, offh, lead_full);
#pragma unroll 1
          for (int j = 0; j < nx; ++j)
            tma_load_2d_pair_elect(sa + mt * ts.a_bytes + j * BH * 128, &tmap_wh, ts.skoff[st] + j * p.C + cb * kBK, wrow, lead_full);
          if (++st == ts.n_strips) { st = 0; ++cb; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================================================================== MMA ISSUER (leader CTA only)
    if (rank == 0) {
      const uint32_t idesc = make_idesc_m(BN, 2 * kBM, false, false);
      uint32_t it = 0, tcount = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs, ++tcount) {
        const uint32_t buf = nbuf == 2 ? (tcount & 1) : 0u;
        const uint32_t use = nbuf == 2 ? (tcount >> 1) : tcount;
        if (use >= 1) mbar_wait(tempty_bar(buf), (use - 1) & 1, p.err_flag, 5);
        tc_fence_after();
        const uint32_t tacc = __shfl_sync(0xffffffffu, tmem_base, 0) + buf * acc_cols;     // warp-uniform for the compiler
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kStages;
          mbar_wait(full_bar(s), (it / kStages) & 1, p.err_flag, 2);
          tc_fence_after();
          // Descriptors are built by the WHOLE warp from warp-uniform values (so they live in uniform registers) and only the
          // instruction itself is predicated on lane 0: built inside a lane-0 branch, every tcgen05.mma dragged a 16-instruction
          // R2UR / ELECT sequence along — ~120 cycles per issue, more than a 256 x N x 16 MMA takes at N <= 128.
          const uint32_t sa = base + s * kStage;
          const uint64_t a0 = make_desc(sa, 16, 1024), b0 = make_desc(sa + mt * ts.a_bytes, 16, 1024);
          for (int j = 0; j < nx; ++j) {
            const uint64_t bj = b0 + (uint64_t)((j * BH * 128) >> 4);
            for (int u = 0; u < mt; ++u) {
              const uint64_t aj = a0 + (uint64_t)((u * ts.a_bytes + j * 128) >> 4);
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k)
                umma_bf16_pair_elect(tacc + u * BN, aj + 2 * k, bj + 2 * k, idesc, (kb | j | k) ? 1u : 0u);
            }
          }
          umma_commit_pair_elect(empty_bar(s), 3);
          if (kb == nkb - 1) umma_commit_pair_elect(tfull_bar(buf), 3);
          __syncwarp();
        }
      }
      tc_fence_before();
    }
  } else {
    // =========================================================================== EPILOGUE (warps 2-9: TMEM lane quarter = warp % 4;
    // warps 2-5 drain M sub-tile 0, warps 6-9 sub-tile 1 — with the GroupNorm sums in it the epilogue of a 512-row tile is as
    // long as its MMAs, two warp groups keep it off the critical path)
    const int q = warp & 3, u = (warp - 2) >> 2;
    uint32_t tcount = 0;
    if (u < mt)
    for (int tile = pair; tile < total_tiles; tile += n_pairs, ++tcount) {
      TileInfo ti;
      decode_tile(tile, p, s_tiles, s_grp, kTileM, ti);
      const uint32_t buf = nbuf == 2 ? (tcount & 1) : 0u;
      const uint32_t use = nbuf == 2 ? (tcount >> 1) : tcount;
      mbar_wait(tfull_bar(buf), use & 1, p.err_flag, 3);
      tc_fence_after();
      const float* bias = p.bias ? p.bias + (long)ti.slot * p.bias_slot_stride + ti.n0 : nullptr;
      uint32_t r[32];
      {
      const uint32_t t_lane = tmem_base + buf * acc_cols + (uint32_t)(u * BN) + ((uint32_t)(q * 32) << 16);
      const int m = ti.m0 + u * 2 * kBM + (int)rank * kBM + q * 32 + lane;
      bool ok = m < ti.rows * ts.Pp;
      const int smp = ok ? m / ts.Pp : 0, pix = ok ? m - smp * ts.Pp : 0;
      const int oa = pix / ts.Wp, ob = pix - oa * ts.Wp;
      ok = ok && ob < p.Wo;                                      // the nx - 1 pad columns of the padded pitch are not outputs
      const long opix = (long)(ti.row_start + smp) * p.P_full + (oa * p.o_my + p.o_oy) * p.Wo_full + ob * p.o_mx + p.o_ox;
      __nv_bfloat16* yrow = p.out + (opix * p.Nout + ti.n0);
      for (int c = 0; c < BN; c += 32) {
        tmem_ld32(t_lane + c, r);
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(r[j]) + (bias ? __ldg(bias + c + j) : 0.f);
        if (ok) {
          uint4* dst = reinterpret_cast<uint4*>(yrow + c);
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) dst[qq] = pack8(f + 8 * qq);
        }
        if (p.pair_sums) epilogue_pair_sums(f, ok, ti.row_start + smp, p.pair_sums, p.Nout >> 1, ti.n0 + c, lane);
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(tempty_bar(buf), 0));
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// STRIP variant for convolutions WITHOUT an upsample in front whose taps form rows (same dy, consecutive dx): conv3 forward
// (3x3, 128 -> 64) and its data gradient (3x3, 64 -> 128), the layers where the gather is the bound — an im2col row is used
// for only N = 64/128 MACs per element, so the 64 B/clk/SM that cp.async moves through L1TEX caps the kernel at ~50 % of the
// tensor peak however the rest is tuned (r01: 288 / 403 TFLOP/s).
// Here the M axis enumerates output pixels with a PADDED row pitch Wp = Wo + nx - 1, so the kx-shifted windows of one tap
// row are ONE contiguous strip of 128 + nx - 1 source pixels: the strip is gathered once per (tap row, channel block) and
// the nx taps are nx MMAs whose A descriptors start kx rows further down the same strip.  Row shifts of an arbitrary
// number of 16-byte units are exact in the NO-SWIZZLE K-major canonical layout ((8,m),(T,2)):((1T,SBO),(1,LBO)) when
// SBO = 8 x 16 B (rows fully linear, 16 B apart) and LBO = strip pitch between the eight 16-byte k-chunks of a row; B
// (weights) stays a 128B-swizzled TMA box per tap.  Gather traffic per MAC drops by nx (3x for a 3x3), the price is
// Wp/Wo - 1 (7 % at 29 columns) of masked output columns.
constexpr int kSRows = 136;                      // strip rows kept per stage (>= 128 + 4 - 1)
constexpr int kSLbo = (kSRows + 1) * 16;         // 2192 B between k-chunks: 16-byte skew keeps the 8 lanes of a row on different banks
constexpr int kSStageA = 18432;                  // 8 * kSLbo = 17536, rounded up to 1 KB
constexpr int kSMaxB = 384 * 128;                // nx * BN <= 384 rows of 128 B
constexpr int kSStages = 3;
constexpr size_t kSSmem = (size_t)kSStages * (kSStageA + kSMaxB) + 1024 + 2048;
// RESIDENT weights: with the strip gather the weight stream became the larger operand (conv3: 147 KB of weights against
// 100 KB of strips per 128-pixel tile, 64 B/clk/SM of L2->SM reads at the tensor rate of N = 64), but 147 KB is all there is
// per expert — so when KK * N * 2 B fits beside the strip stages the whole matrix is loaded once per expert CHANGE (tiles are
// visited in ascending order, so at most n_groups times per CTA) and only the strips move per tile.
constexpr int kSResidentB = kSStages * kSMaxB;   // 144 KB: the room the per-stage weight tiles would take

struct StripParams {
  int resident;                                  // 1: the expert's whole weight matrix stays in shared memory across tiles
  int stages;                                    // pipeline depth: 3, or 4 when four (strip + nx weight boxes) stages fit
  int n_strips, nx, dx0, Wp, Pp;                 // Pp = Ho * Wp
  signed char sdy[8];
  int skoff[8];                                  // weight column of the strip's first tap; tap i at skoff + i*C
};

__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

__global__ void __launch_bounds__(kGThreads, 1)
igemm_strip_kernel(const __grid_constant__ FwdParams p, const __grid_constant__ StripParams sp,
                   const __grid_constant__ CUtensorMap tmap_w) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = p.BN;
  const int kStageB = sp.nx * BN * 128;
  const bool resident = sp.resident != 0;
  const int kStage = resident ? kSStageA : kSStageA + kStageB;       // resident: stages hold strips only
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int ns = sp.stages;                                           // 3 or 4 (the barrier area holds 4 + 4 stage barriers)
  const uint32_t wbase = base + kSStages * kSStageA;                  // resident weights: [k-step][tap][BN rows x 128 B]
  const uint32_t bar_base = base + kSStages * (kSStageA + kSMaxB);
  const uint32_t wfull_bar = bar_base + 8u * 14, wempty_bar = bar_base + 8u * 15;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (4 + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (8 + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (10 + b); };
  uint8_t* gen = smem_raw + (bar_base - raw);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + 8 * 12);
  const uint32_t tmem_slot = bar_base + 8u * 12;
  int* s_tiles = reinterpret_cast<int*>(gen + 128);
  es_group* s_grp = reinterpret_cast<es_group*>(gen + 384);

  const uint32_t nbuf = 2;                                       // BN <= 128: two accumulators always fit
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * BN) tmem_cols <<= 1;

  if (tid < p.n_groups) {
    const es_group gq = p.grp[tid];
    s_grp[tid] = gq;
    s_tiles[tid] = ceil_div(gq.rows * sp.Pp, kBM);
  }
  if (tid == 0) {
    for (int s = 0; s < ns; ++s) {
      mbar_init(full_bar(s), resident ? kGLoaders : kGLoaders + 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 128);
    }
    mbar_init(wfull_bar, 1);
    mbar_init(wempty_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kGW && lane == 0) tma_prefetch_desc(&tmap_w);
  if (warp == kGW + 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  int total_tiles = 0;
  for (int i = 0; i < p.n_groups; ++i) total_tiles += s_tiles[i];
  total_tiles *= p.n_tiles_n;
  const int cblks = p.C / kBK;
  const int nkb = sp.n_strips * cblks;                           // pipeline steps per tile
  const int srows = kBM + sp.nx - 1;

  if (warp < kGW) {
    // ============================================================================= STRIP GATHER (256 threads, cp.async)
    constexpr int RSTEP = kGLoaders / 8;                         // 32 rows per pass
    constexpr int RPT = (kSRows + RSTEP - 1) / RSTEP;            // 5
    const int chunk = tid & 7, rsub = tid >> 3;
    uint32_t it = 0, signalled = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      TileInfo ti;
      decode_tile(tile, p, s_tiles, s_grp, kBM, ti);
      int r_base[RPT], r_oyx[RPT];                               // source pixel base of the sample (-1: no such row), oy<<8 | xv
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const int r = rsub + RSTEP * j;
        const int v = ti.m0 + r;
        const bool valid = r < srows && v < ti.rows * sp.Pp;
        const int sample = valid ? v / sp.Pp : 0;
        const int rem = valid ? v - sample * sp.Pp : 0;
        const int oy = rem / sp.Wp;
        r_oyx[j] = (oy << 8) | (rem - oy * sp.Wp);
        r_base[j] = valid ? (ti.row_start + sample) * p.Hs * p.Ws : -1;
      }
      int cb = 0, st = 0;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % ns;
        if (it >= ns) mbar_wait(empty_bar(s), ((it / ns) - 1) & 1, p.err_flag, 1);
        const uint32_t sa = base + s * kStage;
        const int c0 = cb * kBK + chunk * 8;
        const int dy = sp.sdy[st];
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
          const int r = rsub + RSTEP * j;
          if (r < srows) {
            const int sy = (r_oyx[j] >> 8) + dy, sx = (r_oyx[j] & 255) + sp.dx0;
            const bool inb = r_base[j] >= 0 && sy >= 0 && sy < p.Hs && sx >= 0 && sx < p.Ws;
            const __nv_bfloat16* src = p.a_src + (inb ? ((long)(r_base[j] + sy * p.Ws + sx) * p.C + c0) : 0L);
            cp_async16_ca(sa + (uint32_t)chunk * kSLbo + (uint32_t)r * 16u, src, inb);
          }
        }
        cp_async_commit();
        if (++st == sp.n_strips) { st = 0; ++cb; }
        if (it - signalled >= (uint32_t)kFLag) {
          cp_async_wait<kFLag>();
          fence_proxy_async();
          mbar_arrive(full_bar(signalled % ns));
          ++signalled;
        }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    while (signalled < it) {
      mbar_arrive(full_bar(signalled % ns));
      ++signalled;
    }
  } else if (warp == kGW) {
    // ============================================================================= TMA PRODUCER (nx weight boxes per step)
    if (lane == 0) {
      uint32_t it = 0, nload = 0;
      int cur_slot = -1;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        TileInfo ti;
        decode_tile(tile, p, s_tiles, s_grp, kBM, ti);
        const int wrow = ti.slot * p.Nout + ti.n0;
        int cb = 0, st = 0;
        if (resident) {
          if (ti.slot == cur_slot) continue;
          if (nload > 0) mbar_wait(wempty_bar, (nload - 1) & 1, p.err_flag, 6);   // every MMA on the old weights has completed
          cur_slot = ti.slot;
          ++nload;
          mbar_arrive_expect_tx(wfull_bar, (uint32_t)(nkb * kStageB));
          for (int kb = 0; kb < nkb; ++kb) {
            for (int i = 0; i < sp.nx; ++i)
              tma_load_2d(wbase + (kb * sp.nx + i) * BN * 128, &tmap_w, sp.skoff[st] + i * p.C + cb * kBK, wrow, wfull_bar);
            if (++st == sp.n_strips) { st = 0; ++cb; }
          }
          continue;
        }
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % ns;
          if (it >= ns) mbar_wait(empty_bar(s), ((it / ns) - 1) & 1, p.err_flag, 4);
          mbar_arrive_expect_tx(full_bar(s), (uint32_t)kStageB);
          for (int i = 0; i < sp.nx; ++i)
            tma_load_2d(base + s * kStage + kSStageA + i * BN * 128, &tmap_w, sp.skoff[st] + i * p.C + cb * kBK, wrow, full_bar(s));
          if (++st == sp.n_strips) { st = 0; ++cb; }
        }
      }
    }
  } else if (warp == kGW + 1) {
    // ============================================================================= MMA ISSUER
    const uint32_t idesc = make_idesc(BN, false, false);
    uint32_t it = 0, tcount = 0, nload = 0;
    int cur_slot = -1;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t buf = tcount & 1, use = tcount >> 1;
      if (use >= 1) mbar_wait(tempty_bar(buf), (use - 1) & 1, p.err_flag, 5);
      bool last_of_slot = false;
      if (resident) {
        TileInfo ti, tn;
        decode_tile(tile, p, s_tiles, s_grp, kBM, ti);
        if (ti.slot != cur_slot) {
          mbar_wait(wfull_bar, nload & 1, p.err_flag, 7);
          cur_slot = ti.slot;
          ++nload;
        }
        last_of_slot = tile + (int)gridDim.x < total_tiles && decode_tile(tile + gridDim.x, p, s_tiles, s_grp, kBM, tn) &&
                       tn.slot != cur_slot;
      }
      tc_fence_after();
      const uint32_t tacc = tmem_base + buf * (uint32_t)BN;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % ns;
        mbar_wait(full_bar(s), (it / ns) & 1, p.err_flag, 2);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = base + s * kStage;
          const uint32_t sb = resident ? wbase + (uint32_t)(kb * sp.nx) * BN * 128 : sa + kSStageA;
          for (int i = 0; i < sp.nx; ++i) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_bf16(tacc, make_desc_nosw(sa + (uint32_t)i * 16u + (uint32_t)k * 2u * kSLbo, kSLbo, 128),
                        make_desc(sb + i * BN * 128 + k * 32, 16, 1024), idesc, (kb | i | k) ? 1u : 0u);
          }
          umma_commit(empty_bar(s));
          if (kb == nkb - 1) {
            umma_commit(tfull_bar(buf));
            if (last_of_slot) umma_commit(wempty_bar);       // the weight buffer may be overwritten once these MMAs retire
          }
        }
        __syncwarp();
      }
    }
    tc_fence_before();
  } else {
    // ============================================================================= EPILOGUE (masked padded columns)
    const int q = warp & 3;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      TileInfo ti;
      decode_tile(tile, p, s_tiles, s_grp, kBM, ti);
      const uint32_t buf = tcount & 1, use = tcount >> 1;
      mbar_wait(tfull_bar(buf), use & 1, p.err_flag, 3);
      tc_fence_after();
      const float* bias = p.bias ? p.bias + (long)ti.slot * p.bias_slot_stride + ti.n0 : nullptr;
      uint32_t r[32];
      const uint32_t t_lane = tmem_base + buf * (uint32_t)BN + ((uint32_t)(q * 32) << 16);
      const int m = ti.m0 + q * 32 + lane;
      bool ok = m < ti.rows * sp.Pp;
      const int smp = ok ? m / sp.Pp : 0, rem = ok ? m - smp * sp.Pp : 0;
      const int oa = rem / sp.Wp, ob = rem - oa * sp.Wp;
      ok = ok && ob < p.Wo;
      const long opix = (long)(ti.row_start + smp) * p.P_full + (oa * p.o_my + p.o_oy) * p.Wo_full + ob * p.o_mx + p.o_ox;
      __nv_bfloat16* yrow = p.out + (opix * p.Nout + ti.n0);
      for (int c = 0; c < BN; c += 32) {
        tmem_ld32(t_lane + c, r);
        if (ok) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(r[j]) + (bias ? __ldg(bias + c + j) : 0.f);
          uint4* dst = reinterpret_cast<uint4*>(yrow + c);
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) dst[qq] = pack8(f + 8 * qq);
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(buf));
    }
  }
  __syncthreads();
  if (warp == kGW + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Weight gradient of a convolution:  dw[slot][n][kk] += sum_pix dy[pix, n] * im2col(x)[pix, kk]   (kk = (tap, channel))
// D tile = 128 kk (M) x N (<= 256); the reduction runs over the group's pixels in blocks of 64.  Both operands are
// MN-major (rows = pixels): A rows are gathered 128-byte channel slices of x (line-coalesced cp.async, zero-filled for
// padding and past the group's end), B rows are dy, fetched by TMA as N/64 boxes of {64 n, 64 pixels}.
// Work units = (group, split of the pixel range, 128-wide kk tile); units with the same pixel range are adjacent so
// that the CTAs running together share dy in L2.  fp32 atomics (RED) combine the splits.
struct WgParams {
  const es_group* grp;
  int n_groups;
  int Hs, Ws, C, Hu, Wu, Ho, Wo, P;
  int n_taps, my, mx;               // tap t gathers x at (oy*my + tdy[t], ox*mx + tdx[t]); its C columns start at tcol[t] of a dw row
  signed char tdy[32], tdx[32];
  int tcol[32];
  int dw_ld;                        // length of a dw row (all taps of all phases)
  int N, KK, tiles_m, splits;
  int kmode;                        // 0: all pixel blocks of a group; 1: only its FULL 64-pixel blocks; 2: only the partial last one;
                                    // 3: the pixels the strip kernel left over (see igemm_wgrad_strip_kernel): [pix_lo, rows*P)
  int s_Wp, s_Pp;                   // kmode 3: padded row pitch / pixels per sample of the strip kernel's enumeration
  int tile_m;                       // kk columns per work unit: 128 (single CTA) or 256 (CTA pair)
  unsigned char ymap[64], xmap[64];
  const __nv_bfloat16* x;
  float* dw;
  long dw_slot_stride;
  int* err_flag;
};

struct WgTile {
  int m0, rows, row_start, slot, kb0, kb1;
  int pix_lo;                       // pixels below this index (group-relative) are not reduced (kmode 3), else 0
};

__device__ __forceinline__ bool wg_decode(int t, const WgParams& p, const es_group* s_grp, WgTile& ti) {
  const int per_g = p.splits * p.tiles_m;
  const int g = t / per_g, rem = t - g * per_g;
  const int sp = rem / p.tiles_m, mt = rem - sp * p.tiles_m;
  ti.m0 = mt * (p.tile_m ? p.tile_m : kBM);
  ti.rows = s_grp[g].rows; ti.row_start = s_grp[g].row_start; ti.slot = s_grp[g].slot;
  const int nall = ceil_div(ti.rows * p.P, kBK), nfull = (ti.rows * p.P) / kBK;
  ti.pix_lo = 0;
  if (p.kmode == 3) {               // valid pixels at or after the last full 64-position block of the PADDED enumeration
    const int pos = ((ti.rows * p.s_Pp) / kBK) * kBK;
    const int smp = pos / p.s_Pp, rem = pos - smp * p.s_Pp;
    const int oy = rem / p.s_Wp, t = rem - oy * p.s_Wp;
    ti.pix_lo = smp * p.P + oy * p.Wo + min(t, p.Wo);
    ti.kb0 = ti.pix_lo / kBK;
    ti.kb1 = nall;
    return ti.pix_lo < ti.rows * p.P;
  }
  if (p.kmode == 2) {               // the partial last block only (one unit per (group, kk tile); splits == 1)
    ti.kb0 = nfull;
    ti.kb1 = nall;
    return ti.kb1 > ti.kb0;
  }
  const int nkb = p.kmode == 1 ? nfull : nall;
  const int per = ceil_div(nkb, p.splits);
  ti.kb0 = sp * per;
  ti.kb1 = min(nkb, ti.kb0 + per);
  return ti.kb1 > ti.kb0;
}

__global__ void __launch_bounds__(kGThreads, 1)
igemm_wgrad_kernel(const __grid_constant__ WgParams p, const __grid_constant__ CUtensorMap tmap_dy) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_base = base + kFStages * kFStage;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kFStages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * kFStages + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * kFStages + 2 + b); };
  uint8_t* gen = smem_raw + (bar_base - raw);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + 8 * (2 * kFStages + 4));
  const uint32_t tmem_slot = bar_base + 8u * (2 * kFStages + 4);
  es_group* s_grp = reinterpret_cast<es_group*>(gen + 384);
  unsigned char* s_ymap = gen + 384 + 1024;
  unsigned char* s_xmap = s_ymap + 64;

  const int BN = p.N;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * BN) tmem_cols <<= 1;

  if (tid < p.n_groups) s_grp[tid] = p.grp[tid];
  if (tid < 64) { s_ymap[tid] = p.ymap[tid]; s_xmap[tid] = p.xmap[tid]; }
  if (tid == 0) {
    for (int s = 0; s < kFStages; ++s) {
      mbar_init(full_bar(s), kGLoaders + 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kGW && lane == 0) tma_prefetch_desc(&tmap_dy);
  if (warp == kGW + 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int total_tiles = p.n_groups * p.splits * p.tiles_m;
  const int nseg = BN >> 6;

  if (warp < kGW) {
    // =========================================================================== A GATHER (8 warps, see igemm_fwd_kernel)
    constexpr int PR = 64 / (kGLoaders / 8);        // pixel rows per thread
    constexpr int PSTEP = kGLoaders / 8;
    const int chunk = tid & 7, rsub = tid >> 3;     // pixel rows rsub + PSTEP*i (i < PR), both 64-wide kk segments
    uint32_t it = 0, signalled = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      WgTile ti;
      if (!wg_decode(tile, p, s_grp, ti)) continue;
      int kyv[2], kxv[2], cv[2];
#pragma unroll
      for (int sg = 0; sg < 2; ++sg) {
        const int kk = ti.m0 + sg * 64;
        const int tap = kk / p.C;
        cv[sg] = kk - tap * p.C + chunk * 8;
        kyv[sg] = p.tdy[tap];
        kxv[sg] = p.tdx[tap];
      }
      const int hw = p.Hs * p.Ws;
      const long src_base = (long)ti.row_start * hw;
      // (sample, oy, ox) of this thread's 4 pixel rows: a mixed-radix counter advanced by 64 pixels per k-block, so the
      // main loop has no integer division
      int smp[PR], oyv[PR], oxv[PR];
#pragma unroll
      for (int i = 0; i < PR; ++i) {
        const int pidx = ti.kb0 * kBK + rsub + PSTEP * i;
        smp[i] = pidx / p.P;
        const int pix = pidx - smp[i] * p.P;
        oyv[i] = pix / p.Wo;
        oxv[i] = pix - oyv[i] * p.Wo;
      }
      const int dy64 = kBK / p.Wo, dx64 = kBK - dy64 * p.Wo;
      for (int kb = ti.kb0; kb < ti.kb1; ++kb, ++it) {
        const int s = it % kFStages;
        if (it >= kFStages) mbar_wait(empty_bar(s), ((it / kFStages) - 1) & 1, p.err_flag, 1);
        const uint32_t sa = base + s * kFStage;
#pragma unroll
        for (int i = 0; i < PR; ++i) {
          const int prow = rsub + PSTEP * i;
          const bool v = smp[i] < ti.rows && kb * kBK + prow >= ti.pix_lo;
          const int oy = oyv[i], ox = oxv[i];
          const long sbase = src_base + (long)smp[i] * hw;
          const uint32_t doff = (uint32_t)prow * 128u + (uint32_t)((chunk ^ (prow & 7)) << 4);
#pragma unroll
          for (int sg = 0; sg < 2; ++sg) {
            const int uy = oy * p.my + kyv[sg], ux = ox * p.mx + kxv[sg];
            const bool inb = v && uy >= 0 && uy < p.Hu && ux >= 0 && ux < p.Wu;
            const int sy = inb ? s_ymap[uy] : 0, sx = inb ? s_xmap[ux] : 0;
            const __nv_bfloat16* src = p.x + (inb ? ((sbase + sy * p.Ws + sx) * p.C + cv[sg]) : 0L);
            cp_async16(sa + sg * 8192u + doff, src, inb);
          }
          oxv[i] += dx64;
          oyv[i] += dy64;
          if (oxv[i] >= p.Wo) { oxv[i] -= p.Wo; ++oyv[i]; }
          while (oyv[i] >= p.Ho) { oyv[i] -= p.Ho; ++smp[i]; }
        }
        cp_async_commit();
        if (it - signalled >= (uint32_t)kFLag) {
          cp_async_wait<kFLag>();
          fence_proxy_async();
          mbar_arrive(full_bar(signalled % kFStages));
          ++signalled;
        }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    while (signalled < it) {
      mbar_arrive(full_bar(signalled % kFStages));
      ++signalled;
    }
  } else if (warp == kGW) {
    // =========================================================================== TMA PRODUCER (dy)
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        WgTile ti;
        if (!wg_decode(tile, p, s_grp, ti)) continue;
        const int prow0 = ti.row_start * p.P;
        for (int kb = ti.kb0; kb < ti.kb1; ++kb, ++it) {
          const int s = it % kFStages;
          if (it >= kFStages) mbar_wait(empty_bar(s), ((it / kFStages) - 1) & 1, p.err_flag, 4);
          mbar_arrive_expect_tx(full_bar(s), (uint32_t)BN * 128u);
          const uint32_t sb = base + s * kFStage + kFStageA;
          for (int sg = 0; sg < nseg; ++sg) tma_load_2d(sb + sg * 8192u, &tmap_dy, sg * 64, prow0 + kb * kBK, full_bar(s));
        }
      }
    }
  } else if (warp == kGW + 1) {
    // =========================================================================== MMA ISSUER
    const uint32_t idesc = make_idesc(BN, true, true);
    uint32_t it = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      WgTile ti;
      if (!wg_decode(tile, p, s_grp, ti)) continue;
      const uint32_t buf = tcount & 1;
      if (tcount >= 2) mbar_wait(tempty_bar(buf), ((tcount >> 1) - 1) & 1, p.err_flag, 5);
      tc_fence_after();
      const uint32_t tacc = tmem_base + buf * (uint32_t)BN;
      for (int kb = ti.kb0; kb < ti.kb1; ++kb, ++it) {
        const int s = it % kFStages;
        mbar_wait(full_bar(s), (it / kFStages) & 1, p.err_flag, 2);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = base + s * kFStage;
          const uint32_t sb = sa + kFStageA;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_bf16(tacc, make_desc(sa + k * 2048, 8192, 1024), make_desc(sb + k * 2048, 8192, 1024), idesc,
                      (kb > ti.kb0 || k) ? 1u : 0u);
          umma_commit(empty_bar(s));
          if (kb == ti.kb1 - 1) umma_commit(tfull_bar(buf));
        }
        __syncwarp();
      }
      ++tcount;
    }
    tc_fence_before();
  } else {
    // =========================================================================== EPILOGUE (warps 6-9): RED into dw
    const int q = warp & 3;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      WgTile ti;
      if (!wg_decode(tile, p, s_grp, ti)) continue;
      const uint32_t buf = tcount & 1;
      mbar_wait(tfull_bar(buf), (tcount >> 1) & 1, p.err_flag, 3);
      tc_fence_after();
      const uint32_t t_lane = tmem_base + buf * (uint32_t)BN + ((uint32_t)(q * 32) << 16);
      const int kk = ti.m0 + q * 32 + lane;                    // this thread's reduction column inside the launch's KK
      const int tap = kk / p.C;
      float* dw = p.dw + (long)ti.slot * p.dw_slot_stride + p.tcol[tap] + (kk - tap * p.C);
      uint32_t r[32];
      for (int c = 0; c < BN; c += 32) {
        tmem_ld32(t_lane + c, r);
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(dw + (long)(c + j) * p.dw_ld, __uint_as_float(r[j]));
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(buf));
      ++tcount;
    }
  }
  __syncthreads();
  if (warp == kGW + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// TMA-FED weight gradient, as a CTA pair (kPair: D tile = 256 kk x N, 128 kk per CTA, tcgen05.mma.cta_group::2) or as a
// single CTA (128 kk x N; used for N = 64, where half of dy per CTA would be a 64-byte row).  Both operands by TMA: A = two
// im2col boxes per CTA and pixel block (64 pixels x 64 channels of x for the (tap, channel block) of each 64-wide kk segment —
// MN-major rows of 128 B, exactly what the cp.async gather used to build), B = dy (this CTA's N/2 channels of it in a pair) x
// 64 pixels.  One elected thread feeds the pipeline; there are no gather warps and no L1TEX traffic.  im2col rows cannot be
// zero-filled past a group's end (they run into the next sample), so this kernel reduces over the FULL 64-pixel blocks of a
// group only (kmode 1); the partial last block of each group goes through the gather kernel (kmode 2, one k-block per
// unit), which zero-fills.  Barrier protocol as igemm_tma_pair_kernel.  Launched with a cluster dimension of 2 (kPair) or 1.
template <int kStages, bool kPair>
__global__ void __launch_bounds__(192, 1)
igemm_wgrad_tma_kernel(const __grid_constant__ WgParams p, const __grid_constant__ TmaAParams ta,
                       const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x) {
  constexpr int kStage = kFStageA + 128 * 128;   // A 16 KB + dy (half) <= 16 KB
  extern __shared__ uint8_t smem_raw[];
  // (shuffle: the warp index becomes warp-uniform for the compiler, so the MMA issuer's descriptors live in uniform registers)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_base = base + kStages * kStage;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (16 + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (18 + b); };
  uint8_t* gen = smem_raw + (bar_base - raw);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + 8 * 20);
  const uint32_t tmem_slot = bar_base + 8u * 20;
  es_group* s_grp = reinterpret_cast<es_group*>(gen + 512);

  const int BN = p.N;
  const int nseg_c = kPair ? (BN >> 7) : (BN >> 6);      // 64-channel dy segments this CTA loads
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * BN) tmem_cols <<= 1;

  if (tid < p.n_groups) s_grp[tid] = p.grp[tid];
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), kPair ? 8 : 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap_dy); tma_prefetch_desc(&tmap_x); }
  if (warp == 1) { if (kPair) tmem_alloc_pair(tmem_slot, tmem_cols); else tmem_alloc(tmem_slot, tmem_cols); }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int total_tiles = p.n_groups * p.splits * p.tiles_m;
  const int first = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, stride = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0) {
    // =========================================================================== TMA PRODUCER
    if (lane == 0) {
      const uint32_t lead_full0 = kPair ? mapa_shared(full_bar(0), 0) : full_bar(0);
      uint32_t it = 0;
      for (int tile = first; tile < total_tiles; tile += stride) {
        WgTile ti;
        if (!wg_decode(tile, p, s_grp, ti)) continue;
        int c0[2], ow[2], oh[2];
#pragma unroll
        for (int sg = 0; sg < 2; ++sg) {
          const int kk = ti.m0 + (int)rank * kBM + sg * 64;
          const int tap = kk / p.C;
          c0[sg] = kk - tap * p.C;
          ow[sg] = p.tdx[tap] - ta.low_w;
          oh[sg] = p.tdy[tap] - ta.low_h;
        }
        const int prow0 = ti.row_start * p.P;
        // (sample, oy, ox) of the block's first pixel: a mixed-radix counter advanced by 64 pixels per block
        const int pidx = ti.kb0 * kBK;
        int smp = pidx / p.P;
        const int pix = pidx - smp * p.P;
        int oy = pix / p.Wo, ox = pix - oy * p.Wo;
        const int dy64 = kBK / p.Wo, dx64 = kBK - dy64 * p.Wo;
        for (int kb = ti.kb0; kb < ti.kb1; ++kb, ++it) {
          const int s = it % kStages;
          if (it >= (uint32_t)kStages) mbar_wait(empty_bar(s), ((it / kStages) - 1) & 1, p.err_flag, 4);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(s), (kPair ? 2u : 1u) * (uint32_t)kFStageA + (uint32_t)BN * 128u);
          const uint32_t lead_full = lead_full0 + 8u * s;
          const uint32_t sa = base + s * kStage;
          const int cn = ti.row_start + smp, ch = ta.low_h + oy * p.my, cw = ta.low_w + ox * p.mx;
#pragma unroll
          for (int sg = 0; sg < 2; ++sg) {
            if (kPair) tma_im2col_4d_pair(sa + sg * 8192u, &tmap_x, c0[sg], cw, ch, cn, (uint32_t)ow[sg], (uint32_t)oh[sg], lead_full);
            else tma_im2col_4d(sa + sg * 8192u, &tmap_x, c0[sg], cw, ch, cn, (uint32_t)ow[sg], (uint32_t)oh[sg], lead_full);
          }
          for (int sg = 0; sg < nseg_c; ++sg) {
            const int col = ((int)rank * nseg_c + sg) * 64;
            if (kPair) tma_load_2d_pair(sa + kFStageA + sg * 8192u, &tmap_dy, col, prow0 + kb * kBK, lead_full);
            else tma_load_2d(sa + kFStageA + sg * 8192u, &tmap_dy, col, prow0 + kb * kBK, lead_full);
          }
          ox += dx64; oy += dy64;
          if (ox >= p.Wo) { ox -= p.Wo; ++oy; }
          while (oy >= p.Ho) { oy -= p.Ho; ++smp; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================================================================== MMA ISSUER (leader CTA only)
    if (rank == 0) {
      const uint32_t idesc = make_idesc_m(BN, kPair ? 2 * kBM : kBM, true, true);
      uint32_t it = 0, tcount = 0;
      for (int tile = first; tile < total_tiles; tile += stride) {
        WgTile ti;
        if (!wg_decode(tile, p, s_grp, ti)) continue;
        const uint32_t buf = tcount & 1;
        if (tcount >= 2) mbar_wait(tempty_bar(buf), ((tcount >> 1) - 1) & 1, p.err_flag, 5);
        tc_fence_after();
        const uint32_t tacc = __shfl_sync(0xffffffffu, tmem_base, 0) + buf * (uint32_t)BN;
        const int kb0 = __shfl_sync(0xffffffffu, ti.kb0, 0), kb1 = __shfl_sync(0xffffffffu, ti.kb1, 0);   // uniform loop bounds
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % kStages;
          mbar_wait(full_bar(s), (it / kStages) & 1, p.err_flag, 2);
          tc_fence_after();
          const uint32_t sa = base + s * kStage;          // descriptors from warp-uniform values, outside the lane-0 branch
          const uint64_t a0 = make_desc(sa, 8192, 1024), b0 = make_desc(sa + kFStageA, 8192, 1024);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint32_t acc = (kb > kb0 || k) ? 1u : 0u;
            if (kPair) umma_bf16_pair_elect(tacc, a0 + 128 * k, b0 + 128 * k, idesc, acc); else umma_bf16_elect(tacc, a0 + 128 * k, b0 + 128 * k, idesc, acc);
          }
          if (kPair) {
            umma_commit_pair_elect(empty_bar(s), 3);
            if (kb == kb1 - 1) umma_commit_pair_elect(tfull_bar(buf), 3);
          } else {
            umma_commit_elect(empty_bar(s));
            if (kb == kb1 - 1) umma_commit_elect(tfull_bar(buf));
          }
          __syncwarp();
        }
        ++tcount;
      }
      tc_fence_before();
    }
  } else {
    // =========================================================================== EPILOGUE (warps 2-5): RED into dw
    const int q = warp & 3;
    uint32_t tcount = 0;
    for (int tile = first; tile < total_tiles; tile += stride) {
      WgTile ti;
      if (!wg_decode(tile, p, s_grp, ti)) continue;
      const uint32_t buf = tcount & 1;
      mbar_wait(tfull_bar(buf), (tcount >> 1) & 1, p.err_flag, 3);
      tc_fence_after();
      const uint32_t t_lane = tmem_base + buf * (uint32_t)BN + ((uint32_t)(q * 32) << 16);
      const int kk = ti.m0 + (int)rank * kBM + q * 32 + lane;
      const int tap = kk / p.C;
      float* dw = p.dw + (long)ti.slot * p.dw_slot_stride + p.tcol[tap] + (kk - tap * p.C);
      uint32_t r[32];
      for (int c = 0; c < BN; c += 32) {
        tmem_ld32(t_lane + c, r);
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(dw + (long)(c + j) * p.dw_ld, __uint_as_float(r[j]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (kPair) mbar_arrive_cluster(mapa_shared(tempty_bar(buf), 0)); else mbar_arrive(tempty_bar(buf)); }
      ++tcount;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc_pair(tmem_base, tmem_cols); else tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// TMA-fed weight gradient with TAP-ROW STRIPS.  igemm_wgrad_tma_kernel moves 16 KB of x + 8 KB of dy per CTA for every
// 128 x N x 64 MACs: at N <= 128 the L2->SM fill (~41 B/clk/SM) is 2-4x what the MMA of that block takes (conv2: 0.75, conv3:
// 0.40 PFLOP/s).  With the reduction (pixel) axis enumerated on the padded row pitch Wp = Wo + nx - 1 (see
// igemm_tma_strip_kernel) the nx taps of a tap row read ONE strip of 64 + nx - 1 pixels — tap j's A operand is the same
// MN-major tile started j rows (j * 128 B) further down — and share the dy tile: a work unit is (group, split, tap row, 128
// channels per CTA) with nx accumulators [128 x N] side by side in TMEM.  dy is read on the same padded pitch by an im2col
// box whose upper corner reaches nx - 1 columns past the image (zero-filled): the pad positions contribute nothing.
// Only the full 64-position blocks of a group are reduced here; the remaining pixels go through the gather kernel (kmode 3).
struct WgStripParams {
  int n_rows, nx, Wp, Pp;           // tap rows, taps per row, padded pitch, Ho * Wp
  int cchunks;                      // channel chunks per tap row: C / 128 (single CTA) or C / 256 (pair)
  int stages, stage_bytes, seg_bytes;   // seg_bytes: one 64-channel strip segment, 1 KB aligned
  int low_w, low_h;
  signed char sdy[16], sdx[16];
  int tap0[16];                     // index (into WgParams::tcol) of the row's first tap; tap j of the row is tap0 + j
};

template <bool kPair>
__global__ void __launch_bounds__(192, 1)
igemm_wgrad_strip_kernel(const __grid_constant__ WgParams p, const __grid_constant__ WgStripParams ws,
                         const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x) {
  extern __shared__ uint8_t smem_raw[];
  // (shuffle: the warp index becomes warp-uniform for the compiler, so the MMA issuer's descriptors live in uniform registers)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int kStages = ws.stages, kStage = ws.stage_bytes;
  const uint32_t bar_base = base + kStages * kStage;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (16 + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (18 + b); };
  uint8_t* gen = smem_raw + (bar_base - raw);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + 8 * 20);
  const uint32_t tmem_slot = bar_base + 8u * 20;
  es_group* s_grp = reinterpret_cast<es_group*>(gen + 512);

  const int BN = p.N, nx = ws.nx;
  const int nseg_c = kPair ? (BN >> 7) : (BN >> 6);      // 64-channel dy segments this CTA loads
  const uint32_t acc_cols = (uint32_t)(nx * BN);         // nx accumulators side by side
  const uint32_t nbuf = 2 * acc_cols <= 512 ? 2u : 1u;
  uint32_t tmem_cols = 32;
  while (tmem_cols < nbuf * acc_cols) tmem_cols <<= 1;

  if (tid < p.n_groups) s_grp[tid] = p.grp[tid];
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), kPair ? 8 : 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap_dy); tma_prefetch_desc(&tmap_x); }
  if (warp == 1) { if (kPair) tmem_alloc_pair(tmem_slot, tmem_cols); else tmem_alloc(tmem_slot, tmem_cols); }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int units_m = ws.n_rows * ws.cchunks;
  const int total_tiles = p.n_groups * p.splits * units_m;
  const int first = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, stride = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const uint32_t strip_bytes = (uint32_t)(kBK + nx - 1) * 128u;

  // unit -> (group, split, tap row, channel chunk) and its range of full padded 64-position blocks
  struct Unit { int rows, row_start, slot, kb0, kb1, rr, cc; };
  auto decode = [&](int t, Unit& u) -> bool {
    const int per_g = p.splits * units_m;
    const int g = t / per_g, rem = t - g * per_g;
    const int sp = rem / units_m, mu = rem - sp * units_m;
    u.rr = mu / ws.cchunks; u.cc = mu - u.rr * ws.cchunks;
    u.rows = s_grp[g].rows; u.row_start = s_grp[g].row_start; u.slot = s_grp[g].slot;
    const int nkb = (u.rows * ws.Pp) / kBK;
    const int per = ceil_div(nkb, p.splits);
    u.kb0 = sp * per;
    u.kb1 = min(nkb, u.kb0 + per);
    return u.kb1 > u.kb0;
  };

  if (warp == 0) {
    // =========================================================================== TMA PRODUCER
    // (one lane runs the whole loop.  The warp-uniform / elected-lane form that helps igemm_tma_strip_kernel was measured here
    // too: conv3 weight gradient 0.562 -> 0.616 ms, conv2 0.249 -> 0.263 — not used)
    if (lane == 0) {
      const uint32_t lead_full0 = kPair ? mapa_shared(full_bar(0), 0) : full_bar(0);
      uint32_t it = 0;
      for (int tile = first; tile < total_tiles; tile += stride) {
        Unit u;
        if (!decode(tile, u)) continue;
        const int cbase = (u.cc * (kPair ? 2 : 1) + (int)rank) * kBM;     // this CTA's 128 channels
        const uint32_t ow = (uint32_t)(ws.sdx[u.rr] - ws.low_w), oh = (uint32_t)(ws.sdy[u.rr] - ws.low_h);
        const int pidx = u.kb0 * kBK;
        int smp = pidx / ws.Pp;
        const int pix = pidx - smp * ws.Pp;
        int oy = pix / ws.Wp, ot = pix - oy * ws.Wp;
        const int dy64 = kBK / ws.Wp, dx64 = kBK - dy64 * ws.Wp;
        for (int kb = u.kb0; kb < u.kb1; ++kb, ++it) {
          const int s = it % kStages;
          if (it >= (uint32_t)kStages) mbar_wait(empty_bar(s), ((it / kStages) - 1) & 1, p.err_flag, 4);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(s), (kPair ? 2u : 1u) * 2u * strip_bytes + (uint32_t)BN * 128u);
          const uint32_t lead_full = lead_full0 + 8u * s;
          const uint32_t sa = base + s * kStage;
          const uint32_t sb = sa + 2u * ws.seg_bytes;
          const int cn = u.row_start + smp;
#pragma unroll
          for (int sg = 0; sg < 2; ++sg) {
            if (kPair) tma_im2col_4d_pair(sa + sg * ws.seg_bytes, &tmap_x, cbase + sg * 64, ws.low_w + ot, ws.low_h + oy * p.my, cn, ow, oh, lead_full);
            else tma_im2col_4d(sa + sg * ws.seg_bytes, &tmap_x, cbase + sg * 64, ws.low_w + ot, ws.low_h + oy * p.my, cn, ow, oh, lead_full);
          }
          for (int sg = 0; sg < nseg_c; ++sg) {
            const int col = ((int)rank * nseg_c + sg) * 64;
            if (kPair) tma_im2col_4d_pair(sb + sg * 8192u, &tmap_dy, col, ot, oy, cn, 0u, 0u, lead_full);
            else tma_im2col_4d(sb + sg * 8192u, &tmap_dy, col, ot, oy, cn, 0u, 0u, lead_full);
          }
          ot += dx64; oy += dy64;
          if (ot >= ws.Wp) { ot -= ws.Wp; ++oy; }
          while (oy >= p.Ho) { oy -= p.Ho; ++smp; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================================================================== MMA ISSUER (leader CTA only)
    if (rank == 0) {
      const uint32_t idesc = make_idesc_m(BN, kPair ? 2 * kBM : kBM, true, true);
      uint32_t it = 0, tcount = 0;
      for (int tile = first; tile < total_tiles; tile += stride) {
        Unit u;
        if (!decode(tile, u)) continue;
        const uint32_t buf = nbuf == 2 ? (tcount & 1) : 0u;
        const uint32_t use = nbuf == 2 ? (tcount >> 1) : tcount;
        if (use >= 1) mbar_wait(tempty_bar(buf), (use - 1) & 1, p.err_flag, 5);
        tc_fence_after();
        const uint32_t tacc = __shfl_sync(0xffffffffu, tmem_base, 0) + buf * acc_cols;
        const int kb0 = __shfl_sync(0xffffffffu, u.kb0, 0), kb1 = __shfl_sync(0xffffffffu, u.kb1, 0);   // uniform loop bounds
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % kStages;
          mbar_wait(full_bar(s), (it / kStages) & 1, p.err_flag, 2);
          tc_fence_after();
          const uint32_t sa = base + s * kStage;          // descriptors from warp-uniform values, outside the lane-0 branch
          const uint64_t a0 = make_desc(sa, (uint32_t)ws.seg_bytes, 1024), b0 = make_desc(sa + 2u * ws.seg_bytes, 8192, 1024);
          const uint32_t acc0 = kb > kb0 ? 1u : 0u;
          for (int j = 0; j < nx; ++j) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              const uint64_t ad = a0 + (uint64_t)(8 * j + 128 * k), bd = b0 + 128 * k;
              if (kPair) umma_bf16_pair_elect(tacc + j * BN, ad, bd, idesc, acc0 | (k ? 1u : 0u));
              else umma_bf16_elect(tacc + j * BN, ad, bd, idesc, acc0 | (k ? 1u : 0u));
            }
          }
          if (kPair) {
            umma_commit_pair_elect(empty_bar(s), 3);
            if (kb == kb1 - 1) umma_commit_pair_elect(tfull_bar(buf), 3);
          } else {
            umma_commit_elect(empty_bar(s));
            if (kb == kb1 - 1) umma_commit_elect(tfull_bar(buf));
          }
          __syncwarp();
        }
        ++tcount;
      }
      tc_fence_before();
    }
  } else {
    // =========================================================================== EPILOGUE (warps 2-5): RED into dw
    const int q = warp & 3;
    uint32_t tcount = 0;
    for (int tile = first; tile < total_tiles; tile += stride) {
      Unit u;
      if (!decode(tile, u)) continue;
      const uint32_t buf = nbuf == 2 ? (tcount & 1) : 0u;
      const uint32_t use = nbuf == 2 ? (tcount >> 1) : tcount;
      mbar_wait(tfull_bar(buf), use & 1, p.err_flag, 3);
      tc_fence_after();
      const int ch = (u.cc * (kPair ? 2 : 1) + (int)rank) * kBM + q * 32 + lane;      // this thread's input channel
      uint32_t r[32];
      for (int j = 0; j < nx; ++j) {
        const uint32_t t_lane = tmem_base + buf * acc_cols + (uint32_t)(j * BN) + ((uint32_t)(q * 32) << 16);
        float* dw = p.dw + (long)u.slot * p.dw_slot_stride + p.tcol[ws.tap0[u.rr] + j] + ch;
        for (int c = 0; c < BN; c += 32) {
          tmem_ld32(t_lane + c, r);
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) atomicAdd(dw + (long)(c + jj) * p.dw_ld, __uint_as_float(r[jj]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (kPair) mbar_arrive_cluster(mapa_shared(tempty_bar(buf), 0)); else mbar_arrive(tempty_bar(buf)); }
      ++tcount;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc_pair(tmem_base, tmem_cols); else tmem_dealloc(tmem_base, tmem_cols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeIm2colFn encode_im2col_fn() {
  static EncodeIm2colFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeIm2colFn>(ptr);
  }
  return fn;
}

int* fwd_err_flag() { return pipeline_err_flag(); }

}  // namespace

void fill_maps_uc(int Hs, int Ws, int Hu, int Wu, unsigned char* ymap, unsigned char* xmap) {
  // torch 'nearest': src = min(floor(dst * (in/out as float)), in-1)
  const float sy = (float)Hs / (float)Hu, sx = (float)Ws / (float)Wu;
  for (int i = 0; i < 64; ++i) {
    const int y = (int)floorf((float)i * sy), x = (int)floorf((float)i * sx);
    ymap[i] = (unsigned char)(y < Hs - 1 ? y : Hs - 1);
    xmap[i] = (unsigned char)(x < Ws - 1 ? x : Ws - 1);
  }
}

}  // namespace es

using namespace es;

// Host-side plan of the strip variant: fills `sp` and returns true when the tap table of `p` (BN, n_tiles_n already set) is a
// set of tap rows the strip kernel can run (see igemm_strip_kernel).  Also exported through es_igemm_fwd_plan (CPU-testable).
static bool strip_plan(const FwdParams& p, long total_rows, StripParams& sp) {
  static const bool strip_on = [] { const char* e = getenv("ES_IGEMM_STRIP"); return !(e && e[0] == '0'); }();
  sp = StripParams{};
  bool okk = strip_on && p.Hu == p.Hs && p.Wu == p.Ws && p.my == 1 && p.mx == 1 && p.BN <= 128 && p.n_taps >= 2;
  if (okk) {
    int t = 0;
    while (okk && t < p.n_taps) {
      int n = 1;
      while (t + n < p.n_taps && p.tdy[t + n] == p.tdy[t] && p.tdx[t + n] == p.tdx[t] + n && p.tkoff[t + n] == p.tkoff[t] + n * p.C) ++n;
      if (sp.n_strips == 0) { sp.nx = n; sp.dx0 = p.tdx[t]; }
      okk = sp.n_strips < 8 && n == sp.nx && p.tdx[t] == sp.dx0;
      if (okk) { sp.sdy[sp.n_strips] = p.tdy[t]; sp.skoff[sp.n_strips] = p.tkoff[t]; ++sp.n_strips; }
      t += n;
    }
    okk = okk && sp.nx >= 2 && sp.nx <= 4 && sp.nx * p.BN <= 384;
  }
  if (okk) {
    sp.Wp = p.Wo + sp.nx - 1;
    sp.Pp = p.Ho * sp.Wp;
    okk = sp.Wp < 256 && total_rows * sp.Pp < 2147483647L;
    static const bool res_on = [] { const char* e = getenv("ES_IGEMM_STRIP_RESIDENT"); return e && e[0] == '1'; }();   // measured slower (conv3: 2.79 vs 2.54 ms): opt-in
    sp.stages = (!res_on && 4L * (kSStageA + sp.nx * p.BN * 128) <= (long)kSStages * (kSStageA + kSMaxB)) ? 4 : 3;
    sp.resident = res_on && p.n_tiles_n == 1 && (long)sp.n_strips * (p.C / kBK) * sp.nx * p.BN * 128 <= (long)kSResidentB;
  }
  return okk;
}

// widest N tile (power of two, >= 32) that divides N
static void pick_bn(FwdParams& p) {
  p.BN = 256;
  while (p.BN > 32 && p.Nout % p.BN != 0) p.BN >>= 1;
  p.n_tiles_n = p.Nout / p.BN;
}

// Host-side plan of the TMA-fed pair variant (igemm_tma_pair_kernel): true when the tap table reads the source directly
// (identity nearest maps), N tiles are >= 64 wide and the taps fit an im2col bounding box; fills the box corners
// (index 0 = W, 1 = H).  Base pixel of output index i along an axis = low + i*m with low <= the smallest tap offset; the
// upper corner makes the number of base pixels per row / column exactly Wo / Ho: Q = (W + up - low - 1) / m + 1; both
// corners are kept <= 0.  The dense product (one tap on a 1x1 grid) is excluded: measured slower (0.440 vs 0.389 ms).
static bool tma_pair_plan(const FwdParams& p, int low[2], int up[2]) {
  // the dense 1x1 product (fc2) through this kernel: measured 0.287 vs 0.293 ms with the single-CTA kernel once both have the
  // staged epilogue (bias row in shared memory, transposed stores) — neither the weight stream nor the gather is its bound;
  // opt-in (ES_FC2_TMA_PAIR=1)
  static const bool dense_ok = [] { const char* e = getenv("ES_FC2_TMA_PAIR"); return e && e[0] == '1'; }();
  if (!(p.BN >= 64 && p.Hu == p.Hs && p.Wu == p.Ws) || (!dense_ok && p.n_taps == 1 && p.Hs * p.Ws == 1)) return false;
  if (p.mx < 1 || p.mx > 8 || p.my < 1 || p.my > 8) return false;
  int dmin_x = 127, dmin_y = 127, dmax_x = -128, dmax_y = -128;
  for (int t = 0; t < p.n_taps; ++t) {
    dmin_x = p.tdx[t] < dmin_x ? p.tdx[t] : dmin_x; dmax_x = p.tdx[t] > dmax_x ? p.tdx[t] : dmax_x;
    dmin_y = p.tdy[t] < dmin_y ? p.tdy[t] : dmin_y; dmax_y = p.tdy[t] > dmax_y ? p.tdy[t] : dmax_y;
  }
  const int ext[2] = {p.Ws, p.Hs}, outn[2] = {p.Wo, p.Ho}, mul[2] = {p.mx, p.my}, dmin[2] = {dmin_x, dmin_y}, dmax[2] = {dmax_x, dmax_y};
  for (int a = 0; a < 2; ++a) {
    int lo = dmin[a] < 0 ? dmin[a] : 0;
    const int lim = ext[a] - 1 - (outn[a] - 1) * mul[a];      // up <= 0  <=>  low <= lim
    if (lo > lim) lo = lim;
    low[a] = lo;
    up[a] = (outn[a] - 1) * mul[a] + 1 + lo - ext[a];
    if (up[a] > 0 || lo < -128 || dmax[a] - lo > 255 || (ext[a] + up[a] - lo - 1) / mul[a] + 1 != outn[a]) return false;
  }
  return true;
}

// Host-side plan of the strip form of the TMA-fed pair variant (igemm_tma_strip_kernel).  The tap table must be rows of L taps
// (same dy, consecutive dx, weight columns C apart), all rows equally long; a row is cut into sub-strips of nx taps, nx the
// largest divisor of L whose stage (strip + nx half weight boxes) still leaves a 3-deep pipeline.  Taken only when the fill
// saved outweighs the masked pad columns by ES_TMA_STRIP_GAIN (default 1.15; 0 disables the variant).
static bool tma_strip_plan(const FwdParams& p, long total_rows, TStripParams& ts, int low[2], int up[2]) {
  static const double min_gain = [] { const char* e = getenv("ES_TMA_STRIP_GAIN"); return e ? atof(e) : 1.15; }();
  ts = TStripParams{};
  if (min_gain <= 0.0 || !(p.BN >= 64 && p.Hu == p.Hs && p.Wu == p.Ws && p.mx == 1 && p.my >= 1 && p.my <= 8 && p.n_taps >= 2)) return false;
  int L = 0, n_rows = 0, row_t[32];
  for (int t = 0; t < p.n_taps;) {
    int n = 1;
    while (t + n < p.n_taps && p.tdy[t + n] == p.tdy[t] && p.tdx[t + n] == p.tdx[t] + n && p.tkoff[t + n] == p.tkoff[t] + n * p.C) ++n;
    if (n_rows == 0) L = n;
    if (n != L) return false;
    row_t[n_rows++] = t;
    t += n;
  }
  if (L < 2) return false;
  const int BH = p.BN / 2;
  const int a_bytes = ((kBM + 8) * 128 + 1023) & ~1023;            // room for up to 128 + 8 strip rows, 1 KB aligned
  const long room = 227L * 1024 - 1024 - 2048;
  static const int mt_max = [] { const char* e = getenv("ES_TMA_STRIP_MT"); return e ? atoi(e) : 2; }();
  static const int bn_max = [] { const char* e = getenv("ES_TMA_STRIP_BN"); return e ? atoi(e) : 128; }();
  if (p.BN > bn_max) return false;    // N = 256 is mostly MMA-bound already; measured: strips lose there (3 stages, 10-20 % pad columns)
  const int mt = (p.BN <= 128 && mt_max >= 2) ? 2 : 1;
  int nx = 0;
  for (int d = L; d >= 2; --d)
    if (L % d == 0 && d <= 8 && 3L * (mt * a_bytes + d * BH * 128) <= room) { nx = d; break; }
  if (nx == 0 || n_rows * (L / nx) > 16) return false;
  const int Wp = p.Wo + nx - 1;
  const double gain = (double)nx * (kBM * 128 + BH * 128) / ((kBM + nx - 1) * 128.0 + nx * BH * 128.0) * p.Wo / Wp;
  if (gain < min_gain) return false;
  ts.nx = nx; ts.Wp = Wp; ts.Pp = p.Ho * Wp;
  if (Wp >= 256 || total_rows * ts.Pp >= 2147483647L) return false;
  for (int r = 0; r < n_rows; ++r)
    for (int j0 = 0; j0 < L; j0 += nx) {
      const int t = row_t[r] + j0;
      ts.sdy[ts.n_strips] = p.tdy[t]; ts.sdx[ts.n_strips] = p.tdx[t]; ts.skoff[ts.n_strips] = p.tkoff[t];
      ++ts.n_strips;
    }
  ts.a_bytes = a_bytes;
  ts.mt = mt;
  ts.stage_bytes = mt * a_bytes + nx * BH * 128;
  ts.stages = (int)(room / ts.stage_bytes);
  if (ts.stages > 6) ts.stages = 6;
  FwdParams q = p;                                                 // the im2col box enumerates Wp base pixels per row
  q.Wo = Wp;
  if (!tma_pair_plan(q, low, up)) return false;
  ts.low_w = low[0]; ts.low_h = low[1];
  return true;
}

static void conv_taps(const es_conv_geom* g, FwdParams& p) {
  p.Hs = g->Hs; p.Ws = g->Ws; p.C = g->C; p.Hu = g->Hu; p.Wu = g->Wu; p.Ho = g->Ho; p.Wo = g->Wo;
  p.n_taps = g->KH * g->KW; p.my = 1; p.mx = 1;
  for (int ky = 0; ky < g->KH; ++ky)
    for (int kx = 0; kx < g->KW; ++kx) {
      const int t = ky * g->KW + kx;
      p.tdy[t] = (signed char)(ky - g->pad); p.tdx[t] = (signed char)(kx - g->pad); p.tkoff[t] = t * g->C;
    }
  p.KK = g->KH * g->KW * g->C;
  p.Nout = g->N;
  p.o_my = 1; p.o_oy = 0; p.o_mx = 1; p.o_ox = 0; p.Wo_full = g->Wo; p.P_full = g->Ho * g->Wo;
}

static int launch_fwd(FwdParams& p, const void* x, const void* w, int total_rows, int n_groups, void* stream, int32_t* fused = nullptr) {
  if (fused) *fused = 0;     // set when the launched variant accumulated p.pair_sums (the TMA-fed pair kernels do)
  ES_REQUIRE(p.C > 0 && p.C % 64 == 0 && p.Hu <= 64 && p.Wu <= 64 && p.Hu >= p.Hs && p.Wu >= p.Ws && p.Ho > 0 && p.Wo > 0 &&
                 p.Ho < 256 && p.Wo < 256 && p.n_taps >= 1 && p.n_taps <= 32 && p.KK % 8 == 0,
             "unsupported geometry (need C % 64 == 0, Hu,Wu <= 64, <= 32 taps)");
  ES_REQUIRE(n_groups >= 1 && n_groups <= kFMaxGroups && total_rows > 0, "bad group count / rows");
  p.n_groups = n_groups;
  p.P = p.Ho * p.Wo;
  pick_bn(p);
  ES_REQUIRE(p.Nout % p.BN == 0, "N must be a multiple of 32");
  ES_REQUIRE((long)total_rows * p.Hs * p.Ws < 2147483647L && (long)total_rows * p.P_full < 2147483647L, "too many pixels");
  fill_maps_uc(p.Hs, p.Ws, p.Hu, p.Wu, p.ymap, p.xmap);
  p.a_src = (const __nv_bfloat16*)x;
  p.err_flag = fwd_err_flag();

  // weights as a 2-D tensor [slots*N rows][KK] bf16; the number of slots is not known here, so the row extent is the
  // largest one the group table can address (TMA never reads rows the kernel does not ask for)
  EncodeTiledFn enc = encode_fn();
  ES_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  alignas(64) CUtensorMap tmap, tmap_a;
  const cuuint32_t estr[2] = {1, 1};
  {
    const cuuint64_t dims[2] = {(cuuint64_t)p.KK, (cuuint64_t)kFMaxGroups * (cuuint64_t)p.Nout};
    const cuuint64_t strides[1] = {(cuuint64_t)p.KK * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)p.BN};
    const CUresult rc = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ES_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (weights must be 16-byte aligned, KK*2 a multiple of 16)");
  }
  {  // activations as [source pixels][C] for tile::gather4 (box = one row of 64 channels)
    const cuuint64_t dims[2] = {(cuuint64_t)p.C, (cuuint64_t)total_rows * p.Hs * p.Ws};
    const cuuint64_t strides[1] = {(cuuint64_t)p.C * 2};
    const cuuint32_t box[2] = {64, 1};
    const CUresult rc = enc(&tmap_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ES_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed for the activations");
  }

  // CTA-pair variant with the cp.async gather (see igemm_pair_kernel).  Measured on B200: on par with the single-CTA kernel
  // (N = 256: 0.870 vs 0.858 ms, N = 128: 0.572 vs 0.542 ms) — the gather, not the weight stream, is the bound — so it is
  // opt-in.  ES_IGEMM_PAIR: 0 (default) = off, 1 = every launch the strip / TMA-fed variants do not take, 2 = also instead of
  // the strip variant (A/B measurements).
  static const int pair_mode = [] { const char* e = getenv("ES_IGEMM_PAIR"); return e ? atoi(e) : 0; }();
  auto launch_pair = [&]() -> int {
    alignas(64) CUtensorMap tmap_h;
    const cuuint64_t dims[2] = {(cuuint64_t)p.KK, (cuuint64_t)kFMaxGroups * (cuuint64_t)p.Nout};
    const cuuint64_t strides[1] = {(cuuint64_t)p.KK * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)(p.BN / 2)};
    const CUresult rc = enc(&tmap_h, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ES_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (half weight tile)");
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long max_tiles = (ceil_div_l((long)total_rows * p.P, 2L * kBM) + n_groups) * p.n_tiles_n;
    const int pairs = (int)(max_tiles < sms / 2 ? max_tiles : sms / 2);
    constexpr int kPStages = 6;
    constexpr size_t kPSmem = (size_t)kPStages * (kFStageA + 128 * 128) + 1024 + 2048;
    static bool attr_set = false;
    if (!attr_set) {
      ES_CUDA(cudaFuncSetAttribute(igemm_pair_kernel<kPStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPSmem));
      attr_set = true;
    }
    igemm_pair_kernel<kPStages><<<2 * pairs, kGThreads, kPSmem, as_stream(stream)>>>(p, tmap_h);
    ES_LAUNCH_CHECK();
    return ES_OK;
  };
  const bool pair_ok = pair_mode > 0 && p.BN >= 64;

  // TMA-fed pair variant (see igemm_tma_pair_kernel): the conv reads the source directly (identity nearest maps).  Measured on
  // B200 against the cp.async-gather kernels (ms per launch, batch 1024, E = 8): conv1 forward classes 0.858/0.656/0.683/0.401
  // -> 0.593/0.464/0.487/0.288, conv2 data-gradient classes 0.582/0.486 -> 0.385/0.326, conv1 data gradient 2.262 -> 1.584,
  // conv3 data gradient (strip variant 0.910) -> 0.628, conv3 forward (strip 1.147) -> 1.054.  The dense fc2 product (one tap on
  // a 1x1 grid: pure weight streaming) is the one shape that lost (0.389 -> 0.440) and keeps the single-CTA kernel.
  // ES_IGEMM_TMA_A: 0 = off, 1 = wherever eligible except launches the strip variant takes, 2 (default) = also those.
  static const int tma_a_mode = [] { const char* e = getenv("ES_IGEMM_TMA_A"); return e ? atoi(e) : 2; }();
  static const bool trace = [] { const char* e = getenv("ES_IGEMM_TRACE"); return e && e[0] == '1'; }();   // variant per launch on stderr
  auto launch_tma_pair = [&](bool& taken) -> int {
    taken = false;
    TmaAParams ta{};
    int low[2], up[2];
    if (tma_a_mode <= 0 || !tma_pair_plan(p, low, up)) return ES_OK;
    EncodeIm2colFn enc_i = encode_im2col_fn();
    if (!enc_i) return ES_OK;
    ta.low_w = low[0]; ta.low_h = low[1];
    p.epi_staged = (p.n_taps == 1 && p.Hs * p.Ws == 1 && !p.pair_sums) ? 1 : 0;      // the dense product (fc2)
    alignas(64) CUtensorMap tmap_h, tmap_x;
    {
      const cuuint64_t dims[2] = {(cuuint64_t)p.KK, (cuuint64_t)kFMaxGroups * (cuuint64_t)p.Nout};
      const cuuint64_t strides[1] = {(cuuint64_t)p.KK * 2};
      const cuuint32_t box[2] = {64, (cuuint32_t)(p.BN / 2)};
      const CUresult rc = enc(&tmap_h, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      ES_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (half weight tile)");
    }
    {
      const cuuint64_t dims[4] = {(cuuint64_t)p.C, (cuuint64_t)p.Ws, (cuuint64_t)p.Hs, (cuuint64_t)total_rows};
      const cuuint64_t strides[3] = {(cuuint64_t)p.C * 2, (cuuint64_t)p.Ws * p.C * 2, (cuuint64_t)p.Hs * p.Ws * p.C * 2};
      const cuuint32_t trav[4] = {1, (cuuint32_t)p.mx, (cuuint32_t)p.my, 1};
      const CUresult rc = enc_i(&tmap_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, low, up, 64, kBM, trav,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (rc != CUDA_SUCCESS) return ES_OK;                     // geometry the driver refuses: fall back to the gather variants
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long max_tiles = (ceil_div_l((long)total_rows * p.P, 2L * kBM) + n_groups) * p.n_tiles_n;
    const int pairs = (int)(max_tiles < sms / 2 ? max_tiles : sms / 2);
    constexpr int kPStages = 6;
    constexpr size_t kPSmem = (size_t)kPStages * (kFStageA + 128 * 128) + 1024 + 2048 + 4096 + 4 * 32 * 80;   // + staged epilogue
    static bool attr_set = false;
    if (!attr_set) {
      ES_CUDA(cudaFuncSetAttribute(igemm_tma_pair_kernel<kPStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPSmem));
      attr_set = true;
    }
    igemm_tma_pair_kernel<kPStages><<<2 * pairs, 192, kPSmem, as_stream(stream)>>>(p, ta, tmap_h, tmap_x);
    ES_LAUNCH_CHECK();
    if (fused) *fused = p.pair_sums != nullptr;
    if (trace) fprintf(stderr, "[igemm] tma_pair C=%d N=%d taps=%d Ho=%d Wo=%d my=%d mx=%d rows=%d\n", p.C, p.Nout, p.n_taps, p.Ho, p.Wo,
                       p.my, p.mx, total_rows);
    taken = true;
    return ES_OK;
  };

  // Strip form of the TMA-fed pair variant (see igemm_tma_strip_kernel / tma_strip_plan): one im2col strip per tap row.
  auto launch_tma_strip = [&](bool& taken) -> int {
    taken = false;
    TStripParams ts{};
    int low[2], up[2];
    if (tma_a_mode <= 0 || !tma_strip_plan(p, total_rows, ts, low, up)) return ES_OK;
    EncodeIm2colFn enc_i = encode_im2col_fn();
    if (!enc_i) return ES_OK;
    alignas(64) CUtensorMap tmap_h, tmap_x;
    {
      const cuuint64_t dims[2] = {(cuuint64_t)p.KK, (cuuint64_t)kFMaxGroups * (cuuint64_t)p.Nout};
      const cuuint64_t strides[1] = {(cuuint64_t)p.KK * 2};
      const cuuint32_t box[2] = {64, (cuuint32_t)(p.BN / 2)};
      const CUresult rc = enc(&tmap_h, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      ES_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (half weight tile)");
    }
    {
      const cuuint64_t dims[4] = {(cuuint64_t)p.C, (cuuint64_t)p.Ws, (cuuint64_t)p.Hs, (cuuint64_t)total_rows};
      const cuuint64_t strides[3] = {(cuuint64_t)p.C * 2, (cuuint64_t)p.Ws * p.C * 2, (cuuint64_t)p.Hs * p.Ws * p.C * 2};
      const cuuint32_t trav[4] = {1, 1, (cuuint32_t)p.my, 1};
      const CUresult rc = enc_i(&tmap_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, low, up, 64,
                                (cuuint32_t)(kBM + ts.nx - 1), trav, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (rc != CUDA_SUCCESS) return ES_OK;                     // geometry the driver refuses: the per-tap variant takes it
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long max_tiles = (ceil_div_l((long)total_rows * ts.Pp, 2L * kBM * ts.mt) + n_groups) * p.n_tiles_n;
    const int pairs = (int)(max_tiles < sms / 2 ? max_tiles : sms / 2);
    const size_t smem = (size_t)ts.stages * ts.stage_bytes + 1024 + 2048;
    static bool attr_set = false;
    if (!attr_set) {
      ES_CUDA(cudaFuncSetAttribute(igemm_tma_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_set = true;
    }
    igemm_tma_strip_kernel<<<2 * pairs, kTSThreads, smem, as_stream(stream)>>>(p, ts, tmap_h, tmap_x);
    ES_LAUNCH_CHECK();
    if (fused) *fused = p.pair_sums != nullptr;
    if (trace) fprintf(stderr, "[igemm] tma_strip C=%d N=%d taps=%d mt=%d nx=%d strips=%d stages=%d Ho=%d Wo=%d my=%d rows=%d\n", p.C, p.Nout,
                       p.n_taps, ts.mt, ts.nx, ts.n_strips, ts.stages, p.Ho, p.Wo, p.my, total_rows);
    taken = true;
    return ES_OK;
  };
  if (!getenv("ES_IGEMM_FWD_VARIANT") && tma_a_mode >= 2) {
    bool taken = false;
    const int rc = launch_tma_strip(taken);
    if (rc != ES_OK || taken) return rc;
  }

  // Strip variant (see igemm_strip_kernel): no upsample, taps form rows of consecutive dx with consecutive weight columns,
  // N <= 128.  ES_IGEMM_STRIP=0 disables it (A/B measurements).
  {
    StripParams sp{};
    bool okk = strip_plan(p, total_rows, sp);
    if (okk && tma_a_mode >= 2) {
      bool taken = false;
      const int rc = launch_tma_pair(taken);
      if (rc != ES_OK || taken) return rc;
    }
    if (okk && pair_ok && pair_mode >= 2) return launch_pair();
    if (okk) {
      int dev = 0, sms = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      const long max_tiles = (ceil_div_l((long)total_rows * sp.Pp, (long)kBM) + n_groups) * p.n_tiles_n;
      const int grid = (int)(max_tiles < sms ? max_tiles : sms);
      static bool attr_set = false;
      if (!attr_set) {
        ES_CUDA(cudaFuncSetAttribute(igemm_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSSmem));
        attr_set = true;
      }
      igemm_strip_kernel<<<grid, kGThreads, kSSmem, as_stream(stream)>>>(p, sp, tmap);
      ES_LAUNCH_CHECK();
      return ES_OK;
    }
  }

  // Variant.  Measured on B200 (r01, per-launch CUDA events, batch 1024, E = 8; TFLOP/s for MT=1/cp.async, MT=2/cp.async,
  // MT=2/gather4): conv1 fwd 959 / 748 / 528, conv2 fwd 503 / 469 / 259, conv3 fwd 236 / 229 / 129, conv1 dgrad 927 / 731 /
  // 525.  The 128-row tile with double-buffered accumulators and the cp.async gather wins everywhere (the 3-stage pipeline
  // and the exposed epilogue of MT=2 cost more than the halved weight traffic saves; the TMA gather engine sustains about
  // half the row rate of cp.async), so it is the default; the other variants stay selectable for tuning.
  if (!getenv("ES_IGEMM_FWD_VARIANT")) {
    bool taken = false;
    const int rc = launch_tma_pair(taken);
    if (rc != ES_OK || taken) return rc;
  }
  if (pair_ok && !getenv("ES_IGEMM_FWD_VARIANT")) return launch_pair();
  int mt = 1, g4 = 0;
  if (const char* ov = getenv("ES_IGEMM_FWD_VARIANT")) {   // "mt,g4" — tuning aid
    int a = 0, b = 0;
    if (sscanf(ov, "%d,%d", &a, &b) == 2 && (a == 1 || a == 2) && (b == 0 || b == 1)) { mt = a; g4 = b; }
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long max_tiles = (ceil_div_l((long)total_rows * p.P, (long)mt * kBM) + n_groups) * p.n_tiles_n;
  const int grid = (int)(max_tiles < sms ? max_tiles : sms);
#define ES_FWD_LAUNCH(MTV, G4V)                                                                                          \
  {                                                                                                                      \
    static bool attr_set = false;                                                                                        \
    if (!attr_set) {                                                                                                     \
      ES_CUDA(cudaFuncSetAttribute(igemm_fwd_kernel<MTV, G4V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFSmem)); \
      attr_set = true;                                                                                                   \
    }                                                                                                                    \
    igemm_fwd_kernel<MTV, G4V><<<grid, kGThreads, kFSmem, as_stream(stream)>>>(p, tmap, tmap_a);                         \
  }
  if (mt == 1 && !g4) ES_FWD_LAUNCH(1, false)
  else if (mt == 1) ES_FWD_LAUNCH(1, true)
  else if (!g4) ES_FWD_LAUNCH(2, false)
  else ES_FWD_LAUNCH(2, true)
#undef ES_FWD_LAUNCH
  ES_LAUNCH_CHECK();
  return ES_OK;
}

static int igemm_fwd_impl(const void* x, const void* w, const float* bias, long bias_slot_stride, void* y, const es_conv_geom* g,
                          const es_group* grp, int n_groups, int total_rows, float* pair_sums, int32_t* fused, void* stream) {
  ES_REQUIRE(x && w && y && grp && g, "null pointer");
  ES_REQUIRE(g->KH >= 1 && g->KW >= 1 && g->KH * g->KW <= 32 && g->Ho == g->Hu + 2 * g->pad - g->KH + 1 &&
                 g->Wo == g->Wu + 2 * g->pad - g->KW + 1, "unsupported window (stride 1, <= 32 taps)");
  FwdParams p{};
  p.grp = grp;
  conv_taps(g, p);
  p.bias = bias; p.bias_slot_stride = bias_slot_stride; p.out = (__nv_bfloat16*)y;
  p.pair_sums = pair_sums;
  return launch_fwd(p, x, w, total_rows, n_groups, stream, fused);
}

extern "C" int es_igemm_fwd(const void* x, const void* w, const float* bias, long bias_slot_stride, void* y,
                            const es_conv_geom* g, const es_group* grp, int n_groups, int total_rows, void* stream) {
  return igemm_fwd_impl(x, w, bias, bias_slot_stride, y, g, grp, n_groups, total_rows, nullptr, nullptr, stream);
}

extern "C" int es_igemm_fwd_sums(const void* x, const void* w, const float* bias, long bias_slot_stride, void* y,
                                 const es_conv_geom* g, const es_group* grp, int n_groups, int total_rows, float* pair_sums,
                                 int32_t* fused, void* stream) {
  ES_REQUIRE(pair_sums && fused, "null pointer");
  return igemm_fwd_impl(x, w, bias, bias_slot_stride, y, g, grp, n_groups, total_rows, pair_sums, fused, stream);
}

extern "C" int es_igemm_fwd_plan(const es_conv_geom* g, int total_rows, int32_t* plan8) {
  ES_REQUIRE(g && plan8 && total_rows > 0, "null pointer");
  ES_REQUIRE(g->KH >= 1 && g->KW >= 1 && g->KH * g->KW <= 32 && g->Ho == g->Hu + 2 * g->pad - g->KH + 1 &&
                 g->Wo == g->Wu + 2 * g->pad - g->KW + 1 && g->N % 32 == 0, "unsupported window (stride 1, <= 32 taps, N % 32 == 0)");
  FwdParams p{};
  conv_taps(g, p);
  pick_bn(p);
  StripParams sp{};
  int low[2], up[2];
  TStripParams ts{};
  if (tma_strip_plan(p, total_rows, ts, low, up)) {   // variant 3: TMA-fed pair kernel with tap-row strips
    plan8[0] = 3; plan8[1] = p.BN; plan8[2] = ts.n_strips; plan8[3] = ts.nx; plan8[4] = ts.Wp; plan8[5] = ts.stages;
    plan8[6] = ts.n_strips * (p.C / kBK);
    plan8[7] = (int32_t)ceil_div_l((long)ts.Pp, kBM);
    return ES_OK;
  }
  if (tma_pair_plan(p, low, up)) {      // else the per-tap TMA-fed pair variant wherever it applies
    plan8[0] = 2; plan8[1] = p.BN; plan8[2] = p.n_taps; plan8[3] = 0; plan8[4] = p.Wo; plan8[5] = 6;
    plan8[6] = p.n_taps * (p.C / kBK);
    plan8[7] = (int32_t)ceil_div_l((long)p.Ho * p.Wo, kBM);
    return ES_OK;
  }
  const bool strip = strip_plan(p, total_rows, sp);
  plan8[0] = strip ? 1 : 0;
  plan8[1] = p.BN;
  plan8[2] = strip ? sp.n_strips : 0;
  plan8[3] = strip ? sp.nx : 0;
  plan8[4] = strip ? sp.Wp : p.Wo;
  plan8[5] = strip ? sp.stages : kFStages;
  plan8[6] = (strip ? sp.n_strips : p.n_taps) * (p.C / kBK);                    /* pipeline steps per tile */
  plan8[7] = (int32_t)ceil_div_l((long)(strip ? sp.Pp : p.Ho * p.Wo), kBM);     /* M tiles per row (per n tile) */
  return ES_OK;
}

static int igemm_taps_fwd_impl(const void* x, const void* w, const float* bias, long bias_slot_stride, void* y, const es_tap_geom* g,
                               const es_group* grp, int n_groups, int total_rows, float* pair_sums, int32_t* fused, void* stream) {
  ES_REQUIRE(x && w && y && grp && g, "null pointer");
  ES_REQUIRE(g->n_taps >= 1 && g->n_taps <= 32 && g->my >= 1 && g->mx >= 1 && g->o_my >= 1 && g->o_mx >= 1, "bad tap table");
  ES_REQUIRE((g->Ho - 1) * g->o_my + g->o_oy < g->Ho_full && (g->Wo - 1) * g->o_mx + g->o_ox < g->Wo_full, "output scatter out of range");
  FwdParams p{};
  p.grp = grp;
  p.Hs = g->Hs; p.Ws = g->Ws; p.C = g->C; p.Hu = g->Hu; p.Wu = g->Wu; p.Ho = g->Ho; p.Wo = g->Wo;
  p.n_taps = g->n_taps; p.my = g->my; p.mx = g->mx;
  for (int t = 0; t < g->n_taps; ++t) {
    p.tdy[t] = g->tap_dy[t]; p.tdx[t] = g->tap_dx[t]; p.tkoff[t] = g->tap_koff[t];
    ES_REQUIRE(g->tap_koff[t] >= 0 && g->tap_koff[t] % 64 == 0 && g->tap_koff[t] + g->C <= g->KK, "tap weight offset out of range");
  }
  p.KK = g->KK;
  p.Nout = g->N;
  p.o_my = g->o_my; p.o_oy = g->o_oy; p.o_mx = g->o_mx; p.o_ox = g->o_ox; p.Wo_full = g->Wo_full; p.P_full = g->Ho_full * g->Wo_full;
  p.bias = bias; p.bias_slot_stride = bias_slot_stride; p.out = (__nv_bfloat16*)y;
  p.pair_sums = pair_sums;
  return launch_fwd(p, x, w, total_rows, n_groups, stream, fused);
}

extern "C" int es_igemm_taps_fwd(const void* x, const void* w, const float* bias, long bias_slot_stride, void* y,
                                 const es_tap_geom* g, const es_group* grp, int n_groups, int total_rows, void* stream) {
  return igemm_taps_fwd_impl(x, w, bias, bias_slot_stride, y, g, grp, n_groups, total_rows, nullptr, nullptr, stream);
}

extern "C" int es_igemm_taps_fwd_sums(const void* x, const void* w, const float* bias, long bias_slot_stride, void* y,
                                      const es_tap_geom* g, const es_group* grp, int n_groups, int total_rows, float* pair_sums,
                                      int32_t* fused, void* stream) {
  ES_REQUIRE(pair_sums && fused, "null pointer");
  return igemm_taps_fwd_impl(x, w, bias, bias_slot_stride, y, g, grp, n_groups, total_rows, pair_sums, fused, stream);
}

// Host-side plan of the strip form of the TMA-fed weight gradient (igemm_wgrad_strip_kernel): equal tap rows of consecutive dx
// on a source read directly (mx = 1), N <= 128 (where the per-tap kernel is fill-bound), whole 128-channel chunks per CTA.
static bool wgrad_strip_plan(const WgParams& p, long total_rows, bool as_pair, WgStripParams& ws, int low[2], int up[2]) {
  static const int on = [] { const char* e = getenv("ES_WG_STRIP"); return e ? atoi(e) : 1; }();
  ws = WgStripParams{};
  static const int n_max = [] { const char* e = getenv("ES_WG_STRIP_N"); return e ? atoi(e) : 128; }();
  if (!on || p.N > n_max || p.mx != 1 || p.my < 1 || p.my > 8 || p.Hu != p.Hs || p.Wu != p.Ws || p.n_taps < 2) return false;
  if (p.C % (as_pair ? 2 * kBM : kBM) != 0) return false;
  int L = 0, n_rows = 0, row_t[32];
  for (int t = 0; t < p.n_taps;) {
    int n = 1;
    while (t + n < p.n_taps && p.tdy[t + n] == p.tdy[t] && p.tdx[t + n] == p.tdx[t] + n) ++n;
    if (n_rows == 0) L = n;
    if (n != L) return false;
    row_t[n_rows++] = t;
    t += n;
  }
  int nx = 0;
  for (int d = L; d >= 2; --d)
    if (L % d == 0 && d <= 8 && d * p.N <= 512) { nx = d; break; }
  if (nx == 0 || n_rows * (L / nx) > 16) return false;
  ws.nx = nx; ws.Wp = p.Wo + nx - 1; ws.Pp = p.Ho * ws.Wp;
  if (ws.Wp >= 256 || total_rows * ws.Pp >= 2147483647L) return false;
  for (int r = 0; r < n_rows; ++r)
    for (int j0 = 0; j0 < L; j0 += nx) {
      const int t = row_t[r] + j0;
      ws.sdy[ws.n_rows] = p.tdy[t]; ws.sdx[ws.n_rows] = p.tdx[t]; ws.tap0[ws.n_rows] = t;
      ++ws.n_rows;
    }
  ws.cchunks = p.C / (as_pair ? 2 * kBM : kBM);
  ws.seg_bytes = ((kBK + nx - 1) * 128 + 1023) & ~1023;
  ws.stage_bytes = 2 * ws.seg_bytes + (as_pair ? p.N / 2 : p.N) * 128;
  ws.stages = (int)((227L * 1024 - 1024 - 2048) / ws.stage_bytes);
  if (ws.stages > 8) ws.stages = 8;
  if (ws.stages < 3) return false;
  FwdParams q{};                                                   // the x box enumerates Wp base pixels per row
  q.Hs = p.Hs; q.Ws = p.Ws; q.C = p.C; q.Hu = p.Hu; q.Wu = p.Wu; q.Ho = p.Ho; q.Wo = ws.Wp;
  q.n_taps = p.n_taps; q.my = p.my; q.mx = 1; q.BN = 64;
  for (int t = 0; t < p.n_taps; ++t) { q.tdy[t] = p.tdy[t]; q.tdx[t] = p.tdx[t]; }
  if (!tma_pair_plan(q, low, up)) return false;
  ws.low_w = low[0]; ws.low_h = low[1];
  return true;
}

static int launch_wgrad(WgParams& p, const void* x, const void* dy, float* dw, long dw_slot_stride, int total_rows,
                        int n_groups, void* stream) {
  ES_REQUIRE(p.C > 0 && p.C % 64 == 0 && p.Hu <= 64 && p.Wu <= 64 && p.Hu >= p.Hs && p.Wu >= p.Ws && p.Ho > 0 && p.Wo > 0 &&
                 p.n_taps >= 1 && p.n_taps <= 32, "unsupported geometry");
  ES_REQUIRE(n_groups >= 1 && n_groups <= kFMaxGroups && total_rows > 0, "bad group count / rows");
  ES_REQUIRE(p.N % 64 == 0 && p.N <= 256 && (p.N & (p.N - 1)) == 0, "dy channels must be 64, 128 or 256");
  p.n_groups = n_groups;
  p.P = p.Ho * p.Wo;
  p.KK = p.n_taps * p.C;
  ES_REQUIRE(p.KK % kBM == 0, "taps*C must be a multiple of 128");
  ES_REQUIRE((long)total_rows * p.P < 2147483647L && (long)total_rows * p.Hs * p.Ws * p.C < (1L << 40), "too many pixels");
  p.tiles_m = p.KK / kBM;
  int splits = ceil_div(1024, p.tiles_m * n_groups);
  const long kblocks = ceil_div_l((long)total_rows * p.P, kBK);
  if (splits > kblocks) splits = (int)kblocks;
  if (splits < 1) splits = 1;
  if (splits > 64) splits = 64;
  p.splits = splits;
  fill_maps_uc(p.Hs, p.Ws, p.Hu, p.Wu, p.ymap, p.xmap);
  p.x = (const __nv_bfloat16*)x; p.dw = dw; p.dw_slot_stride = dw_slot_stride; p.err_flag = fwd_err_flag();

  EncodeTiledFn enc = encode_fn();
  ES_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  alignas(64) CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)p.N, (cuuint64_t)total_rows * (cuuint64_t)p.P};
  const cuuint64_t strides[1] = {(cuuint64_t)p.N * 2};
  const cuuint32_t box[2] = {64, 64};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult rc = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(dy), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ES_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed for dy");
  static bool attr_set = false;
  if (!attr_set) {
    ES_CUDA(cudaFuncSetAttribute(igemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFSmem));
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

  // TMA-fed variant (igemm_wgrad_tma_kernel): x is read directly (identity nearest maps).  As a CTA pair when dy has 128 or
  // 256 channels and KK is a multiple of 256, as a single CTA for 64 channels.  It reduces over the full 64-pixel blocks of
  // every group; the partial last block of each group goes through the gather kernel below (kmode 2).
  static const int tma_a_mode = [] { const char* e = getenv("ES_IGEMM_TMA_A"); return e ? atoi(e) : 2; }();
  static const bool trace = [] { const char* e = getenv("ES_IGEMM_TRACE"); return e && e[0] == '1'; }();   // variant per launch on stderr
  {
    FwdParams fp{};
    fp.Hs = p.Hs; fp.Ws = p.Ws; fp.C = p.C; fp.Hu = p.Hu; fp.Wu = p.Wu; fp.Ho = p.Ho; fp.Wo = p.Wo;
    fp.n_taps = p.n_taps; fp.my = p.my; fp.mx = p.mx; fp.BN = p.N;
    for (int t = 0; t < p.n_taps; ++t) { fp.tdy[t] = p.tdy[t]; fp.tdx[t] = p.tdx[t]; }
    int low[2], up[2];
    EncodeIm2colFn enc_i = encode_im2col_fn();
    const bool as_pair = (p.N == 128 || p.N == 256) && p.KK % (2 * kBM) == 0;
    // strip form (igemm_wgrad_strip_kernel): one x strip per tap row, the taps of the row share it and the dy tile
    WgStripParams ws{};
    if (tma_a_mode > 0 && enc_i && (as_pair || p.N == 64) && wgrad_strip_plan(p, total_rows, as_pair, ws, low, up)) {
      alignas(64) CUtensorMap tmap_x, tmap_dyi;
      const cuuint64_t xdims[4] = {(cuuint64_t)p.C, (cuuint64_t)p.Ws, (cuuint64_t)p.Hs, (cuuint64_t)total_rows};
      const cuuint64_t xstrides[3] = {(cuuint64_t)p.C * 2, (cuuint64_t)p.Ws * p.C * 2, (cuuint64_t)p.Hs * p.Ws * p.C * 2};
      const cuuint32_t trav[4] = {1, 1, (cuuint32_t)p.my, 1};
      CUresult rcx = enc_i(&tmap_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), xdims, xstrides, low, up, 64,
                           (cuuint32_t)(kBK + ws.nx - 1), trav, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (rcx == CUDA_SUCCESS) {     // dy on the padded pitch: the box reaches nx - 1 columns past the image (zero-filled)
        const cuuint64_t ddims[4] = {(cuuint64_t)p.N, (cuuint64_t)p.Wo, (cuuint64_t)p.Ho, (cuuint64_t)total_rows};
        const cuuint64_t dstrides[3] = {(cuuint64_t)p.N * 2, (cuuint64_t)p.Wo * p.N * 2, (cuuint64_t)p.Ho * p.Wo * p.N * 2};
        const cuuint32_t trav1[4] = {1, 1, 1, 1};
        const int dlow[2] = {0, 0}, dup[2] = {ws.nx - 1, 0};
        rcx = enc_i(&tmap_dyi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dy), ddims, dstrides, dlow, dup, 64, kBK, trav1,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      }
      if (rcx == CUDA_SUCCESS) {
        WgParams pm = p;
        const int units_m = ws.n_rows * ws.cchunks, workers = as_pair ? sms / 2 : sms;
        int sp2 = (2 * workers + n_groups * units_m / 2) / (n_groups * units_m);
        const long full_blocks = ((long)total_rows * ws.Pp) / kBK / n_groups;      // per group, roughly
        if (sp2 > full_blocks) sp2 = (int)full_blocks;
        if (sp2 < 1) sp2 = 1;
        if (sp2 > 64) sp2 = 64;
        pm.splits = sp2;
        const size_t smem = (size_t)ws.stages * ws.stage_bytes + 1024 + 2048;
        static bool attr3 = false;
        if (!attr3) {
          ES_CUDA(cudaFuncSetAttribute(igemm_wgrad_strip_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
          ES_CUDA(cudaFuncSetAttribute(igemm_wgrad_strip_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
          attr3 = true;
        }
        const int units = n_groups * pm.splits * units_m;
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(192);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = as_stream(stream);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = as_pair ? 2 : 1;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        if (as_pair) {
          cfg.gridDim = dim3(2 * (units < workers ? units : workers));
          ES_CUDA(cudaLaunchKernelEx(&cfg, igemm_wgrad_strip_kernel<true>, pm, ws, tmap_dyi, tmap_x));
        } else {
          cfg.gridDim = dim3(units < workers ? units : workers);
          ES_CUDA(cudaLaunchKernelEx(&cfg, igemm_wgrad_strip_kernel<false>, pm, ws, tmap_dyi, tmap_x));
        }
        ES_LAUNCH_CHECK();
        if (trace) fprintf(stderr, "[igemm] wgrad_strip C=%d N=%d taps=%d nx=%d rows=%d pair=%d stages=%d splits=%d Ho=%d Wo=%d my=%d\n", p.C, p.N,
                           p.n_taps, ws.nx, ws.n_rows, (int)as_pair, ws.stages, pm.splits, p.Ho, p.Wo, p.my);
        WgParams pt = p;                  // what the padded full blocks did not cover: gather kernel, masked below pix_lo
        pt.kmode = 3; pt.splits = 1; pt.s_Wp = ws.Wp; pt.s_Pp = ws.Pp;
        const int tail_units = n_groups * pt.tiles_m;
        igemm_wgrad_kernel<<<tail_units < sms ? tail_units : sms, kGThreads, kFSmem, as_stream(stream)>>>(pt, tmap);
        ES_LAUNCH_CHECK();
        return ES_OK;
      }
    }
    if (tma_a_mode > 0 && enc_i && (as_pair || p.N == 64) && tma_pair_plan(fp, low, up)) {
      alignas(64) CUtensorMap tmap_x;
      const cuuint64_t xdims[4] = {(cuuint64_t)p.C, (cuuint64_t)p.Ws, (cuuint64_t)p.Hs, (cuuint64_t)total_rows};
      const cuuint64_t xstrides[3] = {(cuuint64_t)p.C * 2, (cuuint64_t)p.Ws * p.C * 2, (cuuint64_t)p.Hs * p.Ws * p.C * 2};
      const cuuint32_t trav[4] = {1, (cuuint32_t)p.mx, (cuuint32_t)p.my, 1};
      const CUresult rcx = enc_i(&tmap_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), xdims, xstrides, low, up, 64, kBK, trav,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (rcx == CUDA_SUCCESS) {
        TmaAParams ta{};
        ta.low_w = low[0]; ta.low_h = low[1];
        WgParams pm = p;
        pm.kmode = 1;
        pm.tile_m = as_pair ? 2 * kBM : kBM;
        pm.tiles_m = p.KK / pm.tile_m;
        int sp2 = ceil_div(as_pair ? 512 : 1024, pm.tiles_m * n_groups);
        const long full_blocks = ((long)total_rows * p.P) / kBK;
        if (sp2 > full_blocks) sp2 = (int)full_blocks;
        if (sp2 < 1) sp2 = 1;
        if (sp2 > 64) sp2 = 64;
        pm.splits = sp2;
        constexpr int kPStages = 6;
        constexpr size_t kPSmem = (size_t)kPStages * (kFStageA + 128 * 128) + 1024 + 2048;
        static bool attr2 = false;
        if (!attr2) {
          ES_CUDA(cudaFuncSetAttribute(igemm_wgrad_tma_kernel<kPStages, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPSmem));
          ES_CUDA(cudaFuncSetAttribute(igemm_wgrad_tma_kernel<kPStages, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPSmem));
          attr2 = true;
        }
        const int units = n_groups * pm.splits * pm.tiles_m;
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(192);
        cfg.dynamicSmemBytes = kPSmem;
        cfg.stream = as_stream(stream);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = as_pair ? 2 : 1;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        if (as_pair) {
          const int pairs = units < sms / 2 ? units : sms / 2;
          cfg.gridDim = dim3(2 * pairs);
          ES_CUDA(cudaLaunchKernelEx(&cfg, igemm_wgrad_tma_kernel<kPStages, true>, pm, ta, tmap, tmap_x));
        } else {
          cfg.gridDim = dim3(units < sms ? units : sms);
          ES_CUDA(cudaLaunchKernelEx(&cfg, igemm_wgrad_tma_kernel<kPStages, false>, pm, ta, tmap, tmap_x));
        }
        ES_LAUNCH_CHECK();
        WgParams pt = p;                  // partial last block of every group: gather kernel, one k-block per unit
        pt.kmode = 2; pt.splits = 1;
        const int tail_units = n_groups * pt.tiles_m;
        igemm_wgrad_kernel<<<tail_units < sms ? tail_units : sms, kGThreads, kFSmem, as_stream(stream)>>>(pt, tmap);
        ES_LAUNCH_CHECK();
        return ES_OK;
      }
    }
  }
  const int total = n_groups * p.splits * p.tiles_m;
  igemm_wgrad_kernel<<<total < sms ? total : sms, kGThreads, kFSmem, as_stream(stream)>>>(p, tmap);
  ES_LAUNCH_CHECK();
  return ES_OK;
}

extern "C" int es_igemm_wgrad(const void* x, const void* dy, float* dw, const es_conv_geom* g, const es_group* grp,
                              int n_groups, int total_rows, void* stream) {
  ES_REQUIRE(x && dy && dw && grp && g, "null pointer");
  ES_REQUIRE(g->KH >= 1 && g->KW >= 1 && g->KH * g->KW <= 32 && g->Ho == g->Hu + 2 * g->pad - g->KH + 1 &&
                 g->Wo == g->Wu + 2 * g->pad - g->KW + 1, "unsupported window (stride 1, <= 32 taps)");
  WgParams p{};
  p.grp = grp;
  p.Hs = g->Hs; p.Ws = g->Ws; p.C = g->C; p.Hu = g->Hu; p.Wu = g->Wu; p.Ho = g->Ho; p.Wo = g->Wo;
  p.n_taps = g->KH * g->KW; p.my = 1; p.mx = 1;
  for (int ky = 0; ky < g->KH; ++ky)
    for (int kx = 0; kx < g->KW; ++kx) {
      const int t = ky * g->KW + kx;
      p.tdy[t] = (signed char)(ky - g->pad); p.tdx[t] = (signed char)(kx - g->pad); p.tcol[t] = t * g->C;
    }
  p.N = g->N;
  p.dw_ld = g->KH * g->KW * g->C;
  return launch_wgrad(p, x, dy, dw, (long)g->N * p.dw_ld, total_rows, n_groups, stream);
}

/* tap-table weight gradient: dy is [rows, Ho*Wo, N] (the pixels of ONE output phase, contiguous), x the low-resolution source;
 * dw[slot][n][tap_koff[t] + c] += sum_pix dy[pix, n] * x[(oy + tap_dy[t], ox + tap_dx[t]), c];  a dw row has g->KK columns */
extern "C" int es_igemm_taps_wgrad(const void* x, const void* dy, float* dw, const es_tap_geom* g, const es_group* grp,
                                   int n_groups, int total_rows, void* stream) {
  ES_REQUIRE(x && dy && dw && grp && g, "null pointer");
  ES_REQUIRE(g->my >= 1 && g->mx >= 1, "bad pixel stride");
  WgParams p{};
  p.grp = grp;
  p.Hs = g->Hs; p.Ws = g->Ws; p.C = g->C; p.Hu = g->Hu; p.Wu = g->Wu; p.Ho = g->Ho; p.Wo = g->Wo;
  p.n_taps = g->n_taps; p.my = g->my; p.mx = g->mx;
  ES_REQUIRE(g->n_taps >= 1 && g->n_taps <= 32, "bad tap table");
  for (int t = 0; t < g->n_taps; ++t) {
    p.tdy[t] = g->tap_dy[t]; p.tdx[t] = g->tap_dx[t]; p.tcol[t] = g->tap_koff[t];
    ES_REQUIRE(g->tap_koff[t] >= 0 && g->tap_koff[t] + g->C <= g->KK, "tap column offset out of range");
  }
  p.N = g->N;
  p.dw_ld = g->KK;
  return launch_wgrad(p, x, dy, dw, (long)g->N * g->KK, total_rows, n_groups, stream);
}
