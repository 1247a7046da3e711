// Probe: semantics of TMA im2col mode on sm_100a (cuTensorMapEncodeIm2col + cp.async.bulk.tensor.4d...im2col): which source
// pixel lands in which shared-memory row for a given start coordinate, bounding box, traversal stride and tap offset, how the
// traversal crosses rows / samples, what is zero-filled.  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/i2c tools/experiments/im2col_probe.cu && /tmp/i2c
// Expected model (checked below): a tensor [N][H][W][C]; base pixels are enumerated w fastest, then h, then n over the box
//   w in {lowW + i*sW : i in [0, Q)},  Q = (W + upW - lowW - 1) / sW + 1   (same for h), starting at the instruction's
//   coordinate {c, w, h, n}; smem row r holds channels [c, c+64) of the r-th base pixel shifted by the instruction's
//   offsets {offW, offH}; pixels outside [0,W)x[0,H) are zero; 128B swizzle = chunk ^ (row & 7).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int PIX = 128;

__global__ void k(const __grid_constant__ CUtensorMap tm, int c, int w, int h, int n, int offw, int offh, unsigned short* out,
                  int* status) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* buf = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(buf), b = (uint32_t)__cvta_generic_to_shared(&bar);
  for (int i = threadIdx.x; i < PIX * 64; i += blockDim.x) ((unsigned short*)buf)[i] = 0xDEAD;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(b), "r"(PIX * 128) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6], {%7, %8};"
                 ::"r"(dst), "l"(&tm), "r"(c), "r"(w), "r"(h), "r"(n), "r"(b), "h"((unsigned short)offw), "h"((unsigned short)offh) : "memory");
    uint32_t ok = 0;
    long long t0 = clock64();
    while (!ok && clock64() - t0 < 400000000LL) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(0) : "memory");
    }
    *status = ok;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PIX * 64; i += blockDim.x) out[i] = ((unsigned short*)buf)[i];
}

struct Case { const char* name; int N, H, W, C; int lowW, lowH, upW, upH, sW, sH; int c, w0, h0, n0, offW, offH; };

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fp, cudaEnableDefault, &q);
  EncodeIm2colFn enc = (EncodeIm2colFn)fp;
  if (!enc) { printf("no cuTensorMapEncodeIm2col\n"); return 1; }
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, PIX * 128 + 1024);
  const Case cases[] = {
      // 3x3 pad 1 on an 18x10 grid (conv1 phase (0,0)): tap (dy,dx) = (-1,-1) -> offsets (0,0); start at output pixel (oy 3, ox 4) of sample 2
      {"3x3 pad1 tap(-1,-1)", 6, 18, 10, 128, -1, -1, -1, -1, 1, 1, 64, 4 - 1, 3 - 1, 2, 0, 0},
      {"3x3 pad1 tap(+1,0) ", 6, 18, 10, 128, -1, -1, -1, -1, 1, 1, 0, 4 - 1, 3 - 1, 2, 1, 2},
      // taps dx in {0,1}, dy in {0,1}, 17x10 outputs from an 18x10 source: lower 0, upper chosen so that Q = outputs
      {"2x2 lower0 Q=10x17 ", 6, 18, 10, 128, 0, 0, 0, -1, 1, 1, 0, 7, 16, 1, 1, 1},
      // traversal stride 2 (data gradient of the x2-folded conv: dy pixel = 2*s + d), source 36x20, outputs 18x10, d in [-2, 1]
      {"stride2 d in[-2,1]   ", 4, 36, 20, 64, -2, -2, (10 - 1) * 2 + 1 - 2 - 20, (18 - 1) * 2 + 1 - 2 - 36, 2, 2, 0, -2 + 2 * 3, -2 + 2 * 5, 1, 3, 0},
      // last sample: the 128-pixel run leaves the tensor -> zeros
      {"runs past the tensor ", 3, 18, 10, 64, -1, -1, -1, -1, 1, 1, 0, -1, 10 - 1, 2, 1, 1},
  };
  int bad_total = 0;
  for (const Case& cs : cases) {
    const size_t n_el = (size_t)cs.N * cs.H * cs.W * cs.C;
    std::vector<unsigned short> hsrc(n_el);
    // 16-bit pattern: n(3 bits) h(5) w(5) c/8(3) -> unique per (pixel, 8-channel chunk); never 0 and never 0xDEAD
    for (int n = 0; n < cs.N; ++n) for (int h = 0; h < cs.H; ++h) for (int w = 0; w < cs.W; ++w) for (int c = 0; c < cs.C; ++c)
      hsrc[(((size_t)n * cs.H + h) * cs.W + w) * cs.C + c] = (unsigned short)(0x8000 | (n << 12) | (h << 7) | (w << 2) | ((c / 16) & 3));
    unsigned short *d, *o; int* st;
    size_t bytes = n_el * 2 < 262144 ? 262144 : n_el * 2;      // keep the allocation >= 128 KB (driver quirk for tiny tensors, see CUTLASS)
    cudaMalloc(&d, bytes); cudaMalloc(&o, PIX * 128); cudaMalloc(&st, 4);
    cudaMemset(d, 0, bytes);
    cudaMemcpy(d, hsrc.data(), n_el * 2, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)cs.C, (cuuint64_t)cs.W, (cuuint64_t)cs.H, (cuuint64_t)cs.N};
    cuuint64_t strides[3] = {(cuuint64_t)cs.C * 2, (cuuint64_t)cs.W * cs.C * 2, (cuuint64_t)cs.H * cs.W * cs.C * 2};
    int lower[2] = {cs.lowW, cs.lowH}, upper[2] = {cs.upW, cs.upH};
    cuuint32_t es[4] = {1, (cuuint32_t)cs.sW, (cuuint32_t)cs.sH, 1};
    CUresult rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, strides, lower, upper, 64, PIX, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("%s: encode rc=%d (lower %d,%d upper %d,%d stride %d,%d)\n", cs.name, (int)rc, cs.lowW, cs.lowH, cs.upW, cs.upH, cs.sW, cs.sH);
    if (rc) { ++bad_total; continue; }
    cudaMemset(st, 0, 4);
    k<<<1, 256, PIX * 128 + 1024>>>(tm, cs.c, cs.w0, cs.h0, cs.n0, cs.offW, cs.offH, o, st);
    cudaError_t e = cudaDeviceSynchronize();
    int s = -1; cudaMemcpy(&s, st, 4, cudaMemcpyDeviceToHost);
    printf("  kernel: %s, barrier completed=%d\n", cudaGetErrorString(e), s);
    if (e != cudaSuccess) return 1;
    std::vector<unsigned short> r(PIX * 64);
    cudaMemcpy(r.data(), o, PIX * 128, cudaMemcpyDeviceToHost);
    // expectation
    const int Qw = (cs.W + cs.upW - cs.lowW - 1) / cs.sW + 1, Qh = (cs.H + cs.upH - cs.lowH - 1) / cs.sH + 1;
    int iw = (cs.w0 - cs.lowW) / cs.sW, ih = (cs.h0 - cs.lowH) / cs.sH, in = cs.n0, bad = 0;
    printf("  model: Q = %d x %d, start index (w %d, h %d, n %d)\n", Qw, Qh, iw, ih, in);
    for (int row = 0; row < PIX; ++row) {
      const int w = cs.lowW + iw * cs.sW + cs.offW, h = cs.lowH + ih * cs.sH + cs.offH;
      const bool inb = in < cs.N && w >= 0 && w < cs.W && h >= 0 && h < cs.H;
      for (int ch = 0; ch < 8; ++ch) {
        const unsigned short want = inb ? (unsigned short)(0x8000 | (in << 12) | (h << 7) | (w << 2) | (((cs.c + ch * 8) / 16) & 3)) : 0;
        const unsigned short got = r[row * 64 + ((ch ^ (row & 7)) * 8)];
        if (got != want && bad < 6) {
          printf("  MISMATCH row %d chunk %d: got %04x (n%d h%d w%d c16=%d) want %04x (n%d h%d w%d)%s\n", row, ch, got, (got >> 12) & 7, (got >> 7) & 31,
                 (got >> 2) & 31, got & 3, want, in, h, w, inb ? "" : " [zero]");
        }
        bad += got != want;
      }
      if (++iw == Qw) { iw = 0; if (++ih == Qh) { ih = 0; ++in; } }
    }
    printf("  %s (%d mismatching chunks of %d)\n", bad ? "MODEL WRONG" : "model confirmed", bad, PIX * 8);
    bad_total += bad != 0;
    if (bad) {
      for (int row : {0, 1, 9, 10, 11}) {
        printf("  smem row %3d:", row);
        for (int ch = 0; ch < 8; ++ch) { unsigned short g = r[row * 64 + ch * 8]; printf(" %04x(n%d h%d w%d)", g, (g >> 12) & 7, (g >> 7) & 31, (g >> 2) & 31); }
        printf("\n");
      }
    }
    cudaFree(d); cudaFree(o); cudaFree(st);
  }
  printf("%d case(s) off the model\n", bad_total);
  return 0;
}
