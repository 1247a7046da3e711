// Probe: semantics of TMA tile::gather4 on sm_100a (box shape, OOB rows, 128B swizzle layout).  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/g4 tools/experiments/gather4_probe.cu && /tmp/g4
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void k(const __grid_constant__ CUtensorMap tm, int4 rows, int col, unsigned short* out, int* status) {
  __shared__ __align__(1024) uint8_t buf[1024];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(buf), b = (uint32_t)__cvta_generic_to_shared(&bar);
  for (int i = threadIdx.x; i < 512; i += blockDim.x) ((unsigned short*)buf)[i] = 0xDEAD;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(b), "r"(512) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(&tm), "r"(col), "r"(rows.x), "r"(rows.y), "r"(rows.z), "r"(rows.w), "r"(b) : "memory");
    uint32_t ok = 0;
    long long t0 = clock64();
    while (!ok && clock64() - t0 < 200000000LL) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(0) : "memory");
    }
    *status = ok;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = ((unsigned short*)buf)[i];
}

int main() {
  const int R = 64, C = 128;
  std::vector<unsigned short> h(R * C);
  for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[r * C + c] = (unsigned short)(r * 256 + c);   // raw 16-bit patterns
  unsigned short *d, *o; int* st;
  cudaMalloc(&d, R * C * 2); cudaMalloc(&o, 1024); cudaMalloc(&st, 4);
  cudaMemcpy(d, h.data(), R * C * 2, cudaMemcpyHostToDevice);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  for (int boxh : {1, 4}) {
    for (int sw = 0; sw < 2; ++sw) {
      CUtensorMap tm;
      cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R}; cuuint64_t strides[1] = {(cuuint64_t)C * 2};
      cuuint32_t box[2] = {64, (cuuint32_t)boxh}; cuuint32_t es[2] = {1, 1};
      CUresult rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("box {64,%d} swizzle %d: encode rc=%d\n", boxh, sw, (int)rc);
      if (rc) continue;
      cudaMemset(st, 0, 4);
      k<<<1, 128>>>(tm, make_int4(5, 40, -1, 17), 64, o, st);
      cudaError_t e = cudaDeviceSynchronize();
      int s = -1; cudaMemcpy(&s, st, 4, cudaMemcpyDeviceToHost);
      std::vector<unsigned short> r(512);
      cudaMemcpy(r.data(), o, 1024, cudaMemcpyDeviceToHost);
      printf("  kernel: %s, barrier completed=%d\n", cudaGetErrorString(e), s);
      if (e != cudaSuccess) return 1;
      for (int row = 0; row < 4; ++row) {
        printf("  smem row %d:", row);
        for (int ch = 0; ch < 8; ++ch) printf(" [%04x..%04x]", r[row * 64 + ch * 8], r[row * 64 + ch * 8 + 7]);
        printf("\n");
      }
    }
  }
  return 0;
}
