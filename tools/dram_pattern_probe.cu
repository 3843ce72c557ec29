// tools/dram_pattern_probe.cu -- does the access pattern of the GEMM2 epilogue (128-byte pieces of rows 12 KB apart, each
// 1 KB row segment visited by 8 different warp-instructions at different times) reach the HBM bandwidth a contiguous
// stream reaches?  Reads (and optionally writes) a (rows, L*d) fp32 state like k2_chunk does, vs the same bytes laid out
// tile-contiguously.  One CTA of 512 threads per (128-row block, level, 256-column tile), 16 warps = 4 row bands x 4 parts.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/dram_pattern_probe tools/dram_pattern_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

template <int BLOCKED, int WRITE>
__global__ void __launch_bounds__(512) k(const float* __restrict__ src, float* __restrict__ dst, int rows, int L, int d, int iters) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = warp & 3, part = warp >> 2;
  const int c = lane & 7, rsub = lane >> 3;
  const int tiles_n = d / 256, tiles_m = rows / 128;
  const int ntiles = tiles_m * L * tiles_n;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it)
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int nb = tile % tiles_n, l = (tile / tiles_n) % L, mb = tile / (tiles_n * L);
    for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = quad * 32 + (h * 4 + j) * 4 + rsub;
          const int col = nb * 256 + part * 64 + ch * 32 + c * 4;
          size_t off;
          if (BLOCKED) off = ((size_t)tile * 128 + r) * 256 + (part * 64 + ch * 32 + c * 4);   // tile-contiguous: 128 x 256 block
          else off = ((size_t)(mb * 128 + r) * L + l) * d + col;
          v[j] = __ldcs(reinterpret_cast<const float4*>(src + off));
          if (WRITE) __stcs(reinterpret_cast<float4*>(dst + off), make_float4(v[j].x * 0.25f, v[j].y, v[j].z, v[j].w));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc += v[j].x + v[j].w;
      }
    }
  }
  if (acc == 123.456f) dst[0] = acc;
}

template <int BLOCKED, int WRITE>
static void run(const char* name, const float* src, float* dst, int rows, int L, int d, int grid) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<BLOCKED, WRITE><<<grid, 512>>>(src, dst, rows, L, d, 1);
  cudaDeviceSynchronize();
  const int iters = 20;
  cudaEventRecord(e0);
  k<BLOCKED, WRITE><<<grid, 512>>>(src, dst, rows, L, d, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = (double)rows * L * d * 4 * iters * (WRITE ? 2 : 1);
  printf("%-52s grid %4d  %7.1f GB/s\n", name, grid, bytes / (ms * 1e-3) / 1e9);
}

int main() {
  const int rows = 6272 * 4, L = 6, d = 512;     // 4 x configs[1]: 308 MB, larger than L2
  float *src, *dst;
  cudaMalloc(&src, (size_t)rows * L * d * 4); cudaMalloc(&dst, (size_t)rows * L * d * 4);
  cudaMemset(src, 0, (size_t)rows * L * d * 4);
  for (int grid : {148, 296, 592}) {
    run<0, 0>("state layout (rows, L, d), read", src, dst, rows, L, d, grid);
    run<1, 0>("tile-contiguous layout, read", src, dst, rows, L, d, grid);
    run<0, 1>("state layout, read + write", src, dst, rows, L, d, grid);
    run<1, 1>("tile-contiguous layout, read + write", src, dst, rows, L, d, grid);
  }
  return 0;
}
