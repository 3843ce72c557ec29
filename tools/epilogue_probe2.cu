// tools/epilogue_probe.cu -- which part of the GEMM1 epilogue slows the main loop down?
//
// mainloop_probe.cu: TMA ring + tcgen05.mma cta_group::2 256x256x64 k-blocks with no epilogue run at 88 % of the
// tensor peak; the real GEMM1 kernel (K = 512: 8 k-blocks per 256 x 256 tile, bias + exact-erf GELU + bf16 H out) at
// ~60 %.  This probe adds the epilogue back one component at a time (16 epilogue warps per CTA, two TMEM accumulator
// stages, the tfull / tempty hand-off of gemm_kernel) and reports the MMA rate for each combination:
//   T  tcgen05.ld of the warp's 32 x 64 accumulator part                 (TMEM read-out)
//   A  + bias + GELU + bf16 packing on those values                      (ALU / XU work)
//   S  + the 2 KB per-warp transpose through shared memory (STS.128 x 4, LDS.128 x 4 per 32 x 32 chunk)
//   G  + 64-byte row-segment global stores of the packed tile            (LSU / L2 write traffic)
//   R  row-per-thread global stores straight from registers instead of S + G (no shared-memory transpose)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/epilogue_probe tools/epilogue_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../glom_pytorch_b200/csrc/ptx.cuh"

using namespace glom;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);


__device__ __forceinline__ void tma_load_2d_2sm_a(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(m), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}

// probe-only GELU variants: (X) the two MUFU.EX2 replaced by one FFMA2 (no MUFU at all; result is meaningless),
// (W) ex2 in software on the FMA pipe (Cody-Waite split + degree-3 polynomial + exponent insertion)
__device__ __forceinline__ uint64_t poly5(uint64_t u) {
  uint64_t q = f2_fma(f2_pack(0.00036467931931838393f, 0.00036467931931838393f), u, f2_pack(0.006363349035382271f, 0.006363349035382271f));
  q = f2_fma(q, u, f2_pack(0.05013200640678406f, 0.05013200640678406f));
  q = f2_fma(q, u, f2_pack(-0.4617065489292145f, -0.4617065489292145f));
  q = f2_fma(q, u, f2_pack(1.150075078010559f, 1.150075078010559f));
  q = f2_fma(q, u, f2_pack(-1.0001276731491089f, -1.0001276731491089f));
  return q;
}
__device__ __forceinline__ uint32_t gelu_pair_nomufu(float acc0, float acc1, float bias0, float bias1) {
  const uint64_t x = f2_add(f2_pack(acc0, acc1), f2_pack(bias0, bias1));
  float x0, x1; f2_unpack(x, x0, x1);
  const uint64_t u = f2_pack(__uint_as_float(__float_as_uint(x0) | 0x80000000u), __uint_as_float(__float_as_uint(x1) | 0x80000000u));
  uint64_t q = poly5(u);
  const uint64_t e = f2_fma(q, q, f2_pack(0.5f, 0.5f));
  const uint64_t g = f2_fma(u, e, f2_pack(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f)));
  float g0, g1; f2_unpack(g, g0, g1);
  return pack_bf16x2(g0, g1);
}
__device__ __forceinline__ uint32_t gelu_pair_swexp(float acc0, float acc1, float bias0, float bias1) {
  const uint64_t x = f2_add(f2_pack(acc0, acc1), f2_pack(bias0, bias1));
  float x0, x1; f2_unpack(x, x0, x1);
  const uint64_t u = f2_pack(__uint_as_float(__float_as_uint(x0) | 0x80000000u), __uint_as_float(__float_as_uint(x1) | 0x80000000u));
  uint64_t q = poly5(u);
  float q0, q1; f2_unpack(q, q0, q1);
  q = f2_pack(fmaxf(q0, -125.f), fmaxf(q1, -125.f));
  const uint64_t magic = f2_pack(12582912.f, 12582912.f);
  const uint64_t t = f2_add(q, magic);
  const uint64_t n = f2_add(t, f2_pack(-12582912.f, -12582912.f));
  const uint64_t f = f2_fma(n, f2_pack(-1.f, -1.f), q);
  uint64_t p = f2_fma(f2_pack(0.05550357f, 0.05550357f), f, f2_pack(0.24022652f, 0.24022652f));
  p = f2_fma(p, f, f2_pack(0.69314720f, 0.69314720f));
  p = f2_fma(p, f, f2_pack(1.0f, 1.0f));
  float p0, p1, t0, t1; f2_unpack(p, p0, p1); f2_unpack(t, t0, t1);
  const float e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
  const float e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
  const uint64_t g = f2_fma(u, f2_pack(e0, e1), f2_pack(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f)));
  float g0, g1; f2_unpack(g, g0, g1);
  return pack_bf16x2(g0, g1);
}

struct Result { unsigned long long cycles, kblocks, wait_full, wait_acc; };

constexpr uint32_t STAGE_BYTES = 32768;
constexpr int STAGES = 5;
constexpr int EPI_WARPS = 16;
constexpr int THREADS = 32 * (EPI_WARPS + 4);
enum { F_T = 1, F_A = 2, F_S = 4, F_G = 8, F_R = 16, F_H = 32, F_B = 64, F_D = 128, F_X = 256, F_W = 512, F_L = 1024, F_P = 2048 };

template <int FLAGS, int DIET>
__global__ void __launch_bounds__(THREADS, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, Result* out, int tiles, int nkb,
             int a_rows_total, uint8_t* hbuf, size_t hbytes, const float* bias) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* patches = smem + (size_t)STAGES * STAGE_BYTES;
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float bias_s[256];
  const uint32_t rank = cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  constexpr int W_TMA = EPI_WARPS, W_MMA = EPI_WARPS + 1, W_ALLOC = EPI_WARPS + 2;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 2 * EPI_WARPS); }
    fence_barrier_init();
  }
  if (threadIdx.x < 256) bias_s[threadIdx.x] = bias[threadIdx.x];
  if (warp == W_ALLOC) tmem_alloc_2sm(&tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  if (DIET && warp == W_TMA) {
    // warp-converged producer: every lane polls, one elected lane issues (no per-instruction ELECT loops in SASS)
    const uint32_t elected = elect_one();
    int stage = 0; uint32_t phase = 0;
    const uint32_t bar0 = mapa_shared(smem_u32(&full_bar[0]), 0);
    const uint32_t smem0 = smem_u32(smem);
    const int b_row = (int)rank * 128;
    for (int t = 0; t < tiles; ++t) {
      const int a_row = (int)(((long long)(cluster_id + (long long)t * num_clusters) * 256) % (a_rows_total - 256)) / 256 * 256 + (int)rank * 128;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elected) {
          const uint32_t sa = smem0 + (uint32_t)stage * STAGE_BYTES;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
          const uint32_t bar = bar0 + 8u * (uint32_t)stage;
          tma_load_2d_2sm_a(sa, &map_a, bar, kb * 64, a_row);
          tma_load_2d_2sm_a(sa + 16384, &map_b, bar, kb * 64, b_row);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (DIET && warp == W_MMA && rank == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(256, 256, 0, 0);
    const uint32_t elected = elect_one();
    const uint64_t a_desc0 = umma_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t b_desc0 = umma_desc_sw128(smem_u32(smem) + 16384, 16, 1024);
    int stage = 0; uint32_t phase = 0;
    int as = 0; uint32_t aphase = 0;
    unsigned long long wf = 0, wa = 0;
    const long long c0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      long long t0 = clock64();
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      wa += (unsigned long long)(clock64() - t0);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + (uint32_t)as * 256u;
      for (int kb = 0; kb < nkb; ++kb) {
        t0 = clock64();
        mbar_wait(&full_bar[stage], phase);
        wf += (unsigned long long)(clock64() - t0);
        tc_fence_after_sync();
        if (elected) {
          const uint64_t ad = a_desc0 + (uint64_t)(stage * (STAGE_BYTES >> 4));
          const uint64_t bd = b_desc0 + (uint64_t)(stage * (STAGE_BYTES >> 4));
          umma_bf16_2sm(d_tmem, ad, bd, idesc, kb != 0 ? 1u : 0u);
          umma_bf16_2sm(d_tmem, ad + 2, bd + 2, idesc, 1u);
          umma_bf16_2sm(d_tmem, ad + 4, bd + 4, idesc, 1u);
          umma_bf16_2sm(d_tmem, ad + 6, bd + 6, idesc, 1u);
          umma_commit_2sm(&empty_bar[stage], 3);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (elected) umma_commit_2sm(&tfull_bar[as], 3);
      __syncwarp();
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (lane == 0) {
      out[blockIdx.x].cycles = (unsigned long long)(clock64() - c0);
      out[blockIdx.x].kblocks = (unsigned long long)tiles * nkb;
      out[blockIdx.x].wait_full = wf;
      out[blockIdx.x].wait_acc = wa;
    }
  } else if (!DIET && warp == W_TMA && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int t = 0; t < tiles; ++t) {
      const int a_row = (int)(((long long)(cluster_id + (long long)t * num_clusters) * 256) % (a_rows_total - 256)) / 256 * 256 + (int)rank * 128;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + (size_t)stage * STAGE_BYTES;
        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
        const uint32_t bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
        tma_load_2d_2sm(sa, &map_a, bar, kb * 64, a_row);
        tma_load_2d_2sm(sa + 16384, &map_b, bar, kb * 64, (int)rank * 128);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (!DIET && warp == W_MMA && lane == 0 && rank == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(256, 256, 0, 0);
    int stage = 0; uint32_t phase = 0;
    int as = 0; uint32_t aphase = 0;
    unsigned long long wf = 0, wa = 0;
    const long long c0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      long long t0 = clock64();
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      wa += (unsigned long long)(clock64() - t0);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + (uint32_t)as * 256u;
      for (int kb = 0; kb < nkb; ++kb) {
        t0 = clock64();
        mbar_wait(&full_bar[stage], phase);
        wf += (unsigned long long)(clock64() - t0);
        tc_fence_after_sync();
        const uint32_t a_addr = smem_u32(smem + (size_t)stage * STAGE_BYTES);
        const uint32_t b_addr = a_addr + 16384;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_2sm(d_tmem, umma_desc_sw128(a_addr + k * 32, 16, 1024), umma_desc_sw128(b_addr + k * 32, 16, 1024), idesc,
                        (kb | k) ? 1u : 0u);
        umma_commit_2sm(&empty_bar[stage], 3);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit_2sm(&tfull_bar[as], 3);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    out[blockIdx.x].cycles = (unsigned long long)(clock64() - c0);
    out[blockIdx.x].kblocks = (unsigned long long)tiles * nkb;
    out[blockIdx.x].wait_full = wf;
    out[blockIdx.x].wait_acc = wa;
  } else if ((FLAGS & F_P) && warp < EPI_WARPS) {
    // T+A+S+G with the TMEM read-out software-pipelined in 16-column quarters: the next quarter's tcgen05.ld is in flight
    // while the current one goes through the GELU
    const int quad = warp & 3, part = warp >> 2;
    uint8_t* patch = patches + (size_t)warp * 4096;
    int as = 0; uint32_t aphase = 0;
    for (int t = 0; t < tiles; ++t) {
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after_sync();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * 256 + part * 64);
      const size_t blk = ((size_t)(cluster_id + (size_t)t * num_clusters) * 2 + rank) * 4 + part;
      uint8_t* hrow = hbuf + (blk * 16384) % hbytes + (size_t)quad * 32 * 128;
      uint32_t qa[16], qb[16], pk[16];
      tmem_ld16(t_addr, qa);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        tmem_ld_wait();
        if (q < 3) { if (q & 1) tmem_ld16(t_addr + 16 * (q + 1), qa); else tmem_ld16(t_addr + 16 * (q + 1), qb); }
        const float* b = bias_s + part * 64 + q * 16;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t x0 = (q & 1) ? qb[2 * i] : qa[2 * i], x1 = (q & 1) ? qb[2 * i + 1] : qa[2 * i + 1];
          pk[8 * (q & 1) + i] = gelu_pair_bf16(__uint_as_float(x0), __uint_as_float(x1), b[2 * i], b[2 * i + 1]);
        }
        if (q & 1) {
          const int c0 = (q >> 1) * 32;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(patch + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          __syncwarp();
          const int c = lane & 3;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = i * 8 + (lane >> 2);
            const uint4 val = *reinterpret_cast<const uint4*>(patch + r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
            __stcs(reinterpret_cast<uint4*>(hrow + (size_t)r * 128 + c0 * 2 + c * 16), val);
          }
          __syncwarp();
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[as]), 0));
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else if ((FLAGS & F_D) && warp < EPI_WARPS) {
    // T+A+S+G with the warps split into two groups half a period apart: "late" warps (odd column parts) store a chunk's
    // packed results one chunk later, right after issuing the next TMEM load -- while the "early" warps are in their
    // math phase -- so the MUFU-bound math and the LSU-bound stores of the two groups overlap.
    const int quad = warp & 3, part = warp >> 2;
    const bool late = (FLAGS & F_R) ? false : (part & 1);
    uint8_t* patch = patches + (size_t)warp * 4096;
    int as = 0; uint32_t aphase = 0;
    uint32_t pend[16];
    uint8_t* pend_dst = nullptr;
    auto store_chunk = [&](const uint32_t (&pk)[16], uint8_t* dst) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(patch + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      __syncwarp();
      const int c = lane & 3;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = i * 8 + (lane >> 2);
        const uint4 val = *reinterpret_cast<const uint4*>(patch + r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
        __stcs(reinterpret_cast<uint4*>(dst + (size_t)r * 128 + c * 16), val);
      }
      __syncwarp();
    };
    for (int t = 0; t < tiles; ++t) {
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after_sync();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * 256 + part * 64);
      const size_t blk = ((size_t)(cluster_id + (size_t)t * num_clusters) * 2 + rank) * 4 + part;
      uint8_t* hrow = hbuf + (blk * 16384) % hbytes + (size_t)quad * 32 * 128;
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(t_addr + c0, v);
        if (late && pend_dst) store_chunk(pend, pend_dst);
        tmem_ld_wait();
        uint32_t pk[16];
        const float* b = bias_s + part * 64 + c0;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          pk[i] = gelu_pair_bf16(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]), b[2 * i], b[2 * i + 1]);
        if (late) {
#pragma unroll
          for (int i = 0; i < 16; ++i) pend[i] = pk[i];
          pend_dst = hrow + c0 * 2;
        } else {
          store_chunk(pk, hrow + c0 * 2);
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[as]), 0));
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (late && pend_dst) store_chunk(pend, pend_dst);
  } else if (warp < EPI_WARPS) {
    const int quad = warp & 3, part = warp >> 2;
    uint8_t* patch = patches + (size_t)warp * 4096;
    int as = 0; uint32_t aphase = 0;
    unsigned sink = 0;
    for (int t = 0; t < tiles; ++t) {
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after_sync();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * 256 + part * 64);
      // destination: a 16 KB block per (tile, CTA, part) of a ring-shaped H-like buffer
      const size_t blk = ((size_t)(cluster_id + (size_t)t * num_clusters) * 2 + rank) * 4 + part;
      uint8_t* hrow = hbuf + (blk * 16384) % hbytes + (size_t)quad * 32 * 128;
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        if (FLAGS & F_T) { tmem_ld32(t_addr + c0, v); tmem_ld_wait(); }
        else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0x3f000000u + lane + i + t;
        }
        uint32_t pk[16];
        if (FLAGS & F_H) {
          // half of the GELU's ALU work: same data flow, ~3.5 instead of ~7 instructions per element
          const float* b = bias_s + part * 64 + c0;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            uint64_t x = f2_add(f2_pack(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), f2_pack(b[2 * i], b[2 * i + 1]));
            uint64_t q = f2_fma(x, x, f2_pack(0.5f, 0.5f));
            q = f2_fma(q, x, f2_pack(0.25f, 0.25f));
            float q0, q1; f2_unpack(q, q0, q1);
            q0 = ex2_approx(q0);
            pk[i] = pack_bf16x2(q0, q1);
          }
        } else if (FLAGS & F_X) {
          const float* b = bias_s + part * 64 + c0;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            pk[i] = gelu_pair_nomufu(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]), b[2 * i], b[2 * i + 1]);
        } else if (FLAGS & F_W) {
          const float* b = bias_s + part * 64 + c0;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            pk[i] = gelu_pair_swexp(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]), b[2 * i], b[2 * i + 1]);
        } else if (FLAGS & F_A) {
          const float* b = bias_s + part * 64 + c0;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            pk[i] = gelu_pair_bf16(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]), b[2 * i], b[2 * i + 1]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = v[2 * i] ^ v[2 * i + 1];
        }
        if (FLAGS & F_B) {
          // row-per-thread into the warp's 4 KB staging block in the 128B-swizzled image (conflict-free), one bulk copy per tile
          if (c0 == 0) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
          }
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(patch + lane * 128 + ((((c0 >> 3) + c) ^ (lane & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          if (c0 == 32) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 4096;" ::"l"(hrow), "r"(smem_u32(patch)) : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
        } else if (FLAGS & F_L) {
          // full-line variant: both chunks are staged as 128-byte rows (16-byte chunk j of row r at j ^ (r & 7)), then every
          // STG.128 covers 4 whole 128-byte lines
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(patch + lane * 128 + ((((c0 >> 3) + c) ^ (lane & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          if (c0 == 32) {
            __syncwarp();
            const int c = lane & 7;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = i * 4 + (lane >> 3);
              const uint4 val = *reinterpret_cast<const uint4*>(patch + r * 128 + ((c ^ (r & 7)) << 4));
              __stcs(reinterpret_cast<uint4*>(hrow + (size_t)r * 128 + c * 16), val);
            }
            __syncwarp();
          }
        } else if (FLAGS & F_S) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(patch + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          __syncwarp();
          const int c = lane & 3;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = i * 8 + (lane >> 2);
            const uint4 val = *reinterpret_cast<const uint4*>(patch + r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
            if (FLAGS & F_G) __stcs(reinterpret_cast<uint4*>(hrow + (size_t)r * 128 + c0 * 2 + c * 16), val);
            else sink ^= val.x ^ val.y ^ val.z ^ val.w;
          }
          __syncwarp();
        } else if (FLAGS & F_R) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            __stcs(reinterpret_cast<uint4*>(hrow + (size_t)lane * 128 + c0 * 2 + c * 16), make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]));
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) sink ^= pk[i];
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[as]), 0));
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if ((FLAGS & F_B) && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (sink == 0x12345u) out[0].kblocks = sink;      // keep the synthetic work alive
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == W_ALLOC) { tc_fence_after_sync(); tmem_dealloc_2sm(tmem_base, 512); }
}

template <int FLAGS, int DIET>
static void run(const char* name, const CUtensorMap& ma, const CUtensorMap& mb, int sms, int nkb, int tiles, int a_rows_total,
                Result* dres, uint8_t* hbuf, size_t hbytes, const float* bias) {
  const size_t smem = 1024 + (size_t)STAGES * STAGE_BYTES + EPI_WARPS * 4096;
  cudaFuncSetAttribute(probe_kernel<FLAGS, DIET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(sms / 2 * 2); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(dres, 0, sizeof(Result) * sms);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe_kernel<FLAGS, DIET>, ma, mb, dres, tiles, nkb, a_rows_total, hbuf, hbytes, bias);
    cudaEventRecord(e1);
    e = e == cudaSuccess ? cudaDeviceSynchronize() : e;
    if (e != cudaSuccess) { printf("%s FAILED: %s\n", name, cudaGetErrorString(e)); exit(1); }
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  Result* h = (Result*)malloc(sizeof(Result) * sms);
  cudaMemcpy(h, dres, sizeof(Result) * sms, cudaMemcpyDeviceToHost);
  double cyc = 0, wf = 0, wa = 0;
  const int pairs = sms / 2;
  for (int i = 0; i < pairs; ++i) { cyc += (double)h[2 * i].cycles; wf += (double)h[2 * i].wait_full; wa += (double)h[2 * i].wait_acc; }
  cyc /= pairs; wf /= pairs; wa /= pairs;
  const double kb = (double)tiles * nkb;
  const double rate = kb * 256.0 * 256.0 * 64.0 / cyc / 2.0;
  printf("K = %4d  %-34s %5.0f MAC/clk/SM (%5.1f %%)  %6.0f clk per tile   MMA lane waits: operands %4.1f %%, accumulator %4.1f %%   %.2f ms, %4.0f TFLOP/s at %4.0f MHz\n",
         nkb * 64, name, rate, 100.0 * rate / 4096.0, cyc / tiles, 100.0 * wf / cyc, 100.0 * wa / cyc, ms,
         2.0 * kb * pairs * 256.0 * 256.0 * 64.0 / (ms * 1e-3) / 1e12, cyc / (ms * 1e-3) / 1e6);
  free(h);
}

int main(int argc, char** argv) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int a_rows_total = 16384, KMAX = 4096;
  void *a, *b;
  cudaMalloc(&a, (size_t)a_rows_total * KMAX * 2);
  cudaMalloc(&b, (size_t)256 * KMAX * 2);
  cudaMemset(a, 0, (size_t)a_rows_total * KMAX * 2);
  cudaMemset(b, 0, (size_t)256 * KMAX * 2);
  const size_t hbytes = (size_t)384 << 20;           // H-like destination, larger than L2
  uint8_t* hbuf;
  cudaMalloc(&hbuf, hbytes);
  float* bias;
  cudaMalloc(&bias, 1024);
  cudaMemset(bias, 0, 1024);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  CUtensorMap ma, mb;
  {
    cuuint64_t gd[2] = {(cuuint64_t)KMAX, (cuuint64_t)a_rows_total};
    cuuint64_t gs[1] = {(cuuint64_t)KMAX * 2};
    cuuint32_t bx[2] = {64, 128};
    cuuint32_t es[2] = {1, 1};
    if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 1;
    cuuint64_t gdb[2] = {(cuuint64_t)KMAX, 256};
    if (enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, b, gdb, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 1;
  }
  Result* dres;
  cudaMalloc(&dres, sizeof(Result) * sms);
  printf("GEMM1-shaped tiles (256 x 256 per CTA pair) with the epilogue added back piece by piece, %d SMs\n", sms);
  const int nkb = argc > 1 ? atoi(argv[1]) : 8, tiles = argc > 2 ? atoi(argv[2]) : 400;
#define RUN(F, D, NAME) run<F, D>(NAME, ma, mb, sms, nkb, tiles, a_rows_total, dres, hbuf, hbytes, bias)
  RUN(F_T | F_A | F_S | F_G, 1, "T+A+S+G");
  RUN(F_P, 1, "T+A+S+G, TMEM read-out pipelined in quarters");
  return 0;
}
