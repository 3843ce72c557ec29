// Pieces shared by the tensor-core kernels (tc_kernels.cu: K1 / K2 / K3, mlp_kernel.cu: the merged persistent MLP
// kernel): tile constants, the K1 / K2 epilogue chunk bodies and the tensor-map helpers.
#pragma once
#include "engine.h"
#include "ptx.cuh"

#include <stdio.h>

namespace glom {

constexpr int BM = 128;            // UMMA M (rows of the state per tile)
constexpr int BK = 64;             // bf16 elements per 128-byte swizzle row
constexpr uint32_t A_STAGE_BYTES = BM * BK * 2;   // 16 KB

// Sum of squares of a 32-column chunk row, in the canonical order shared with prep_state_kernel:
// 8 lanes hold 4 consecutive columns each (sequential fmaf), then an xor tree over the 8 lanes.
__device__ __forceinline__ float row_chunk_sumsq(float a, float b, float c, float d) {
  float q = a * a;
  q = fmaf(b, b, q);
  q = fmaf(c, c, q);
  q = fmaf(d, d, q);
  q += __shfl_xor_sync(0xffffffffu, q, 1);
  q += __shfl_xor_sync(0xffffffffu, q, 2);
  q += __shfl_xor_sync(0xffffffffu, q, 4);
  return q;
}

// ---- K1 epilogue chunk: 32 rows x 32 columns.  Row-per-thread bias + exact-erf GELU + bf16 pack, transpose
// through the warp's 2 KB patch (16-byte chunk c of row r stored at chunk c ^ ((r >> 1) & 3)), then 64-byte
// row segments out (8 rows x 64 B per store instruction).
// HPOL: 0 = streaming stores (two-kernel step: H is consumed by the NEXT launch, keep it out of L2's way),
//       1 = L2 evict-last policy `pol` (merged MLP kernel: H is consumed ~10 us later by GEMM2 tiles of the same launch
//           and must survive the rest of the traffic until then; the consumer's evict-first loads demote it again)
template <bool FULL, int HPOL = 0>
__device__ __forceinline__ void k1_chunk(const uint32_t (&v)[32], const float* bias, uint8_t* patch,
                                         __nv_bfloat16* hdst /* &H[row0][col] */, size_t pitch, int lane, int rows_left,
                                         uint64_t pol = 0) {
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    // bias slice read per use (broadcast LDS.128): 32 fewer live registers, so the polynomial's constant pairs stay in
    // registers instead of being re-materialised for every pair
    const float4 b = *reinterpret_cast<const float4*>(bias + 4 * i);
    pk[2 * i] = gelu_pair_bf16(__uint_as_float(v[4 * i + 0]), __uint_as_float(v[4 * i + 1]), b.x, b.y);
    pk[2 * i + 1] = gelu_pair_bf16(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]), b.z, b.w);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(patch + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) =
        make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
  __syncwarp();
  const int c = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + (lane >> 2);
    const uint4 val = *reinterpret_cast<const uint4*>(patch + r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
    // streaming (evict-first) stores: H (369 MB per step) never fits L2, and letting it through the normal policy
    // evicts the state shadows and weights the GEMMs and the consensus kernel re-read (measured: K1 -4 %)
    if (FULL || r < rows_left) {
      if (HPOL == 0) __stcs(reinterpret_cast<uint4*>(hdst + (size_t)r * pitch + c * 8), val);
      else st_global_v4_hint(hdst + (size_t)r * pitch + c * 8, val, pol);
    }
  }
  __syncwarp();
}

// ---- K2 epilogue chunk: the 4-way combine (glom_pytorch.py:141-142) on a 32 x 32 accumulator chunk.
// The accumulators go through the warp's 4 KB patch (f32, 128-byte rows, chunk c of row r at c ^ (r & 7)) so
// that each lane then owns 4 consecutive columns of 8 rows and every global access covers whole 128-byte lines.
struct K2Chunk {
  int l, L, d, n, row0;
  int prow0;            // row0 % n: patch index of the band's first row (position table row), computed once per tile
  int s_bcast;          // 1: s32_in is init_levels (L, d), the same for every row (first step of a call without carried state)
  const float* s32_in; const __nv_bfloat16* c_in; const float* pos;
  float* s32_out; __nv_bfloat16* sb_out; __nv_bfloat16* sp_out;
};
template <bool FULL>
__device__ __forceinline__ void k2_chunk(const uint32_t (&v)[32], const float4 b4, uint8_t* patch, const K2Chunk& k,
                                         int col, int lane, int rows_left, float (&rowsq)[8]) {
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<uint4*>(patch + lane * 128 + ((c ^ (lane & 7)) << 4)) =
        make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
  __syncwarp();
  const int c = lane & 7, rsub = lane >> 3;
  const bool top = (k.l == k.L - 1);                  // 3 contributions on the top level, 4 elsewhere (:128-129)
  const bool has_td = (k.l >= 1);
  const size_t ld = (size_t)k.L * k.d;
  const size_t base = ((size_t)k.row0 * k.L + k.l) * k.d + col + c * 4;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    // global loads of four rows first (independent, all in flight), then combine + store
    float4 sv[4], pp[4];
    uint2 cw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = (h * 4 + j) * 4 + rsub;
      sv[j] = make_float4(0.f, 0.f, 0.f, 0.f); pp[j] = sv[j]; cw[j] = make_uint2(0u, 0u);
      if (FULL || r < rows_left) {
        sv[j] = k.s_bcast ? __ldg(reinterpret_cast<const float4*>(k.s32_in + (size_t)k.l * k.d + col + c * 4))
                          : __ldcs(reinterpret_cast<const float4*>(k.s32_in + base + (size_t)r * ld));
        cw[j] = __ldcs(reinterpret_cast<const uint2*>(k.c_in + base + (size_t)r * ld));
        if (has_td) {
          int pr = k.prow0 + r;                       // (row0 + r) % n without a division per row (r < 32)
          if (k.n >= 32) { if (pr >= k.n) pr -= k.n; } else pr %= k.n;
          pp[j] = __ldg(reinterpret_cast<const float4*>(k.pos + (size_t)pr * k.d + col + c * 4));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = h * 4 + j;
      const int r = i * 4 + rsub;
      const float4 acc = *reinterpret_cast<const float4*>(patch + r * 128 + ((c ^ (r & 7)) << 4));
      float o0 = (sv[j].x + (acc.x + b4.x)) + __uint_as_float(cw[j].x << 16);               // (:141)
      float o1 = (sv[j].y + (acc.y + b4.y)) + __uint_as_float(cw[j].x & 0xFFFF0000u);
      float o2 = (sv[j].z + (acc.z + b4.z)) + __uint_as_float(cw[j].y << 16);
      float o3 = (sv[j].w + (acc.w + b4.w)) + __uint_as_float(cw[j].y & 0xFFFF0000u);
      if (top) { o0 = o0 / 3.0f; o1 = o1 / 3.0f; o2 = o2 / 3.0f; o3 = o3 / 3.0f; }          // (:142) IEEE division
      else { o0 *= 0.25f; o1 *= 0.25f; o2 *= 0.25f; o3 *= 0.25f; }                          // x/4 == x*0.25 exactly
      if (FULL || r < rows_left) {
        const size_t o = base + (size_t)r * ld;
        __stcs(reinterpret_cast<float4*>(k.s32_out + o), make_float4(o0, o1, o2, o3));
        *reinterpret_cast<uint2*>(k.sb_out + o) = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
        if (has_td)
          *reinterpret_cast<uint2*>(k.sp_out + ((size_t)(k.row0 + r) * (k.L - 1) + (k.l - 1)) * k.d + col + c * 4) =
              make_uint2(pack_bf16x2(o0 + pp[j].x, o1 + pp[j].y), pack_bf16x2(o2 + pp[j].z, o3 + pp[j].w));
      } else {
        o0 = o1 = o2 = o3 = 0.f;
      }
      rowsq[i] += row_chunk_sumsq(o0, o1, o2, o3);
    }
  }
  __syncwarp();
}

static inline bool encode_map(EncodeTiledFn enc, CUtensorMap* m, const void* base, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes /* rank-1 */, const uint32_t* box, char* err, size_t errlen,
                       const char* what) {
  cuuint64_t gd[3]; cuuint64_t gs[2]; cuuint32_t bx[3]; cuuint32_t es[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
    return false;
  }
  return true;
}

static inline bool map2d(EncodeTiledFn enc, CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                  char* err, size_t errlen, const char* what) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t strides[1] = {cols * 2};
  const uint32_t box[2] = {(uint32_t)BK, box_rows};
  return encode_map(enc, m, base, 2, dims, strides, box, err, errlen, what);
}

}  // namespace glom
