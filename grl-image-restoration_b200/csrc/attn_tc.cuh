// attn_tc.cuh -- constants and shared-memory layout common to the attention kernels (attn_tc.cu, attn_tc_split.cu).
#pragma once
#include <stdint.h>

#include "tc_common.cuh"

namespace grl {
namespace tc {

constexpr int kQT = 128;
constexpr int kDP = 32;  // padded head dim (slot width)
constexpr float kMaskLog2 = -100.0f * 1.4426950408889634f;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// byte offset of 16-byte chunk c of row r in a 64-byte-row SWIZZLE_64B tile
__device__ __forceinline__ uint32_t sw64(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

template <int KT>
struct AttnSmem {
  static constexpr int Q_BYTES = kQT * 64;
  static constexpr int KV_BYTES = KT * 64;
  static constexpr int P_BYTES = kQT * KT * 2;
  static constexpr int OFF_K = Q_BYTES;
  static constexpr int OFF_V = OFF_K + 2 * KV_BYTES;
  static constexpr int OFF_P = (OFF_V + 2 * KV_BYTES + 1023) / 1024 * 1024;  // two P buffers
  static constexpr int OFF_META = OFF_P + 2 * P_BYTES;                       // int koff[3][KT], rid[3][KT] (tile % 3)
  static constexpr int OFF_BAR = OFF_META + 6 * KT * 4;
  static constexpr int TOTAL = OFF_BAR + 128 + 1024;
  static_assert(P_BYTES % 1024 == 0, "P tiles must be 1024-byte aligned");
};

}  // namespace tc
}  // namespace grl
