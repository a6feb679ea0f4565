// gemm_tc.cu -- bf16 tcgen05 GEMM / implicit-GEMM 3x3 convolution with fused epilogues (the non-attention half
// of a GRL block on the throughput path).
//
//   D[128 x BN] (fp32, TMEM) = A[128 x K] (bf16, TMA -> smem, SWIZZLE_128B) * W[BN x K]^T (bf16, TMA -> smem)
//
// A is either a row-major (tokens x Kpad) activation matrix (nn.Linear: QKVProjection, AnchorLinear, proj, Mlp) or
// the channels-last image itself read through a 4-D tensor map: one CTA owns an 8x16 pixel patch and each of the
// 9 taps is the same TMA box shifted by (dy, dx) -- the zero padding of the convolution is TMA's out-of-bounds
// fill, no im2col buffer exists (CAB convs mixed_attn_block.py:973-977, TransformerStage.conv grl.py:164-170).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2-5 = epilogue (one accumulator row per thread, read with tcgen05.ld 32x32b).
// Two CTAs are co-resident per SM (<= 100 KB smem, <= 256 TMEM columns each) so one CTA's epilogue overlaps the
// other's loads and MMAs; these GEMMs are short-K and HBM/epilogue bound, not tensor bound.
//
// Epilogues:
//   EPI_BIAS_ACT : y = act(acc + b) (+ res)                       -> bf16 and/or fp32     (fc1, CAB, convs, heads)
//   EPI_QKV      : per 32-wide head slot  y = (acc + b) * scale / max(||.||, 1e-12)  -> bf16 (q^, k^, a^; v untouched)
//                  (F.normalize + logit scale of Attention.attn / AffineTransform, efficient.py:39,:85)
//   EPI_LN       : x' = x + rs * LayerNorm(acc + b) (+ cab_y * gate) -> fp32 residual stream + bf16 operand copy
//                  (efficient.py:543-554)
#include "grl_common.cuh"
#include "tc_common.cuh"
#include "ops_tc.h"

namespace grl {
namespace tc {

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

constexpr int kStages = 2;
constexpr int kBM = 128, kBK = 64;
constexpr int kTH = 8, kTW = 16;  // conv patch (kTH * kTW == kBM)

template <int BN>
struct GemmSmem {
  static constexpr int A_BYTES = kBM * kBK * 2;
  static constexpr int B_BYTES = BN * kBK * 2;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int TOTAL = kStages * STAGE + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ uint32_t tmem_cols_for(int bn) { return bn <= 32 ? 32 : bn <= 64 ? 64 : bn <= 128 ? 128 : 256; }

template <int BN, int EPI, bool CONV>
__global__ void __launch_bounds__(192, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using S = GemmSmem<BN>;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * S::STAGE);
  uint64_t* empty = full + kStages;
  uint64_t* tmem_full = empty + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * BN;
  const int nk_total = a.taps * a.nk;

  // tile coordinates
  int m0 = 0, tb = 0, ty0 = 0, tx0 = 0;
  if (CONV) {
    int t = blockIdx.x;
    const int tx = t % a.tiles_x;
    t /= a.tiles_x;
    const int ty = t % a.tiles_y;
    tb = t / a.tiles_y;
    ty0 = ty * kTH;
    tx0 = tx * kTW;
  } else {
    m0 = blockIdx.x * kBM;
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init_fence();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols_for(BN));
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kc = 0; kc < nk_total; ++kc) {
        const int s = kc % kStages;
        const uint32_t ph = (kc / kStages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sa = smem + s * S::STAGE;
        uint8_t* sb = sa + S::A_BYTES;
        mbar_expect_tx(&full[s], S::STAGE);
        if (CONV) {
          const int tap = kc / a.nk, c0 = (kc - tap * a.nk) * kBK;
          tma_load_4d(sa, &tmA, &full[s], c0, tx0 + (tap % 3) - 1, ty0 + (tap / 3) - 1, tb);
        } else {
          tma_load_2d(sa, &tmA, &full[s], kc * kBK, m0);
        }
        tma_load_2d(sb, &tmB, &full[s], kc * kBK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(kBM, BN, 1, 0, 0);
      for (int kc = 0; kc < nk_total; ++kc) {
        const int s = kc % kStages;
        const uint32_t ph = (kc / kStages) & 1;
        mbar_wait(&full[s], ph);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + s * S::STAGE);
        const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          const uint64_t ad = umma_desc(sa + k * 32, 16, 1024, SWZ_128B);
          const uint64_t bd = umma_desc(sb + k * 32, 16, 1024, SWZ_128B);
          umma_ss(tmem, ad, bd, idesc, (kc | k) != 0);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(tmem_full);
    }
  } else {
    // ------------------------------------------------------------------ epilogue: one accumulator row per thread
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;
    long long tok;
    bool row_ok;
    if (CONV) {
      const int y = ty0 + r / kTW, x = tx0 + r % kTW;
      row_ok = (y < a.H) && (x < a.W);
      tok = ((long long)tb * a.H + y) * a.W + x;
    } else {
      tok = (long long)m0 + r;
      row_ok = tok < a.M;
    }
    mbar_wait(tmem_full, 0);
    tcgen05_fence_after();
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t v[32];

    if (EPI == EPI_LN) {
      // pass 1: mean, pass 2: variance (two-pass like ATen's LayerNorm), pass 3: normalise + residual
      const int C = a.C;
      float sum = 0.f;
      for (int c0 = 0; c0 < C; c0 += 32) {
        tmem_ld32(trow + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c0 + j < C) sum += __uint_as_float(v[j]) + a.bias[c0 + j];
      }
      const float mean = sum / (float)C;
      float var = 0.f;
      for (int c0 = 0; c0 < C; c0 += 32) {
        tmem_ld32(trow + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c0 + j < C) {
            const float d = __uint_as_float(v[j]) + a.bias[c0 + j] - mean;
            var = fmaf(d, d, var);
          }
      }
      const float rstd = rsqrtf(var / (float)C + a.eps);
      const long long bimg = row_ok ? tok / a.L : 0;
      for (int c0 = 0; c0 < BN; c0 += 32) {
        tmem_ld32(trow + c0, v);
        tmem_ld_wait();
        if (!row_ok) continue;
        float o[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int c = c0 + j;
          float val = 0.f;
          if (c < C) {
            val = (__uint_as_float(v[j]) + a.bias[c] - mean) * rstd * a.gamma[c] + a.beta[c];
            val = val * a.res_scale + a.res_f32[tok * a.ldr + c];
            if (a.cab_y) val += __bfloat162float(a.cab_y[tok * a.ld_caby + c]) * a.cab_gate[bimg * C + c];
          }
          o[j] = val;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int c = c0 + j;
          if (c + 3 < C) {
            *reinterpret_cast<float4*>(a.out_f32 + tok * a.ldo_f32 + c) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
          } else {
            for (int e = 0; e < 4; ++e)
              if (c + e < C) a.out_f32[tok * a.ldo_f32 + c + e] = o[j + e];
          }
        }
        if (c0 < a.ldo_bf16) {  // operand copy, zero in the pad channels
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 pk = make_uint4(pack_bf16(o[j], o[j + 1]), pack_bf16(o[j + 2], o[j + 3]), pack_bf16(o[j + 4], o[j + 5]),
                                  pack_bf16(o[j + 6], o[j + 7]));
            if (c0 + j < a.ldo_bf16) *reinterpret_cast<uint4*>(a.out_bf16 + tok * a.ldo_bf16 + c0 + j) = pk;
          }
        }
      }
    } else {
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= a.N) break;  // warp-uniform
        tmem_ld32(trow + c0, v);
        tmem_ld_wait();
        if (!row_ok) continue;
        float o[32];
        if (EPI == EPI_QKV) {
          const int slot = (n0 + c0) >> 5;
          float ss = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            o[j] = __uint_as_float(v[j]) + a.bias[n0 + c0 + j];
            ss = fmaf(o[j], o[j], ss);
          }
          const float sc = a.slot_scale[slot];
          const float mul = sc > 0.f ? sc / fmaxf(sqrtf(ss), 1e-12f) : 1.0f;  // scale <= 0 marks a value slot
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] *= mul;
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + c0 + j;
            float val = __uint_as_float(v[j]) + (n < a.N ? a.bias[n] : 0.f);
            val = apply_act(val, a.act, a.slope);
            if (a.res_f32 && n < a.N_f32) val += a.res_f32[tok * a.ldr + n];
            o[j] = (n < a.N) ? val : 0.f;
          }
        }
        if (a.out_bf16) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (n0 + c0 + j < a.ldo_bf16) {
              uint4 pk = make_uint4(pack_bf16(o[j], o[j + 1]), pack_bf16(o[j + 2], o[j + 3]),
                                    pack_bf16(o[j + 4], o[j + 5]), pack_bf16(o[j + 6], o[j + 7]));
              *reinterpret_cast<uint4*>(a.out_bf16 + tok * a.ldo_bf16 + n0 + c0 + j) = pk;
            }
          }
        }
        if (EPI == EPI_BIAS_ACT && a.out_f32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + c0 + j;
            if (n < a.N_f32) a.out_f32[tok * a.ldo_f32 + n] = o[j];
          }
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem, tmem_cols_for(BN));
  }
}

// -------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------
static int make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                    const cuuint32_t* box) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(GRL_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes,
                   box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return fail(GRL_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)rc);
  return GRL_OK;
}

template <int BN, int EPI, bool CONV>
static int launch_one(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmTcArgs& a, dim3 grid, cudaStream_t st) {
  auto kern = gemm_tc_kernel<BN, EPI, CONV>;
  static bool configured = false;
  if (!configured) {
    GRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<BN>::TOTAL));
    configured = true;
  }
  kern<<<grid, 192, GemmSmem<BN>::TOTAL, st>>>(tmA, tmB, a);
  GRL_LAUNCH_CHECK("gemm_tc_kernel");
  return GRL_OK;
}

template <int EPI, bool CONV>
static int dispatch_bn(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmTcArgs& a, dim3 grid,
                       cudaStream_t st) {
  switch (bn) {
    case 64: return launch_one<64, EPI, CONV>(tmA, tmB, a, grid, st);
    case 128: return launch_one<128, EPI, CONV>(tmA, tmB, a, grid, st);
    case 192: return launch_one<192, EPI, CONV>(tmA, tmB, a, grid, st);
    case 256: return launch_one<256, EPI, CONV>(tmA, tmB, a, grid, st);
  }
  return fail(GRL_ERR_INVALID, "gemm_tc: unsupported tile width %d", bn);
}

int pick_bn(int npad) {
  if (npad <= 64) return 64;
  if (npad <= 128) return 128;
  if (npad % 192 == 0 || npad <= 192) return 192;
  if (npad % 256 == 0) return 256;
  return npad % 128 == 0 ? 128 : 192;
}

// x: bf16 (M, Kpad) row-major or (B, H, W, Kpad) channels-last; w: bf16 (Npad, taps*Kpad) K-major.
int launch_gemm_tc(const GemmTcProblem& p, GemmTcArgs a, cudaStream_t st) {
  GRL_REQUIRE(p.kpad % kBK == 0 && p.kpad > 0, "gemm_tc: K pad %d must be a multiple of 64", p.kpad);
  GRL_REQUIRE(p.npad % 32 == 0 && p.npad > 0, "gemm_tc: N pad %d must be a multiple of 32", p.npad);
  const int bn = (p.epi == EPI_LN) ? (p.npad <= 64 ? 64 : p.npad <= 128 ? 128 : p.npad <= 192 ? 192 : 256) : pick_bn(p.npad);
  GRL_REQUIRE(p.epi != EPI_LN || p.npad <= 256, "gemm_tc: LayerNorm epilogue needs the whole row in one tile (N=%d)",
              p.npad);
  const bool conv = p.taps == 9;
  GRL_REQUIRE(p.taps == 1 || p.taps == 9, "gemm_tc: taps must be 1 or 9");
  CUtensorMap tmA, tmB;
  int rc;
  dim3 grid;
  a.nk = p.kpad / kBK;
  a.taps = p.taps;
  if (conv) {
    cuuint64_t dims[4] = {(cuuint64_t)p.kpad, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    cuuint64_t str[3] = {(cuuint64_t)p.kpad * 2, (cuuint64_t)p.W * p.kpad * 2, (cuuint64_t)p.H * p.W * p.kpad * 2};
    cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)kTW, (cuuint32_t)kTH, 1};
    if ((rc = make_map(&tmA, p.x, 4, dims, str, box)) != GRL_OK) return rc;
    a.H = p.H, a.W = p.W;
    a.tiles_x = ceil_div(p.W, kTW), a.tiles_y = ceil_div(p.H, kTH);
    a.M = (long long)p.B * p.H * p.W;
    grid = dim3((unsigned)(a.tiles_x * a.tiles_y * p.B), ceil_div(p.npad, bn));
  } else {
    cuuint64_t dims[2] = {(cuuint64_t)p.kpad, (cuuint64_t)p.M};
    cuuint64_t str[1] = {(cuuint64_t)p.kpad * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)kBM};
    if ((rc = make_map(&tmA, p.x, 2, dims, str, box)) != GRL_OK) return rc;
    a.M = p.M;
    grid = dim3((unsigned)ceil_div(p.M, kBM), ceil_div(p.npad, bn));
  }
  if (a.M == 0) return GRL_OK;
  {
    cuuint64_t dims[2] = {(cuuint64_t)p.kpad * p.taps, (cuuint64_t)p.npad};
    cuuint64_t str[1] = {(cuuint64_t)p.kpad * p.taps * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)bn};
    if ((rc = make_map(&tmB, p.w, 2, dims, str, box)) != GRL_OK) return rc;
  }
  switch (p.epi) {
    case EPI_BIAS_ACT:
      return conv ? dispatch_bn<EPI_BIAS_ACT, true>(bn, tmA, tmB, a, grid, st)
                  : dispatch_bn<EPI_BIAS_ACT, false>(bn, tmA, tmB, a, grid, st);
    case EPI_QKV:
      GRL_REQUIRE(!conv, "gemm_tc: QKV epilogue is linear-only");
      return dispatch_bn<EPI_QKV, false>(bn, tmA, tmB, a, grid, st);
    case EPI_LN:
      GRL_REQUIRE(!conv, "gemm_tc: LN epilogue is linear-only");
      return dispatch_bn<EPI_LN, false>(bn, tmA, tmB, a, grid, st);
  }
  return fail(GRL_ERR_INVALID, "gemm_tc: unknown epilogue %d", p.epi);
}

}  // namespace tc
}  // namespace grl
