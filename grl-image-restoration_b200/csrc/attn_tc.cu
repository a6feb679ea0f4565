// attn_tc.cu -- fused cosine attention on tcgen05 tensor cores (the throughput path of WindowAttention and both
// passes of AnchorStripeAttention; mixed_attn_block_efficient.py:77-94,:128-165,:215-270).
//
// Inputs are the packed bf16 head slots written by the QKV / anchor projection epilogue (gemm_tc.cu, EPI_QKV):
// every head is a 32-wide slot (head_dim zero-padded), q^ / k^ / a^ are already L2-normalised and the side that
// carries the learned logit scale is pre-multiplied by exp(min(logit_scale, ln 100)) * log2(e); the bias table is
// 16*sigmoid(CPB(.))*log2(e).  So   S' = Q K^T   (one tcgen05.mma, fp32 in TMEM)   and
// P = exp2(S' + bias' + mask' - rowmax')   is softmax(cos*scale + bias + mask) exactly.
//
// One CTA = one (window|stripe, head, 128-query tile); one query row per thread (TMEM lane == row, tcgen05.ld
// 32x32b), keys stream through shared memory in tiles of KT (cp.async gather with the roll / partition address
// arithmetic of grl_geometry.h folded in, 64-byte swizzle so the tiles are valid UMMA operands as they land).
//   S  = Q K^T        A = Q [128 x 32] K-major SW64,  B = K tile [KT x 32] K-major SW64     -> TMEM cols [0, KT)
//   P                 softmax numerators, bf16, written to smem as [128 x KT] K-major SW128
//   Oj = P V          A = P,  B = V tile [KT keys x 32] MN-major SW64                        -> TMEM cols [KT, KT+32)
// The running output lives in registers (o = o * corr + Oj), so TMEM is never read-modify-written.
// Several CTAs are co-resident per SM (3 at KT = 64), which is what overlaps one CTA's MMAs with another's softmax.
#include "grl_common.cuh"
#include "ops_f32.h"
#include "ops_tc.h"
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
  static constexpr int OFF_P = OFF_V + 2 * KV_BYTES;
  static constexpr int OFF_META = OFF_P + P_BYTES;           // int koff[2][KT], rid[2][KT]
  static constexpr int OFF_BAR = OFF_META + 4 * KT * 4;
  static constexpr int TOTAL = OFF_BAR + 64 + 1024;
  static_assert(OFF_P % 1024 == 0, "P tile must be 1024-byte aligned for SWIZZLE_128B");
};

// KW: key-window width when it is a power of two that tiles KT (bias addressing with immediates), 0 = generic
template <int KT, int KW>
__global__ void __launch_bounds__(kQT, 3) attn_tc_kernel(const AttnTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using S = AttnSmem<KT>;
  uint8_t* Qs = smem;
  uint8_t* Ks = smem + S::OFF_K;
  uint8_t* Vs = smem + S::OFF_V;
  uint8_t* Ps = smem + S::OFF_P;
  int* koff_s = reinterpret_cast<int*>(smem + S::OFF_META);  // [2][KT]
  int* krid_s = koff_s + 2 * KT;                              // [2][KT]
  uint64_t* bar_s = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* bar_o = bar_s + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_o + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int Nq = a.gq.wh * a.gq.ww, Nk = a.gk.wh * a.gk.ww;
  const int nqt = (Nq + kQT - 1) / kQT;
  const int nww = a.gq.W / a.gq.ww;
  const int nwh = a.gq.H / a.gq.wh;
  const int nW = nwh * nww;
  int bid = blockIdx.x;
  const int qt = bid % nqt;
  bid /= nqt;
  const int h = bid % a.heads;
  const int bw = bid / a.heads;
  const int b = bw / nW, w = bw - b * nW;
  const int wr = w / nww, wc = w - wr * nww;
  const int Wt = a.gq.ww + a.gk.ww - 1;
  const int ntiles = (Nk + KT - 1) / KT;
  constexpr uint32_t TMEM_COLS = (KT + kDP <= 128) ? 128 : 256;

  if (tid == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_init_fence();
  }
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);

  // ---- this thread's query row
  const int qi = qt * kQT + tid;
  const bool q_ok = qi < Nq;
  const Tok tq = locate(a.gq, wr, wc, q_ok ? qi : 0);
  const long long q_tok = (long long)(b * a.gq.H + tq.y) * a.gq.W + tq.x;
  {
    const __nv_bfloat16* src = a.q + q_tok * a.ldq + a.q_off + h * kDP;
#pragma unroll
    for (int c = 0; c < 4; ++c) cp_async_16(Qs + sw64(tid, c), src + c * 8, q_ok);
  }
  // ---- K / V tile loader: thread t < KT gathers key row t of the tile (K and V)
  auto load_tile = [&](int tile, int buf) {
    const int k0 = tile * KT;
    for (int r = tid; r < KT; r += kQT) {
      const int kj = k0 + r;
      const bool ok = kj < Nk;
      const Tok tk = locate(a.gk, wr, wc, ok ? kj : 0);
      const long long tok = (long long)(b * a.gk.H + tk.y) * a.gk.W + tk.x;
      const __nv_bfloat16* ksrc = a.k + tok * a.ldk + a.k_off + h * kDP;
      const __nv_bfloat16* vsrc = a.v_dense ? a.v + (((long long)bw * a.heads + h) * Nk + (ok ? kj : 0)) * kDP
                                            : a.v + tok * a.ldv + a.v_off + h * kDP;
      uint8_t* kd = Ks + buf * S::KV_BYTES;
      uint8_t* vd = Vs + buf * S::KV_BYTES;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        cp_async_16(kd + sw64(r, c), ksrc + c * 8, ok);
        cp_async_16(vd + sw64(r, c), vsrc + c * 8, ok);
      }
      koff_s[buf * KT + r] = tk.ih * Wt + tk.iw;
      krid_s[buf * KT + r] = region_id(a.gk, tk.r, tk.c);
    }
  };
  load_tile(0, 0);
  cp_async_commit();
  if (ntiles > 1) load_tile(1, 1);
  cp_async_commit();

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);

  // bias row base for this query:  idx(i, j) = base_i - koff_j   (grl_geometry.h rel_index)
  const float* bias_h = a.bias + (size_t)h * a.rows;
  const int base_i = (tq.ih + a.gk.wh - 1) * Wt + tq.iw + a.gk.ww - 1;
  const int q_rid = region_id(a.gq, tq.r, tq.c);
  const bool need_mask = a.use_mask && (wr == nwh - 1 || wc == nww - 1);

  float o[kDP];
#pragma unroll
  for (int e = 0; e < kDP; ++e) o[e] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;

  const uint32_t idesc_qk = umma_idesc(kQT, KT, a.fmt, 0, 0);
  const uint32_t idesc_pv = umma_idesc(kQT, kDP, a.fmt, 0, 1);
  const int fmt = a.fmt;
  const uint32_t q_sa = smem_u32(Qs), p_sa = smem_u32(Ps);

  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    const int k0 = t * KT;
    if (t + 1 < ntiles) cp_async_wait<1>(); else cp_async_wait<0>();
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tcgen05_fence_after();
      const uint32_t k_sa = smem_u32(Ks + buf * S::KV_BYTES);
#pragma unroll
      for (int k = 0; k < kDP / 16; ++k)
        umma_ss(tmem, umma_desc(q_sa + k * 32, 16, 512, SWZ_64B), umma_desc(k_sa + k * 32, 16, 512, SWZ_64B), idesc_qk,
                k != 0);
      umma_commit(bar_s);
    }
    mbar_wait(bar_s, t & 1);
    tcgen05_fence_after();

    // ---- logits of this tile (log2 domain), row max
    float lg[KT];
    float m_tile = -INFINITY;
#pragma unroll
    for (int c0 = 0; c0 < KT; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(trow + c0, v);
      tmem_ld_wait();
      if (KW > 0 && k0 + KT <= Nk) {  // full tile: immediate-offset runs (tail tiles take the clamped generic path)
        constexpr int KWS = KW > 0 ? KW : 1;  // (KW == 0 never reaches this branch)
        constexpr int RW = (KWS >= 32) ? 32 : KWS;  // run length of consecutive keys in one key row inside this chunk
#pragma unroll
        for (int r0 = 0; r0 < 32; r0 += RW) {
          const int kj = k0 + c0 + r0;  // first key of the run (CTA-uniform)
          const float* bp = bias_h + base_i - ((kj / KWS) * Wt + (kj % KWS));
#pragma unroll
          for (int j = 0; j < RW; ++j) lg[c0 + r0 + j] = __uint_as_float(v[r0 + j]) + __ldg(bp - j);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          lg[c0 + j] = __uint_as_float(v[j]) + __ldg(bias_h + base_i - koff_s[buf * KT + c0 + j]);
      }
    }
    if (need_mask) {
#pragma unroll
      for (int j = 0; j < KT; ++j)
        if (krid_s[buf * KT + j] != q_rid) lg[j] += kMaskLog2;
    }
    if (k0 + KT > Nk) {
#pragma unroll
      for (int j = 0; j < KT; ++j)
        if (k0 + j >= Nk) lg[j] = -INFINITY;
    }
#pragma unroll
    for (int j = 0; j < KT; ++j) m_tile = fmaxf(m_tile, lg[j]);
    const float m_new = fmaxf(m_run, m_tile);
    const float corr = ex2(m_run - m_new);
    m_run = m_new;
    float psum = 0.f;
#pragma unroll
    for (int c = 0; c < KT / 8; ++c) {
      float p[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        p[e] = ex2(lg[c * 8 + e] - m_new);
        psum += p[e];
      }
      const uint4 pk = make_uint4(pack16(p[0], p[1], fmt), pack16(p[2], p[3], fmt), pack16(p[4], p[5], fmt), pack16(p[6], p[7], fmt));
      // [128 x KT] K-major SWIZZLE_128B, 64-key sub-tiles of 16 KB
      const int sub = c >> 3, cc = c & 7;
      *reinterpret_cast<uint4*>(Ps + sub * (kQT * 128) + tid * 128 + ((cc ^ (tid & 7)) << 4)) = pk;
    }
    l_run = l_run * corr + psum;

    tcgen05_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tcgen05_fence_after();
      const uint32_t v_sa = smem_u32(Vs + buf * S::KV_BYTES);
#pragma unroll
      for (int k = 0; k < KT / 16; ++k) {
        const uint32_t pa = p_sa + (k >> 2) * (kQT * 128) + (k & 3) * 32;
        umma_ss(tmem + KT, umma_desc(pa, 16, 1024, SWZ_128B), umma_desc(v_sa + k * 1024, 16, 512, SWZ_64B), idesc_pv,
                k != 0);
      }
      umma_commit(bar_o);
    }
    mbar_wait(bar_o, t & 1);
    tcgen05_fence_after();
    {
      uint32_t v[32];
      tmem_ld32(trow + KT, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < kDP; ++e) o[e] = fmaf(o[e], corr, __uint_as_float(v[e]));
    }
    tcgen05_fence_before();
    __syncthreads();  // buffers `buf`, P and the TMEM columns are free again
    if (t + 2 < ntiles) load_tile(t + 2, buf);
    cp_async_commit();
  }

  if (q_ok) {
    const float inv = 1.0f / l_run;
    __nv_bfloat16* dst = a.o_dense ? a.out + (((long long)bw * a.heads + h) * Nq + qi) * kDP
                                   : a.out + q_tok * a.ldo + a.o_off + h * kDP;
#pragma unroll
    for (int e = 0; e < kDP; e += 8)
      *reinterpret_cast<uint4*>(dst + e) =
          make_uint4(pack16(o[e] * inv, o[e + 1] * inv, fmt), pack16(o[e + 2] * inv, o[e + 3] * inv, fmt),
                     pack16(o[e + 4] * inv, o[e + 5] * inv, fmt), pack16(o[e + 6] * inv, o[e + 7] * inv, fmt));
  }
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

template <int KT, int KW>
static int launch_attn_one(const AttnTcArgs& a, unsigned nblk, cudaStream_t st) {
  auto kern = attn_tc_kernel<KT, KW>;
  static bool configured = false;
  if (!configured) {
    GRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<KT>::TOTAL));
    configured = true;
  }
  kern<<<nblk, kQT, AttnSmem<KT>::TOTAL, st>>>(a);
  GRL_LAUNCH_CHECK("attn_tc_kernel");
  return GRL_OK;
}

int launch_attn_tc(const AttnTcArgs& a, cudaStream_t st) {
  if (a.B == 0) return GRL_OK;
  int rc;
  if ((rc = check_grid(a.gq, "attn_tc(q grid)")) != GRL_OK) return rc;
  if ((rc = check_grid(a.gk, "attn_tc(k grid)")) != GRL_OK) return rc;
  GRL_REQUIRE(a.gq.H / a.gq.wh == a.gk.H / a.gk.wh && a.gq.W / a.gq.ww == a.gk.W / a.gk.ww,
              "attn_tc: query and key grids have different window counts");
  GRL_REQUIRE(a.heads >= 1 && a.heads <= 8, "attn_tc: heads=%d unsupported", a.heads);
  const int Nq = a.gq.wh * a.gq.ww;
  const long long nblk = (long long)a.B * (a.gq.H / a.gq.wh) * (a.gq.W / a.gq.ww) * a.heads * ceil_div(Nq, kQT);
  GRL_REQUIRE(nblk < (1ll << 31), "attn_tc: grid too large");
  constexpr int KT = 64;
  switch (a.gk.ww) {
    case 8: return launch_attn_one<KT, 8>(a, (unsigned)nblk, st);
    case 16: return launch_attn_one<KT, 16>(a, (unsigned)nblk, st);
    case 32: return launch_attn_one<KT, 32>(a, (unsigned)nblk, st);
    case 64: return launch_attn_one<KT, 64>(a, (unsigned)nblk, st);
    case 128: return launch_attn_one<KT, 128>(a, (unsigned)nblk, st);
    default: return launch_attn_one<KT, 0>(a, (unsigned)nblk, st);
  }
}

}  // namespace tc
}  // namespace grl
