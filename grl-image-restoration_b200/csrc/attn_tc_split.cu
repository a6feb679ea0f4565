// attn_tc_split.cu -- EXPERIMENTAL variant of attn_tc.cu (same math, same operands, same barriers): every query row
// is handled by TWO threads, each owning 32 of the 64 keys of a tile and 16 of the 32 output columns.
//
// Why: the one-row-per-thread kernel is latency-bound (profiles/r1_tc_path_final.md): a softmax warp needs ~5500 cycles
// for the ~480 instructions of a tile and only 12 such warps fit on an SM, because a row's 64 logits + 32 outputs pin
// 128 registers per thread.  Halving the per-thread state (32 logits + 16 outputs) doubles the number of softmax warps
// per CTA at the same or a lower register count, i.e. 16 (2 CTAs / SM) or 24 (3 CTAs / SM) warps per SM to overlap each
// other's tcgen05.ld / bias-load / MUFU latencies.  The price is one 4-byte exchange of the partial row maximum per tile
// between the two threads of a row (shared memory + a 64-thread named barrier) and one exchange of the denominator at
// the end.
//
// Selected at run time by GRL_ATTN_SPLIT=1 (2 CTAs / SM, <= 112 registers) or =2 (3 CTAs / SM, <= 72 registers);
// unset = attn_tc.cu.  Not yet measured on hardware: off by default.
//
// Warps 0-7: softmax.  warp & 3 = TMEM lane quarter (rows 32*(warp&3) ..+31), warp >> 2 = key half.  Warp 8: producer
// + MMA issuer, identical to attn_tc.cu (cp.async gathers with roll / partition addressing, QK(t+1) at s_free(t),
// PV(t) at p_full(t)).
#include <stdlib.h>

#include "attn_tc.cuh"
#include "grl_common.cuh"
#include "ops_f32.h"
#include "ops_tc.h"
#include "tc_common.cuh"

namespace grl {
namespace tc {

namespace {

constexpr int kSplitKT = 64;
constexpr int kSplitSoftmaxWarps = 8;
constexpr int kSplitThreads = kSplitSoftmaxWarps * 32 + 32;
constexpr int kHalfKeys = kSplitKT / 2;  // keys per thread per tile
constexpr int kHalfOut = kDP / 2;        // output columns per thread

struct SplitSmem : AttnSmem<kSplitKT> {
  static constexpr int OFF_MX = OFF_BAR + 128;            // float mx[2 (tile parity)][2 (half)][128 rows]
  static constexpr int OFF_DEN = OFF_MX + 2 * 2 * kQT * 4;  // float den[2 (half)][128 rows]
  static constexpr int TOTAL = OFF_DEN + 2 * kQT * 4 + 1024;
};

__device__ __forceinline__ void pair_barrier(int quarter) {  // the two warps that share TMEM lane quarter `quarter`
  switch (quarter) {  // immediate barrier ids: a register operand would make ptxas reserve all 16 named barriers
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
}

template <int KW, int VAR, int MINB>
__global__ void __launch_bounds__(kSplitThreads, MINB) attn_tc_split_kernel(const AttnTcArgs a) {
  constexpr int KT = kSplitKT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using S = SplitSmem;
  uint8_t* Qs = smem;
  uint8_t* Ks = smem + S::OFF_K;
  uint8_t* Vs = smem + S::OFF_V;
  uint8_t* Ps = smem + S::OFF_P;
  int* koff_s = reinterpret_cast<int*>(smem + S::OFF_META);  // [3][KT]
  int* krid_s = koff_s + 3 * KT;                              // [3][KT]
  uint64_t* bar_s = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);  // QK(t) complete              (tcgen05.commit)
  uint64_t* bar_o = bar_s + 1;                                        // PV(t) complete              (tcgen05.commit)
  uint64_t* s_free = bar_s + 2;                                       // S_t read by all 256 threads
  uint64_t* p_full = bar_s + 3;                                       // P_t written, O_{t-1} consumed (256 arrivals)
  uint64_t* meta_full = bar_s + 4;                                    // [3] koff / rid of tile t    (32 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s + 8);
  float* mx_s = reinterpret_cast<float*>(smem + S::OFF_MX);
  float* den_s = reinterpret_cast<float*>(smem + S::OFF_DEN);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
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
  constexpr uint32_t TMEM_COLS = 128;  // S: columns [0, 64), O: [64, 96)
  constexpr int kThreadsSoftmax = kSplitSoftmaxWarps * 32;

  if (tid == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_init(s_free, kThreadsSoftmax);
    mbar_init(p_full, kThreadsSoftmax);
    for (int i = 0; i < 3; ++i) mbar_init(&meta_full[i], 32);
    mbar_init_fence();
  }
  if (warp == kSplitSoftmaxWarps) tmem_alloc(tmem_slot, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr int fmt = (VAR & 1) ? FMT_BF16 : FMT_F16;

  if (warp == kSplitSoftmaxWarps) {
    // =============================================================== producer + MMA issuer (as in attn_tc.cu)
    auto load_q = [&]() {
      for (int r = lane; r < kQT; r += 32) {
        const int qi = qt * kQT + r;
        const bool ok = qi < Nq;
        const Tok tq = locate(a.gq, wr, wc, ok ? qi : 0);
        const __nv_bfloat16* src = a.q + ((long long)(b * a.gq.H + tq.y) * a.gq.W + tq.x) * a.ldq + a.q_off + h * kDP;
#pragma unroll
        for (int c = 0; c < 4; ++c) cp_async_16(Qs + sw64(r, c), src + c * 8, ok);
      }
    };
    auto load_k = [&](int tile) {
      const int buf = tile & 1, k0 = tile * KT, slot = tile % 3;
      for (int r = lane; r < KT; r += 32) {
        const int kj = k0 + r;
        const bool ok = kj < Nk;
        const Tok tk = locate(a.gk, wr, wc, ok ? kj : 0);
        const __nv_bfloat16* ksrc = a.k + ((long long)(b * a.gk.H + tk.y) * a.gk.W + tk.x) * a.ldk + a.k_off + h * kDP;
        uint8_t* kd = Ks + buf * S::KV_BYTES;
#pragma unroll
        for (int c = 0; c < 4; ++c) cp_async_16(kd + sw64(r, c), ksrc + c * 8, ok);
        koff_s[slot * KT + r] = tk.ih * Wt + tk.iw;
        krid_s[slot * KT + r] = region_id(a.gk, tk.r, tk.c);
      }
      mbar_arrive(&meta_full[slot]);
    };
    auto load_v = [&](int tile) {
      const int buf = tile & 1, k0 = tile * KT;
      for (int r = lane; r < KT; r += 32) {
        const int kj = k0 + r;
        const bool ok = kj < Nk;
        const __nv_bfloat16* vsrc;
        if (a.v_dense) {
          vsrc = a.v + (((long long)bw * a.heads + h) * Nk + (ok ? kj : 0)) * kDP;
        } else {
          const Tok tk = locate(a.gk, wr, wc, ok ? kj : 0);
          vsrc = a.v + ((long long)(b * a.gk.H + tk.y) * a.gk.W + tk.x) * a.ldv + a.v_off + h * kDP;
        }
        uint8_t* vd = Vs + buf * S::KV_BYTES;
#pragma unroll
        for (int c = 0; c < 4; ++c) cp_async_16(vd + sw64(r, c), vsrc + c * 8, ok);
      }
    };
    const uint32_t idesc_qk = umma_idesc(kQT, KT, fmt, 0, 0);
    const uint32_t idesc_pv = umma_idesc(kQT, kDP, fmt, 0, 1);
    const uint32_t q_sa = smem_u32(Qs), p_sa = smem_u32(Ps);
    auto issue_qk = [&](int tile) {  // lane 0 only
      const uint32_t k_sa = smem_u32(Ks + (tile & 1) * S::KV_BYTES);
#pragma unroll
      for (int k = 0; k < kDP / 16; ++k)
        umma_ss(tmem, umma_desc(q_sa + k * 32, 16, 512, SWZ_64B), umma_desc(k_sa + k * 32, 16, 512, SWZ_64B), idesc_qk, k != 0);
      umma_commit(bar_s);
    };
    auto publish_all = [&]() {
      cp_async_wait<0>();
      fence_proxy_async_smem();
      __syncwarp();
    };
    auto publish_but_last = [&]() {
      cp_async_wait<1>();
      fence_proxy_async_smem();
      __syncwarp();
    };

    load_q();
    load_k(0);
    if (ntiles > 1) load_k(1);
    load_v(0);
    cp_async_commit();
    publish_all();
    if (lane == 0) {
      tcgen05_fence_after();
      issue_qk(0);
    }
    for (int t = 0; t < ntiles; ++t) {
      if (t + 1 < ntiles) {  // QK(t+1) once every thread has pulled its half of S_t out of TMEM
        publish_but_last();
        mbar_wait(s_free, t & 1);
        if (lane == 0) {
          tcgen05_fence_after();
          issue_qk(t + 1);
        }
      }
      if (t + 2 < ntiles) load_k(t + 2);
      cp_async_commit();
      publish_but_last();  // V_t landed
      mbar_wait(p_full, t & 1);
      if (lane == 0) {
        tcgen05_fence_after();
        const uint32_t v_sa = smem_u32(Vs + (t & 1) * S::KV_BYTES);
        const uint32_t pt_sa = p_sa + (t & 1) * S::P_BYTES;
#pragma unroll
        for (int k = 0; k < KT / 16; ++k)
          umma_ss(tmem + KT, umma_desc(pt_sa + (k & 3) * 32, 16, 1024, SWZ_128B), umma_desc(v_sa + k * 1024, 16, 512, SWZ_64B),
                  idesc_pv, k != 0);
        umma_commit(bar_o);
      }
      if (t + 1 < ntiles) load_v(t + 1);
      cp_async_commit();
    }
    mbar_wait(bar_o, (ntiles - 1) & 1);  // keep TMEM alive until the last MMA is done
  } else {
    // =============================================================== softmax warps: two threads per query row
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;  // row of the query tile == TMEM lane
    const int col0 = half * kHalfKeys;    // first key column (within a tile) of this thread
    const int qi = qt * kQT + row;
    const bool q_ok = qi < Nq;
    const Tok tq = locate(a.gq, wr, wc, q_ok ? qi : 0);
    const long long q_tok = (long long)(b * a.gq.H + tq.y) * a.gq.W + tq.x;
    const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16);
    const float* bias_h = a.bias + (size_t)h * 4 * a.rows_pad;
    const int base_i = (tq.ih + a.gk.wh - 1) * Wt + tq.iw + a.gk.ww - 1;
    const int q_rid = region_id(a.gq, tq.r, tq.c);
    const bool need_mask = a.use_mask && (wr == nwh - 1 || wc == nww - 1);
    constexpr bool ones = (VAR & 2) != 0;

    float o[kHalfOut];
#pragma unroll
    for (int e = 0; e < kHalfOut; ++e) o[e] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    float corr_prev = 0.f;

    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1, k0 = t * KT, slot = t % 3;
      const int kc = k0 + col0;  // first key of this thread in this tile
      mbar_wait(bar_s, t & 1);
      tcgen05_fence_after();
      const bool full_tile = (KW > 0) && (k0 + KT <= Nk);
      if (!full_tile || need_mask) mbar_wait(&meta_full[slot], (t / 3) & 1);

      // ---- logits of this thread's 32 keys (log2 domain): S from TMEM + bias
      float lg[kHalfKeys];
      {
        uint32_t v[32];
        tmem_ld32(trow + col0, v);
        tmem_ld_wait();
        if (full_tile) {
          constexpr int KWS = KW > 0 ? KW : 4;
          constexpr int RW = (KWS >= 32) ? 32 : KWS;  // consecutive keys of one key row inside the chunk
#pragma unroll
          for (int r0 = 0; r0 < 32; r0 += RW) {
            const int kj = kc + r0;  // first key of the run (warp-uniform, multiple of 4)
            const int s0 = base_i - ((kj / KWS) * Wt + (kj % KWS)) - 3;  // table index of key kj + 3
            const int cpy = (-s0) & 3;
            const float4* bp = reinterpret_cast<const float4*>(bias_h + (size_t)cpy * a.rows_pad + (s0 + cpy));
#pragma unroll
            for (int qd = 0; qd < RW / 4; ++qd) {
              const float4 bb = __ldg(bp - qd);
              const int j = r0 + 4 * qd;
              lg[j + 0] = __uint_as_float(v[j + 0]) + bb.w;
              lg[j + 1] = __uint_as_float(v[j + 1]) + bb.z;
              lg[j + 2] = __uint_as_float(v[j + 2]) + bb.y;
              lg[j + 3] = __uint_as_float(v[j + 3]) + bb.x;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            lg[j] = __uint_as_float(v[j]) + __ldg(bias_h + base_i - koff_s[slot * KT + col0 + j]);
        }
      }
      tcgen05_fence_before();
      mbar_arrive(s_free);  // this thread no longer needs S_t in TMEM

      if (need_mask) {
#pragma unroll
        for (int j = 0; j < kHalfKeys; ++j)
          if (krid_s[slot * KT + col0 + j] != q_rid) lg[j] += kMaskLog2;
      }
      if (k0 + KT > Nk) {
#pragma unroll
        for (int j = 0; j < kHalfKeys; ++j)
          if (kc + j >= Nk) lg[j] = -INFINITY;
      }
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < kHalfKeys; j += 4) {
        mx[0] = fmaxf(mx[0], lg[j]), mx[1] = fmaxf(mx[1], lg[j + 1]);
        mx[2] = fmaxf(mx[2], lg[j + 2]), mx[3] = fmaxf(mx[3], lg[j + 3]);
      }
      // ---- row maximum across the two halves.  Slot parity t & 1: the partner's read of tile t is ordered before its
      // arrival at the pair barrier of tile t + 1, which this thread passes before it writes the slot again at t + 2.
      const float m_loc = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      float* mxt = mx_s + (t & 1) * (2 * kQT);
      mxt[half * kQT + row] = m_loc;
      pair_barrier(quarter);
      const float m_new = fmaxf(m_run, fmaxf(m_loc, mxt[(half ^ 1) * kQT + row]));
      const float corr = ex2(m_run - m_new);
      m_run = m_new;

      float ps[4] = {0.f, 0.f, 0.f, 0.f};
      uint8_t* Pt = Ps + buf * S::P_BYTES;  // [128 x 64] K-major SWIZZLE_128B
#pragma unroll
      for (int c = 0; c < kHalfKeys / 8; ++c) {
        float p[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          p[e] = ex2(lg[c * 8 + e] - m_new);
          if (!ones) ps[e & 3] += p[e];
        }
        uint4 pk;
        if (fmt == FMT_BF16)
          pk = make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
        else
          pk = make_uint4(pack_f16(p[0], p[1]), pack_f16(p[2], p[3]), pack_f16(p[4], p[5]), pack_f16(p[6], p[7]));
        const int cc = half * (kHalfKeys / 8) + c;  // 16-byte chunk of the 128-byte row
        *reinterpret_cast<uint4*>(Pt + row * 128 + ((cc ^ (row & 7)) << 4)) = pk;
      }
      l_run = l_run * corr + ((ps[0] + ps[1]) + (ps[2] + ps[3]));  // partial denominator of this half (same reference max)

      // ---- fold in this thread's 16 columns of the previous tile's P V
      if (t > 0) {
        mbar_wait(bar_o, (t - 1) & 1);
        tcgen05_fence_after();
        uint32_t v[kHalfOut];
        tmem_ld16(trow + KT + half * kHalfOut, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < kHalfOut; ++e) o[e] = fmaf(o[e], corr_prev, __uint_as_float(v[e]));
      }
      corr_prev = corr;
      tcgen05_fence_before();
      fence_proxy_async_smem();  // P_t (generic-proxy stores) -> visible to the tensor core
      mbar_arrive(p_full);
    }
    {
      mbar_wait(bar_o, (ntiles - 1) & 1);
      tcgen05_fence_after();
      uint32_t v[kHalfOut];
      tmem_ld16(trow + KT + half * kHalfOut, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < kHalfOut; ++e) o[e] = fmaf(o[e], corr_prev, __uint_as_float(v[e]));
    }
    // ---- denominator: ones-column -> output column 31 (held by half 1); otherwise the sum of the two partial row sums
    den_s[half * kQT + row] = ones ? (half == 1 ? o[kHalfOut - 1] : 0.f) : l_run;
    pair_barrier(quarter);
    if (q_ok) {
      const float inv = 1.0f / (den_s[row] + den_s[kQT + row]);
      __nv_bfloat16* dst = (a.o_dense ? a.out + (((long long)bw * a.heads + h) * Nq + qi) * kDP
                                      : a.out + q_tok * a.ldo + a.o_off + h * kDP) +
                           half * kHalfOut;
#pragma unroll
      for (int e = 0; e < kHalfOut; e += 8)
        *reinterpret_cast<uint4*>(dst + e) =
            make_uint4(pack16(o[e] * inv, o[e + 1] * inv, fmt), pack16(o[e + 2] * inv, o[e + 3] * inv, fmt),
                       pack16(o[e + 4] * inv, o[e + 5] * inv, fmt), pack16(o[e + 6] * inv, o[e + 7] * inv, fmt));
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == kSplitSoftmaxWarps) {
    tcgen05_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

template <int KW, int VAR, int MINB>
int launch_split_one(const AttnTcArgs& a, unsigned nblk, cudaStream_t st) {
  auto kern = attn_tc_split_kernel<KW, VAR, MINB>;
  static bool configured = false;
  if (!configured) {
    GRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SplitSmem::TOTAL));
    configured = true;
  }
  kern<<<nblk, kSplitThreads, SplitSmem::TOTAL, st>>>(a);
  GRL_LAUNCH_CHECK("attn_tc_split_kernel");
  return GRL_OK;
}

template <int KW, int MINB>
int launch_split_var(const AttnTcArgs& a, unsigned nblk, cudaStream_t st) {
  switch ((a.fmt == FMT_BF16 ? 1 : 0) | (a.ones_col ? 2 : 0)) {
    case 0: return launch_split_one<KW, 0, MINB>(a, nblk, st);
    case 1: return launch_split_one<KW, 1, MINB>(a, nblk, st);
    case 2: return launch_split_one<KW, 2, MINB>(a, nblk, st);
    default: return launch_split_one<KW, 3, MINB>(a, nblk, st);
  }
}

template <int MINB>
int launch_split_kw(const AttnTcArgs& a, unsigned nblk, cudaStream_t st) {
  switch (a.gk.ww) {
    case 8: return launch_split_var<8, MINB>(a, nblk, st);
    case 16: return launch_split_var<16, MINB>(a, nblk, st);
    case 32: return launch_split_var<32, MINB>(a, nblk, st);
    case 64: return launch_split_var<64, MINB>(a, nblk, st);
    case 128: return launch_split_var<128, MINB>(a, nblk, st);
    default: return launch_split_var<0, MINB>(a, nblk, st);
  }
}

}  // namespace

// mode 1: 2 CTAs / SM (<= 112 registers), mode 2: 3 CTAs / SM (<= 72 registers).  Arguments already validated by
// launch_attn_tc.
int launch_attn_tc_split(const AttnTcArgs& a, unsigned nblk, int mode, cudaStream_t st) {
  return mode == 2 ? launch_split_kw<3>(a, nblk, st) : launch_split_kw<2>(a, nblk, st);
}

}  // namespace tc
}  // namespace grl
