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
#include <stdlib.h>

#include "grl_common.cuh"
#include "ops_f32.h"
#include "ops_tc.h"
#include "attn_tc.cuh"
#include "tc_common.cuh"

namespace grl {
namespace tc {

// ---- differential-timing builds (tools/kernel_diag.py): each GRL_ATTN_DIAG_* define removes ONE ingredient of the kernel
// so that its cost shows up as a time difference.  The results of such a build are WRONG by construction; nothing but
// tools/kernel_diag.py defines these, and the default build is bit-for-bit the production kernel.
#ifdef GRL_ATTN_DIAG_NOBIAS  // no bias-table loads (logits = S)
#define GRL_DIAG_BIAS(ptr) make_float4(0.f, 0.f, 0.f, 0.f)
#else
#define GRL_DIAG_BIAS(ptr) __ldg(ptr)
#endif
#ifdef GRL_ATTN_DIAG_NOEXP  // no MUFU (P = logit - max, garbage but finite)
#define GRL_DIAG_EX2(x) (x)
#else
#define GRL_DIAG_EX2(x) ex2(x)
#endif
#ifdef GRL_ATTN_DIAG_NOPSTORE  // P never written to shared memory
#define GRL_DIAG_PSTORE(stmt)
#else
#define GRL_DIAG_PSTORE(stmt) stmt
#endif
#ifdef GRL_ATTN_DIAG_NOFOLD  // O_{t-1} never read back from TMEM inside the loop (the wait on bar_o stays)
#define GRL_DIAG_FOLD(block)
#else
#define GRL_DIAG_FOLD(block) block
#endif
#ifdef GRL_ATTN_DIAG_NOPFENCE  // no fence.proxy.async before p_full
#define GRL_DIAG_PFENCE(stmt)
#else
#define GRL_DIAG_PFENCE(stmt) stmt
#endif
#ifdef GRL_ATTN_DIAG_NOGATHER  // the producer gathers K / V only for the first tiles (the MMAs re-read stale tiles)
#define GRL_DIAG_GATHER(stmt)
#else
#define GRL_DIAG_GATHER(stmt) stmt
#endif

// One 64-byte row of a Q / K / V tile -> shared memory.  Default: the lane that owns the row issues its four 16-byte
// cp.async (a warp instruction then touches 32 different rows: 32 L1 tag requests, 32 shared-memory wavefronts).
// GRL_ATTN_QUAD_GATHER (A/B build, tools/kernel_diag.py): the row address is handed to a quad of lanes by shuffle and
// each warp instruction copies 8 whole rows (8 tag requests).  Same bytes, same layout, same results.
#ifdef GRL_ATTN_QUAD_GATHER
#define GRL_ROW_COPY(base, r, src, ok)                                                                      \
  do {                                                                                                      \
    const int r_first_ = (r) - lane;                                                                        \
    _Pragma("unroll") for (int sub_ = 0; sub_ < 4; ++sub_) {                                                \
      const int sl_ = sub_ * 8 + (lane >> 2);                                                               \
      const unsigned long long p_ = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)(src), sl_);    \
      const int ok_ = __shfl_sync(0xffffffffu, (int)(ok), sl_);                                             \
      cp_async_16((base) + sw64(r_first_ + sl_, lane & 3),                                                  \
                  reinterpret_cast<const __nv_bfloat16*>((uintptr_t)p_) + (lane & 3) * 8, ok_ != 0);        \
    }                                                                                                       \
  } while (0)
#else
#define GRL_ROW_COPY(base, r, src, ok)                                                      \
  do {                                                                                      \
    _Pragma("unroll") for (int c_ = 0; c_ < 4; ++c_) cp_async_16((base) + sw64((r), c_), (src) + c_ * 8, (ok)); \
  } while (0)
#endif

constexpr int kAttnThreads = kQT + 32;  // 4 softmax warps (one query row per thread) + 1 producer / MMA warp

// Warp-specialised pipeline without block-wide barriers in the loop (tiles t = 0..nt-1 of KT keys):
//   warp 4 (producer + MMA issuer): cp.async gathers of Q / K_t / V_t (roll + partition addressing), then
//       QK(t+1) as soon as every softmax thread has pulled S_t out of TMEM (s_free), PV(t) as soon as P_t is in smem
//       and O_{t-1} has been consumed (p_full); completion is signalled by tcgen05.commit on bar_s / bar_o.
//   warps 0-3 (softmax, thread = query row = TMEM lane): wait bar_s(t) -> S_t + bias -> registers -> arrive s_free ->
//       exp2 / running max / bf16|fp16 P_t -> smem -> wait bar_o(t-1), o = o * corr_{t-1} + O_{t-1} -> arrive p_full.
// The only waits of a softmax warp are on MMA completions.
// KW: key-window width when it is a power of two >= 8 (bias rows read as aligned float4 from the 4-way shifted table
// copies), 0 = generic scalar path.
// VAR: bit 0 = bf16 operands (else fp16), bit 1 = ones-column denominators -- compile-time so the exp / pack loop has no
// uniform branches (they were ~9 % of the softmax warps' issue slots).
template <int KT, int KW, int VAR>
__global__ void __launch_bounds__(kAttnThreads, (KT <= 32 ? 4 : KT <= 64 ? 3 : 2)) attn_tc_kernel(const AttnTcArgs a) {
  static_assert(KT == 32 || KT == 64 || KT == 128, "P tiles: 64-byte rows (SWIZZLE_64B) or 128-byte rows (SWIZZLE_128B)");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using S = AttnSmem<KT>;
  uint8_t* Qs = smem;
  uint8_t* Ks = smem + S::OFF_K;
  uint8_t* Vs = smem + S::OFF_V;
  uint8_t* Ps = smem + S::OFF_P;
  int* koff_s = reinterpret_cast<int*>(smem + S::OFF_META);  // [3][KT]
  int* krid_s = koff_s + 3 * KT;                              // [3][KT]
  uint64_t* bar_s = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);  // QK(t) complete          (tcgen05.commit)
  uint64_t* bar_o = bar_s + 1;                                        // PV(t) complete          (tcgen05.commit)
  uint64_t* s_free = bar_s + 2;                                       // S_t read by all rows    (128 arrivals)
  uint64_t* p_full = bar_s + 3;                                       // P_t written, O_{t-1} consumed (128 arrivals)
  uint64_t* meta_full = bar_s + 4;                                    // [3] koff / rid of tile t (32 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s + 8);

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
  constexpr uint32_t TMEM_COLS = (KT + kDP <= 64) ? 64 : (KT + kDP <= 128) ? 128 : 256;

  if (tid == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_init(s_free, kQT);
    mbar_init(p_full, kQT);
    for (int i = 0; i < 3; ++i) mbar_init(&meta_full[i], 32);
    mbar_init_fence();
  }
  if (warp == 4) tmem_alloc(tmem_slot, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr int fmt = (VAR & 1) ? FMT_BF16 : FMT_F16;

  if (warp == 4) {
    // =============================================================== producer + MMA issuer
    auto load_q = [&]() {
      for (int r = lane; r < kQT; r += 32) {
        const int qi = qt * kQT + r;
        const bool ok = qi < Nq;
        const Tok tq = locate(a.gq, wr, wc, ok ? qi : 0);
        const __nv_bfloat16* src = a.q + ((long long)(b * a.gq.H + tq.y) * a.gq.W + tq.x) * a.ldq + a.q_off + h * kDP;
        GRL_ROW_COPY(Qs, r, src, ok);
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
        GRL_ROW_COPY(kd, r, ksrc, ok);
        koff_s[slot * KT + r] = tk.ih * Wt + tk.iw;
        krid_s[slot * KT + r] = region_id(a.gk, tk.r, tk.c);
      }
      mbar_arrive(&meta_full[slot]);  // release: koff / rid of this tile are published
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
        GRL_ROW_COPY(vd, r, vsrc, ok);
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
    // this warp's gathers -- all of them, or all but the most recently committed group -- have landed and are
    // visible to the tensor core (async proxy)
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
      // ---- QK(t+1): K_{t+1} was requested an iteration ago; the S columns are free once every row has read S_t
      // commit order per iteration is  {K_{t+2}}, {V_{t+1}}:  K_{t+1} is in the second most recent group here
      if (t + 1 < ntiles) {
        publish_but_last();
        mbar_wait(s_free, t & 1);
        if (lane == 0) {
          tcgen05_fence_after();
          issue_qk(t + 1);
        }
      }
      // K buffer t&1 is free (QK(t) completed before anyone could read S_t)
      if (t + 2 < ntiles) GRL_DIAG_GATHER(load_k(t + 2));
      cp_async_commit();
      // ---- PV(t): V_t landed (second most recent group; K_{t+2} may still be in flight); P_t written and O_{t-1}
      // consumed by every row
      publish_but_last();
      mbar_wait(p_full, t & 1);
      if (lane == 0) {
        tcgen05_fence_after();
        const uint32_t v_sa = smem_u32(Vs + (t & 1) * S::KV_BYTES);
        const uint32_t pt_sa = p_sa + (t & 1) * S::P_BYTES;
#pragma unroll
        for (int k = 0; k < KT / 16; ++k) {
          const uint64_t pd = (KT == 32) ? umma_desc(pt_sa + k * 32, 16, 512, SWZ_64B)
                                         : umma_desc(pt_sa + (k >> 2) * (kQT * 128) + (k & 3) * 32, 16, 1024, SWZ_128B);
          umma_ss(tmem + KT, pd, umma_desc(v_sa + k * 1024, 16, 512, SWZ_64B), idesc_pv, k != 0);
        }
        umma_commit(bar_o);
      }
      // V buffer (t+1)&1 held V_{t-1}.  PV(t-1) is complete: every row waited for it before arriving on p_full(t).
      // (Waiting on bar_o here would be wrong as well as redundant: PV(t) may already have flipped the barrier
      // again, and a parity wait on the older phase would then sleep until a phase that needs this warp.)
      if (t + 1 < ntiles) GRL_DIAG_GATHER(load_v(t + 1));
      cp_async_commit();
    }
    mbar_wait(bar_o, (ntiles - 1) & 1);  // keep TMEM alive until the last MMA is done
  } else {
    // =============================================================== softmax warps: thread = query row
    const int qi = qt * kQT + tid;
    const bool q_ok = qi < Nq;
    const Tok tq = locate(a.gq, wr, wc, q_ok ? qi : 0);
    const long long q_tok = (long long)(b * a.gq.H + tq.y) * a.gq.W + tq.x;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    // bias:  idx(i, j) = base_i - koff_j  (grl_geometry.h rel_index); table copy c holds T shifted right by c entries
    const float* bias_h = a.bias + (size_t)h * 4 * a.rows_pad;
    const int base_i = (tq.ih + a.gk.wh - 1) * Wt + tq.iw + a.gk.ww - 1;
    const int q_rid = region_id(a.gq, tq.r, tq.c);
    const bool need_mask = a.use_mask && (wr == nwh - 1 || wc == nww - 1);
    // ones-column: when head_dim < 32 the projection epilogue sets column 31 of every V row to 1, so O[:, 31] =
    // sum_j P_ij is the softmax denominator -- accumulated by the tensor core from the very P it multiplies with V,
    // and rescaled together with the other columns; the 64 FADDs per tile of the explicit row sum disappear.
    constexpr bool ones = (VAR & 2) != 0;

    float o[kDP];
#pragma unroll
    for (int e = 0; e < kDP; ++e) o[e] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    float corr_prev = 0.f;  // exp2(m_{t-2} - m_{t-1}): brings o (relative to m_{t-2}) to the reference of O_{t-1}

    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1, k0 = t * KT, slot = t % 3;
      mbar_wait(bar_s, t & 1);
      tcgen05_fence_after();
      const bool full_tile = (KW > 0) && (k0 + KT <= Nk);
      if (!full_tile || need_mask) mbar_wait(&meta_full[slot], (t / 3) & 1);

      // ---- logits of this tile (log2 domain): S from TMEM + bias
      float lg[KT];
#pragma unroll
      for (int c0 = 0; c0 < KT; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(trow + c0, v);
        tmem_ld_wait();
        if (full_tile) {
          constexpr int KWS = KW > 0 ? KW : 4;
          constexpr int RW = (KWS >= 32) ? 32 : KWS;  // consecutive keys of one key row inside this chunk
#pragma unroll
          for (int r0 = 0; r0 < 32; r0 += RW) {
            const int kj = k0 + c0 + r0;  // first key of the run (CTA-uniform, multiple of 4)
            const int s0 = base_i - ((kj / KWS) * Wt + (kj % KWS)) - 3;  // table index of key kj + 3
            const int cpy = (-s0) & 3;
            const float4* bp = reinterpret_cast<const float4*>(bias_h + (size_t)cpy * a.rows_pad + (s0 + cpy));
#pragma unroll
            for (int qd = 0; qd < RW / 4; ++qd) {
              const float4 bb = GRL_DIAG_BIAS(bp - qd);
              const int j = c0 + r0 + 4 * qd;
              lg[j + 0] = __uint_as_float(v[r0 + 4 * qd + 0]) + bb.w;
              lg[j + 1] = __uint_as_float(v[r0 + 4 * qd + 1]) + bb.z;
              lg[j + 2] = __uint_as_float(v[r0 + 4 * qd + 2]) + bb.y;
              lg[j + 3] = __uint_as_float(v[r0 + 4 * qd + 3]) + bb.x;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            lg[c0 + j] = __uint_as_float(v[j]) + __ldg(bias_h + base_i - koff_s[slot * KT + c0 + j]);
        }
      }
      tcgen05_fence_before();
      mbar_arrive(s_free);  // this row no longer needs S_t in TMEM

      if (need_mask) {
#pragma unroll
        for (int j = 0; j < KT; ++j)
          if (krid_s[slot * KT + j] != q_rid) lg[j] += kMaskLog2;
      }
      if (k0 + KT > Nk) {
#pragma unroll
        for (int j = 0; j < KT; ++j)
          if (k0 + j >= Nk) lg[j] = -INFINITY;
      }
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < KT; j += 4) {
        mx[0] = fmaxf(mx[0], lg[j]), mx[1] = fmaxf(mx[1], lg[j + 1]);
        mx[2] = fmaxf(mx[2], lg[j + 2]), mx[3] = fmaxf(mx[3], lg[j + 3]);
      }
      const float m_new = fmaxf(fmaxf(m_run, fmaxf(mx[0], mx[1])), fmaxf(mx[2], mx[3]));
      const float corr = ex2(m_run - m_new);
      m_run = m_new;
      float ps[4] = {0.f, 0.f, 0.f, 0.f};  // independent partial sums (no 64-long dependent FADD chain)
      uint8_t* Pt = Ps + buf * S::P_BYTES;
#pragma unroll
      for (int c = 0; c < KT / 8; ++c) {
        float p[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          p[e] = GRL_DIAG_EX2(lg[c * 8 + e] - m_new);
          if (!ones) ps[e & 3] += p[e];
        }
        uint4 pk;
        if (fmt == FMT_BF16)
          pk = make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
        else
          pk = make_uint4(pack_f16(p[0], p[1]), pack_f16(p[2], p[3]), pack_f16(p[4], p[5]), pack_f16(p[6], p[7]));
        if (KT == 32) {  // [128 x 32] K-major: 64-byte rows, SWIZZLE_64B
          *reinterpret_cast<uint4*>(Pt + sw64(tid, c)) = pk;
        } else {  // [128 x KT] K-major SWIZZLE_128B, 64-key sub-tiles of 16 KB
          const int sub = c >> 3, cc = c & 7;
          GRL_DIAG_PSTORE(*reinterpret_cast<uint4*>(Pt + sub * (kQT * 128) + tid * 128 + ((cc ^ (tid & 7)) << 4)) = pk);
        }
      }
      l_run = l_run * corr + ((ps[0] + ps[1]) + (ps[2] + ps[3]));

      // ---- fold in the previous tile's P V (it has had a whole softmax to finish).  o is kept relative to the running
      // max at which the last folded O was computed, so the fold is one FFMA per element: o <- o * corr_{t-1} + O_{t-1}
      if (t > 0) {
        mbar_wait(bar_o, (t - 1) & 1);
        tcgen05_fence_after();
        GRL_DIAG_FOLD({
          uint32_t v[32];
          tmem_ld32(trow + KT, v);
          tmem_ld_wait();
          _Pragma("unroll") for (int e = 0; e < kDP; ++e) o[e] = fmaf(o[e], corr_prev, __uint_as_float(v[e]));
        })
      }
      corr_prev = corr;
      tcgen05_fence_before();
      GRL_DIAG_PFENCE(fence_proxy_async_smem());  // P_t (generic-proxy stores) -> visible to the tensor core
      mbar_arrive(p_full);
    }
    {
      mbar_wait(bar_o, (ntiles - 1) & 1);
      tcgen05_fence_after();
      uint32_t v[32];
      tmem_ld32(trow + KT, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < kDP; ++e) o[e] = fmaf(o[e], corr_prev, __uint_as_float(v[e]));  // now relative to the final max
    }
    if (q_ok) {
      const float inv = 1.0f / (ones ? o[kDP - 1] : l_run);
      __nv_bfloat16* dst = a.o_dense ? a.out + (((long long)bw * a.heads + h) * Nq + qi) * kDP
                                     : a.out + q_tok * a.ldo + a.o_off + h * kDP;
#pragma unroll
      for (int e = 0; e < kDP; e += 8)
        *reinterpret_cast<uint4*>(dst + e) =
            make_uint4(pack16(o[e] * inv, o[e + 1] * inv, fmt), pack16(o[e + 2] * inv, o[e + 3] * inv, fmt),
                       pack16(o[e + 4] * inv, o[e + 5] * inv, fmt), pack16(o[e + 6] * inv, o[e + 7] * inv, fmt));
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tcgen05_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

template <int KT, int KW, int VAR>
static int launch_attn_var(const AttnTcArgs& a, unsigned nblk, cudaStream_t st) {
  auto kern = attn_tc_kernel<KT, KW, VAR>;
  // the attribute is per device: a process that drives several GPUs configures each one once
  static bool configured[kMaxDevices] = {false};
  int dev = 0;
  GRL_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices || !configured[dev]) {
    GRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<KT>::TOTAL));
    if (dev >= 0 && dev < kMaxDevices) configured[dev] = true;
  }
  kern<<<nblk, kAttnThreads, AttnSmem<KT>::TOTAL, st>>>(a);
  GRL_LAUNCH_CHECK("attn_tc_kernel");
  return GRL_OK;
}

template <int KT, int KW>
static int launch_attn_one(const AttnTcArgs& a, unsigned nblk, cudaStream_t st) {
  switch ((a.fmt == FMT_BF16 ? 1 : 0) | (a.ones_col ? 2 : 0)) {
    case 0: return launch_attn_var<KT, KW, 0>(a, nblk, st);
    case 1: return launch_attn_var<KT, KW, 1>(a, nblk, st);
    case 2: return launch_attn_var<KT, KW, 2>(a, nblk, st);
    default: return launch_attn_var<KT, KW, 3>(a, nblk, st);
  }
}

int attn_variant(int set) {
  static int variant = -1;
  if (variant < 0) {
    // default 5: the persistent TMEM-resident kernel (attn2.cu) wherever the geometry has TMA boxes, this file's gather
    // kernel otherwise.  GRL_ATTN_SPLIT=0 forces the gather kernel (A/B runs, tools/attn_debug.py).
    const char* e = getenv("GRL_ATTN_SPLIT");
    variant = (e && atoi(e) == 0) ? 0 : 5;
  }
  const int prev = variant;
  if (set == 0 || set == 5) variant = set;
  return prev;
}

int launch_attn_tc(const AttnTcArgs& a, cudaStream_t st) {
  if (a.B == 0) return GRL_OK;
  int rc;
  if ((rc = check_grid(a.gq, "attn_tc(q grid)")) != GRL_OK) return rc;
  if ((rc = check_grid(a.gk, "attn_tc(k grid)")) != GRL_OK) return rc;
  GRL_REQUIRE(a.gq.H / a.gq.wh == a.gk.H / a.gk.wh && a.gq.W / a.gq.ww == a.gk.W / a.gk.ww,
              "attn_tc: query and key grids have different window counts");
  GRL_REQUIRE(a.heads >= 1 && a.heads <= 8, "attn_tc: heads=%d unsupported", a.heads);
  GRL_REQUIRE(a.rows_pad % 4 == 0 && a.rows_pad >= a.rows + 4, "attn_tc: bias table pitch %d too small for %d rows", a.rows_pad,
              a.rows);
  const int Nq = a.gq.wh * a.gq.ww;
  const long long nblk = (long long)a.B * (a.gq.H / a.gq.wh) * (a.gq.W / a.gq.ww) * a.heads * ceil_div(Nq, kQT);
  GRL_REQUIRE(nblk < (1ll << 31), "attn_tc: grid too large");
  if (attn_variant(-1) == 5) {  // persistent TMEM-resident kernel (attn2.cu) where the geometry has TMA boxes
    const int rc = launch_attn2(a, st);
    if (rc <= 0) return rc;
  }
  // 64 keys per tile, 3 CTAs / SM (32- and 128-key tiles were measured slower: profiles/r1_tc_path_final.md)
  switch (a.gk.ww) {
    case 8: return launch_attn_one<64, 8>(a, (unsigned)nblk, st);
    case 16: return launch_attn_one<64, 16>(a, (unsigned)nblk, st);
    case 32: return launch_attn_one<64, 32>(a, (unsigned)nblk, st);
    case 64: return launch_attn_one<64, 64>(a, (unsigned)nblk, st);
    case 128: return launch_attn_one<64, 128>(a, (unsigned)nblk, st);
    default: return launch_attn_one<64, 0>(a, (unsigned)nblk, st);
  }
}

}  // namespace tc
}  // namespace grl
