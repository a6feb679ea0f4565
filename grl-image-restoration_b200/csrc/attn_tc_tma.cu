// attn_tc_tma.cu -- EXPERIMENTAL variant of attn_tc.cu: identical softmax warps, but the producer warp fetches the Q / K /
// V tiles with TMA tensor copies of the (B, H, W, C) activation tensors instead of per-row cp.async gathers.
//
// Why (profiles/r1_tc_path_final.md): the busiest unit of the production kernel is the L1 data pipe (58 % of peak), and
// the per-row gathers are its largest client: each LDGSTS.128 warp instruction is 32 separate 16-byte requests (ncu: 32
// L1 tag requests and 32 shared-memory wavefronts per instruction, 8x the ideal), 26.7 M of each per window-attention
// launch against 23.6 M for the bias loads and 6.3 M for the P stores -- besides ~40 address instructions per gathered
// row in the one warp that also issues every MMA.  A TMA box moves the same bytes without touching the LSU pipe or the
// register file, and one lane issues it.
//
// Selected at run time by grl_tc_attn_variant(3) / GRL_ATTN_SPLIT=3; launches fall back to attn_tc.cu when a geometry
// does not satisfy the box conditions (see attn_tma_geometry).  Not yet measured on hardware: off by default.
#include <stdlib.h>

#include "attn_tc.cuh"
#include "grl_common.cuh"
#include "ops_f32.h"
#include "ops_tc.h"
#include "tc_common.cuh"

namespace grl {
namespace tc {

namespace {

constexpr int kAttnThreads = kQT + 32;

// Optional staging area for the bias rows of a key tile (template SB): two buffers after the barriers
template <int KT, bool SB>
struct AttnTmaSmem : AttnSmem<KT> {
  static constexpr int BIAS_BYTES = 7168;  // per buffer: up to (rows of a tile) x 4 shifted copies x padded row
  static constexpr int OFF_BIAS = AttnSmem<KT>::OFF_BAR + 128;
  static constexpr int TOTAL = SB ? OFF_BIAS + 2 * BIAS_BYTES + 1024 : AttnSmem<KT>::TOTAL;
};

struct AttnTmaGeom {
  int bw_q, bw_k;  // tokens per TMA box (power of two, divides gcd(window width, horizontal shift) and the tile sizes)
};

// SB: the producer stages the few bias-table rows a (query tile, key tile) pair needs in shared memory with 1-D bulk
// copies, and the softmax threads read their aligned float4 runs from there instead of from global memory / L1.
template <int KT, int KW, int VAR, bool SB>
__global__ void __launch_bounds__(kAttnThreads, (KT <= 32 ? 4 : KT <= 64 ? 3 : 2)) attn_tc_tma_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnTcArgs a, const AttnTmaGeom tg) {
  static_assert(KT == 32 || KT == 64 || KT == 128, "P tiles: 64-byte rows (SWIZZLE_64B) or 128-byte rows (SWIZZLE_128B)");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using S = AttnTmaSmem<KT, SB>;
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
  uint64_t* q_full = bar_s + 7;                                       // Q tile landed            (TMA complete_tx)
  uint64_t* k_full = bar_s + 8;                                       // [2] K_t landed
  uint64_t* v_full = bar_s + 10;                                      // [2] V_t landed
  uint64_t* bias_full = bar_s + 12;                                   // [2] staged bias rows of tile t landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s + 14);
  float* bias_s = reinterpret_cast<float*>(smem + S::OFF_BIAS);

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
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) mbar_init(&k_full[i], 1), mbar_init(&v_full[i], 1), mbar_init(&bias_full[i], 1);
    mbar_init_fence();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 4) tmem_alloc(tmem_slot, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const bool need_mask_cta = a.use_mask && (wr == nwh - 1 || wc == nww - 1);
  // ---- staged bias (SB): a (query row ih, key row kh) pair only needs table row ih - kh + KH - 1; a 128-query tile
  // spans few query rows and a 64-key tile few key rows, so a handful of rows of Wt floats covers the whole 128 x 64
  // bias tile.  Four copies, copy c shifted right by c entries, so every thread's reversed runs of 4 consecutive entries
  // are aligned 16-byte shared-memory loads (same trick as the global table).
  constexpr int KWc = KW > 0 ? KW : 4;
  const int RP = ((Wt + 8) + 3) & ~3;                                  // padded row length (floats)
  const int ih_a = (qt * kQT) / a.gq.ww;                                // query rows of this tile
  const int ih_b = min(Nq - 1, qt * kQT + kQT - 1) / a.gq.ww;
  const int krows_max = (KWc >= KT) ? 1 : KT / KWc;
  const bool use_sb = SB && (KW > 0) && ((ih_b - ih_a) + krows_max) * 4 * RP * 4 <= S::BIAS_BYTES && a.rows_pad >= a.rows + 16;
  const float* bias_hd = a.bias + (size_t)h * 4 * a.rows_pad;          // copy c of this head: bias_hd + c * rows_pad
  constexpr int fmt = (VAR & 1) ? FMT_BF16 : FMT_F16;

  if (warp == 4) {
    // =============================================================== producer + MMA issuer: TMA instead of gathers
    // A run of `bw` consecutive tokens of one window row is contiguous in the (B, H, W, C) tensor even after the roll
    // (bw divides gcd(window width, shift): the wrap-around of torch.roll never falls inside a run), so it is ONE
    // 4-D TMA box (32 channels x bw x 1 x 1) that lands as bw rows of 64 bytes, 64-byte swizzled by the copy engine --
    // the layout the UMMA descriptors of this kernel expect.  One lane issues one box; completion by transaction bytes.
    auto tma_rows = [&](const CUtensorMap* map, const GrlGrid& g, int coff, uint8_t* dst, int n0, int cnt, int bw,
                        uint64_t* bar) {
      if (lane == 0) mbar_expect_tx(bar, (uint32_t)cnt * 64u);
      __syncwarp();
      for (int s = lane; s * bw < cnt; s += 32) {
        const int n = n0 + s * bw;
        const int ih = n / g.ww, iw = n - ih * g.ww;
        int y = wr * g.wh + ih + g.sh;
        if (y >= g.H) y -= g.H;
        int x = wc * g.ww + iw + g.sw;
        if (x >= g.W) x -= g.W;
        tma_load_4d(dst + s * bw * 64, map, bar, coff, x, y, b);
      }
    };
    auto load_bias = [&](int tile) {  // rows needed by (this query tile) x (key tile `tile`); full tiles only
      const int k0 = tile * KT;
      if (!use_sb || k0 + KT > Nk) return;
      const int kh_a = k0 / KWc, kh_b = (k0 + KT - 1) / KWc;
      const int dmin = ih_a - kh_b;
      const int nrow = ((ih_b - ih_a) + (kh_b - kh_a) + 1) * 4;  // one bulk copy per (row, copy): RP floats each
      float* R = bias_s + (tile & 1) * (S::BIAS_BYTES / 4);
      uint64_t* bar = &bias_full[tile & 1];
      if (lane == 0) mbar_expect_tx(bar, (uint32_t)(nrow * RP * 4));
      __syncwarp();
      for (int i = lane; i < nrow; i += 32) {
        const int u = (dmin + (i >> 2) + a.gk.wh - 1) * Wt - (i & 3);  // table index that lands at R_c[0]
        const int g0 = (u + 3) & ~3;                                  // aligned start inside global copy g0 - u
        bulk_load_1d(R + i * RP, bias_hd + (size_t)(g0 - u) * a.rows_pad + g0, (uint32_t)(RP * 4), bar);
      }
    };
    auto load_q = [&]() { tma_rows(&tmQ, a.gq, a.q_off + h * kDP, Qs, qt * kQT, min(kQT, Nq - qt * kQT), tg.bw_q, q_full); };
    // koff / rid of a key tile: only tiles that take the generic bias path (ragged last tile, KW == 0) or the shift
    // mask read them
    auto load_meta = [&](int tile) {
      const int k0 = tile * KT, slot = tile % 3;
      if (need_mask_cta || KW == 0 || k0 + KT > Nk) {
        for (int r = lane; r < KT; r += 32) {
          const int kj = k0 + r;
          const Tok tk = locate(a.gk, wr, wc, kj < Nk ? kj : 0);
          koff_s[slot * KT + r] = tk.ih * Wt + tk.iw;
          krid_s[slot * KT + r] = region_id(a.gk, tk.r, tk.c);
        }
      }
      mbar_arrive(&meta_full[slot]);  // every tile arrives (the phase of a slot is tile / 3), needed or not
    };
    auto load_k = [&](int tile) {
      const int k0 = tile * KT;
      tma_rows(&tmK, a.gk, a.k_off + h * kDP, Ks + (tile & 1) * S::KV_BYTES, k0, min(KT, Nk - k0), tg.bw_k, &k_full[tile & 1]);
      load_meta(tile);
      load_bias(tile);  // same buffer discipline as K: every row is done with tile - 2's rows before s_free(tile - 2)
    };
    auto load_v = [&](int tile) {
      const int k0 = tile * KT, cnt = min(KT, Nk - k0);
      uint8_t* vd = Vs + (tile & 1) * S::KV_BYTES;
      if (a.v_dense) {  // (B_, heads, Nk, 32) rows: one 2-D box of KT rows (rows past this head's Nk are finite data of
                        // the next head or TMA zero fill, and meet P == 0)
        if (lane == 0) {
          mbar_expect_tx(&v_full[tile & 1], (uint32_t)KT * 64u);
          tma_load_2d(vd, &tmV, &v_full[tile & 1], 0, (int)(((long long)bw * a.heads + h) * Nk + k0));
        }
      } else {
        if (cnt < KT) {  // ragged tile: rows past Nk must be finite (they meet P == 0): zero them
          for (int i = lane; i < (KT - cnt) * 4; i += 32) *reinterpret_cast<uint4*>(vd + cnt * 64 + i * 16) = make_uint4(0, 0, 0, 0);
          fence_proxy_async_smem();
        }
        tma_rows(&tmV, a.gk, a.v_off + h * kDP, vd, k0, cnt, tg.bw_k, &v_full[tile & 1]);
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

    load_q();
    load_k(0);
    if (ntiles > 1) load_k(1);
    load_v(0);
    // Only lane 0 polls the barriers (it is also the MMA issuer); the other lanes park at __syncwarp, which orders their
    // later TMA issues after what lane 0 observed, instead of burning issue slots in 31 more spin loops.
    if (lane == 0) {
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tcgen05_fence_after();
      issue_qk(0);
    }
    __syncwarp();
    for (int t = 0; t < ntiles; ++t) {
      // ---- QK(t+1): K_{t+1} landed (buffer (t+1)&1, its ((t+1)>>1)-th use); S columns free once every row read S_t
      if (t + 1 < ntiles) {
        if (lane == 0) {
          mbar_wait(&k_full[(t + 1) & 1], ((t + 1) >> 1) & 1);
          mbar_wait(s_free, t & 1);
          tcgen05_fence_after();
          issue_qk(t + 1);
        }
        __syncwarp();
      }
      // K buffer t&1 and the koff / rid slot of tile t-1 are free (QK(t) completed before anyone could read S_t)
      if (t + 2 < ntiles) load_k(t + 2);
      // ---- PV(t): V_t landed; P_t written and O_{t-1} consumed by every row
      if (lane == 0) {
        mbar_wait(&v_full[t & 1], (t >> 1) & 1);
        mbar_wait(p_full, t & 1);
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
      __syncwarp();
      // V buffer (t+1)&1 held V_{t-1}; PV(t-1) is complete: every row waited for it before arriving on p_full(t)
      if (t + 1 < ntiles) load_v(t + 1);
    }
    if (lane == 0) mbar_wait(bar_o, (ntiles - 1) & 1);  // keep TMEM alive until the last MMA is done
    __syncwarp();
  } else {
    // =============================================================== softmax warps: thread = query row
    const int qi = qt * kQT + tid;
    const bool q_ok = qi < Nq;
    const Tok tq = locate(a.gq, wr, wc, q_ok ? qi : qt * kQT);  // padding rows mirror the tile's first row
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
      const bool sb_tile = full_tile && use_sb;
      const float* Rrow = nullptr;  // this thread's copy of row slot 0, positioned so that [.. - 4 qd] are its runs
      int dmin = 0;
      if (sb_tile) {
        mbar_wait(&bias_full[buf], (t >> 1) & 1);
        dmin = ih_a - (k0 + KT - 1) / KWc;
        const int cthr = (-tq.iw) & 3;  // (iw + KW - 1 - 3 + cthr) % 4 == 0 because KW % 4 == 0
        Rrow = bias_s + buf * (S::BIAS_BYTES / 4) + cthr * RP + (tq.iw + KWc - 1 - 3 + cthr);
      }

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
            const float4* bp;
            if (sb_tile) {  // staged rows: row slot (ih - kh) - dmin, column iw - kw + KW - 1 (shifted by the copy)
              bp = reinterpret_cast<const float4*>(Rrow + ((tq.ih - kj / KWS) - dmin) * 4 * RP - (kj % KWS));
            } else {        // straight from the 4-copy table in global memory
              const int s0 = base_i - ((kj / KWS) * Wt + (kj % KWS)) - 3;  // table index of key kj + 3
              const int cpy = (-s0) & 3;
              bp = reinterpret_cast<const float4*>(bias_h + (size_t)cpy * a.rows_pad + (s0 + cpy));
            }
#pragma unroll
            for (int qd = 0; qd < RW / 4; ++qd) {
              const float4 bb = sb_tile ? bp[-qd] : __ldg(bp - qd);
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
          p[e] = ex2(lg[c * 8 + e] - m_new);
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
          *reinterpret_cast<uint4*>(Pt + sub * (kQT * 128) + tid * 128 + ((cc ^ (tid & 7)) << 4)) = pk;
        }
      }
      l_run = l_run * corr + ((ps[0] + ps[1]) + (ps[2] + ps[3]));

      // ---- fold in the previous tile's P V (it has had a whole softmax to finish).  o is kept relative to the running
      // max at which the last folded O was computed, so the fold is one FFMA per element: o <- o * corr_{t-1} + O_{t-1}
      if (t > 0) {
        mbar_wait(bar_o, (t - 1) & 1);
        tcgen05_fence_after();
        {
          uint32_t v[32];
          tmem_ld32(trow + KT, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < kDP; ++e) o[e] = fmaf(o[e], corr_prev, __uint_as_float(v[e]));
        }
      }
      corr_prev = corr;
      tcgen05_fence_before();
      fence_proxy_async_smem();  // P_t (generic-proxy stores) -> visible to the tensor core
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


// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
// Tokens per box for a window of width ww rolled by sw: the largest power of two <= 64 dividing gcd(ww, sw) (ww if the
// grid is not rolled horizontally).  0 = no usable box (odd widths, or runs shorter than 8 tokens = 512 bytes).
int box_tokens_impl(const GrlGrid& g) {
  int d = g.ww;
  if (g.sw > 0) {
    int x = g.ww, y = g.sw;
    while (y) {
      const int t = x % y;
      x = y, y = t;
    }
    d = x;
  }
  int bw = 64;
  while (bw > 1 && d % bw) bw >>= 1;
  return bw >= 8 ? bw : 0;  // >= 512 bytes per box: destinations stay aligned to the 64-byte-swizzle repeat
}

int make_token_map(CUtensorMap* m, const void* base, long long ld, const GrlGrid& g, int B, int bw) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(GRL_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t dims[4] = {(cuuint64_t)ld, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)B};
  const cuuint64_t str[3] = {(cuuint64_t)ld * 2, (cuuint64_t)g.W * ld * 2, (cuuint64_t)g.H * g.W * ld * 2};
  const cuuint32_t box[4] = {(cuuint32_t)kDP, (cuuint32_t)bw, 1, 1};
  const cuuint32_t ones[4] = {1, 1, 1, 1};
  const CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(base), dims, str, box, ones,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return fail(GRL_ERR_CUDA, "cuTensorMapEncodeTiled (attention tokens) failed with CUresult %d", (int)rc);
  return GRL_OK;
}

int make_dense_map(CUtensorMap* m, const void* base, long long rows, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(GRL_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t dims[2] = {(cuuint64_t)kDP, (cuuint64_t)rows};
  const cuuint64_t str[1] = {(cuuint64_t)kDP * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kDP, (cuuint32_t)box_rows};
  const cuuint32_t ones[2] = {1, 1};
  const CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, str, box, ones,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return fail(GRL_ERR_CUDA, "cuTensorMapEncodeTiled (dense V) failed with CUresult %d", (int)rc);
  return GRL_OK;
}

template <int KT, int KW, int VAR, bool SB>
int launch_tma_var(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnTcArgs& a,
                   const AttnTmaGeom& tg, unsigned nblk, cudaStream_t st) {
  auto kern = attn_tc_tma_kernel<KT, KW, VAR, SB>;
  static bool configured = false;
  if (!configured) {
    GRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnTmaSmem<KT, SB>::TOTAL));
    configured = true;
  }
  kern<<<nblk, kAttnThreads, AttnTmaSmem<KT, SB>::TOTAL, st>>>(tq, tk, tv, a, tg);
  GRL_LAUNCH_CHECK("attn_tc_tma_kernel");
  return GRL_OK;
}

template <int KW, bool SB>
int launch_tma_kw(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnTcArgs& a,
                  const AttnTmaGeom& tg, unsigned nblk, cudaStream_t st) {
  switch ((a.fmt == FMT_BF16 ? 1 : 0) | (a.ones_col ? 2 : 0)) {
    case 0: return launch_tma_var<64, KW, 0, SB>(tq, tk, tv, a, tg, nblk, st);
    case 1: return launch_tma_var<64, KW, 1, SB>(tq, tk, tv, a, tg, nblk, st);
    case 2: return launch_tma_var<64, KW, 2, SB>(tq, tk, tv, a, tg, nblk, st);
    default: return launch_tma_var<64, KW, 3, SB>(tq, tk, tv, a, tg, nblk, st);
  }
}

template <bool SB>
int launch_tma_sb(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnTcArgs& a,
                  const AttnTmaGeom& tg, unsigned nblk, cudaStream_t st) {
  switch (a.gk.ww) {
    case 8: return launch_tma_kw<8, SB>(tq, tk, tv, a, tg, nblk, st);
    case 16: return launch_tma_kw<16, SB>(tq, tk, tv, a, tg, nblk, st);
    case 32: return launch_tma_kw<32, SB>(tq, tk, tv, a, tg, nblk, st);
    case 64: return launch_tma_kw<64, SB>(tq, tk, tv, a, tg, nblk, st);
    case 128: return launch_tma_kw<128, SB>(tq, tk, tv, a, tg, nblk, st);
    default: return launch_tma_kw<0, SB>(tq, tk, tv, a, tg, nblk, st);
  }
}

}  // namespace

int attn_tma_box_tokens(const GrlGrid& g) { return box_tokens_impl(g); }

// Returns GRL_OK after launching, a negative error, or +1 when this geometry cannot be expressed as TMA boxes (the
// caller then launches the gather kernel).  Arguments already validated by launch_attn_tc.
int launch_attn_tc_tma(const AttnTcArgs& a, unsigned nblk, bool staged_bias, cudaStream_t st) {
  AttnTmaGeom tg;
  tg.bw_q = box_tokens_impl(a.gq);
  tg.bw_k = box_tokens_impl(a.gk);
  if (tg.bw_q == 0 || tg.bw_k == 0) return 1;
  if (a.gq.W < tg.bw_q || a.gk.W < tg.bw_k) return 1;
  // 16-byte alignment of every box origin / pitch (checked by capi for pitches and offsets; bases come from torch)
  if ((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) | reinterpret_cast<uintptr_t>(a.v)) & 15) return 1;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_token_map(&tq, a.q, a.ldq, a.gq, a.B, tg.bw_q)) != GRL_OK) return rc;
  if ((rc = make_token_map(&tk, a.k, a.ldk, a.gk, a.B, tg.bw_k)) != GRL_OK) return rc;
  if (a.v_dense) {
    const long long rows = (long long)a.B * (a.gk.H / a.gk.wh) * (a.gk.W / a.gk.ww) * a.heads * a.gk.wh * a.gk.ww;
    if (rows < 64) return 1;  // the dense-V box is 64 rows: keep it inside the tensor
    if ((rc = make_dense_map(&tv, a.v, rows, 64)) != GRL_OK) return rc;
  } else {
    if ((rc = make_token_map(&tv, a.v, a.ldv, a.gk, a.B, tg.bw_k)) != GRL_OK) return rc;
  }
  return staged_bias ? launch_tma_sb<true>(tq, tk, tv, a, tg, nblk, st) : launch_tma_sb<false>(tq, tk, tv, a, tg, nblk, st);
}

}  // namespace tc
}  // namespace grl
