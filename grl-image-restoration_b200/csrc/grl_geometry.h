// grl_geometry.h -- closed-form index arithmetic shared by host code and every attention kernel.
//
// The reference builds int64 index tensors and -100 masks on the CPU (models/common/ops.py) and
// keeps ~1.5 GB of them as module buffers (models/networks/grl.py:386-429).  Here they are O(1)
// functions evaluated in registers; grl_*_host() in capi.cu expand them into tensors so the tests can
// check them bit-exactly against the reference's golden digests.
#pragma once
#include <stdint.h>

#include "../../include/grl_b200.h"

#if defined(__CUDACC__)
#define GRL_HD __host__ __device__ __forceinline__
#else
#define GRL_HD inline
#endif

namespace grl {

struct Tok {
  int r, c;  // coordinates in the ROLLED grid
  int y, x;  // coordinates in the un-rolled (memory) grid
  int ih, iw;  // coordinates inside the window
};

// Token n of window (wr, wc):  rolled position, then un-roll  R[r,c] = X[(r+s) mod H]  (torch.roll by -s,
// mixed_attn_block_efficient.py:141-143,:236-241) and window_partition's row-major order (ops.py:45-53).
GRL_HD Tok locate(const GrlGrid& g, int wr, int wc, int n) {
  Tok t;
  t.ih = n / g.ww;
  t.iw = n - t.ih * g.ww;
  t.r = wr * g.wh + t.ih;
  t.c = wc * g.ww + t.iw;
  t.y = t.r + g.sh;
  if (t.y >= g.H) t.y -= g.H;
  t.x = t.c + g.sw;
  if (t.x >= g.W) t.x -= g.W;
  return t;
}

// Region id of a rolled position: ops.py:76-99 (_fill_window) as a closed form.  Only (in)equality of two
// ids inside one window is ever used (ops.py:120-125,:148-155), and that matches the reference's
// sequential-slice construction for every shift, including the degenerate shift == 0.
GRL_HD int region_id(const GrlGrid& g, int r, int c) {
  int a = (r >= g.H - g.wh) + (g.sh > 0 && r >= g.H - g.sh);
  int b = (c >= g.W - g.ww) + (g.sw > 0 && c >= g.W - g.sw);
  return 3 * a + b;
}

// Relative-position index between a query token (qh,qw) of a (.., qww)-wide window and a key token (kh,kw)
// of a (kwh x kww) window: get_relative_position_index_simple + coords_diff_odd (ops.py:308-316,:352-375).
// window->anchor: q = window token, k = anchor; anchor->window: q = anchor, k = window token.
GRL_HD int rel_index(int qh, int qw, int kh, int kw, int qww, int kwh, int kww) {
  return (qh - kh + kwh - 1) * (qww + kww - 1) + (qw - kw + kww - 1);
}

GRL_HD int windows_per_image(const GrlGrid& g) { return (g.H / g.wh) * (g.W / g.ww); }

}  // namespace grl
