"""Algorithmic FLOP / score-element counts of the attention hot path: multiply-add = 2 FLOPs, true
head_dim (no padding), masks ignored.  Used by bench.py for the roofline line."""
from .geometry import stripe_info, to_2tuple


def attention_counts(cfg, x_size):
    """Returns dict(score_elems, f_qk, f_attn) per image for GRL(**cfg) on a padded (H, W) feature map."""
    H, W = x_size
    L = H * W
    C = cfg["embed_dim"]
    ws = to_2tuple(cfg["window_size"])
    df = cfg["anchor_window_down_factor"]
    score, f_qk = 0, 0
    for s, depth in enumerate(cfg["depths"]):
        hw, hs = cfg["num_heads_window"][s], cfg["num_heads_stripe"][s]
        dw, ds = (C // 2) // hw, (C // 2) // hs
        for i in range(depth):
            ss, sg = list(cfg["stripe_size"]), list(cfg["stripe_groups"])
            if i % 2 == 1:
                ss, sg = ss[::-1], sg[::-1]
            st, _ = stripe_info(ss, sg, True, x_size)
            n1 = st[0] * st[1]
            n2 = (st[0] // df) * (st[1] // df)
            s_win = L * ws[0] * ws[1] * hw
            s_str = 2 * (L // n1) * hs * n1 * n2
            score += s_win + s_str
            f_qk += 2 * dw * s_win + 2 * ds * s_str
    return dict(score_elems=score, f_qk=f_qk, f_attn=2 * f_qk)
