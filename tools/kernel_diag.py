"""Dev tool: differential timing of the attention and GEMM kernels (what does each ingredient cost?).

Two steps, because nvcc is in the build container and the GPU is not:

  python tools/kernel_diag.py build     # here: one libgrl_b200.so per GRL_*_DIAG_* define -> ab/lib<name>.so
  gpurun -- python tools/kernel_diag.py run [--batch 8]
                                        # on the B200: times one GRL-Base x4 forward per variant with CUDA events

`run` swaps each library into place, launches tools/time_model.py in a fresh process and restores the production
library at the end.  Variants other than `base` compute garbage on purpose (attn_tc.cu GRL_DIAG_*, gemm_tc.cu
GRL_GDIAG_* macros): only their
TIME means anything.  ab/ is scratch (git-ignored) but travels to the GPU box -- delete it when done (16 MB / variant).
Caveat: with the ones-column denominators nothing else consumes P, so `nopstore` also removes the exp2 (dead code).
"""
import argparse
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "grl-image-restoration_b200")
LIB = os.path.join(PKG, "libgrl_b200.so")
AB = os.path.join(ROOT, "ab")
VARIANTS = {
    "base": "",
    "nobias": "-DGRL_ATTN_DIAG_NOBIAS",
    "noexp": "-DGRL_ATTN_DIAG_NOEXP",
    "nopstore": "-DGRL_ATTN_DIAG_NOPSTORE",
    "nofold": "-DGRL_ATTN_DIAG_NOFOLD",
    "nopfence": "-DGRL_ATTN_DIAG_NOPFENCE",
    "nogather": "-DGRL_ATTN_DIAG_NOGATHER",
    "nobias_noexp": "-DGRL_ATTN_DIAG_NOBIAS -DGRL_ATTN_DIAG_NOEXP",
    "softmax_stub": "-DGRL_ATTN_DIAG_NOBIAS -DGRL_ATTN_DIAG_NOEXP -DGRL_ATTN_DIAG_NOPSTORE -DGRL_ATTN_DIAG_NOFOLD",
    "quadgather": "-DGRL_ATTN_QUAD_GATHER",  # NOT a stub: same results, each LDGSTS warp instruction copies 8 whole rows
    "gemm_nores": "-DGRL_GEMM_DIAG_NORES",
    "gemm_nocab": "-DGRL_GEMM_DIAG_NOCAB",
    "gemm_nost32": "-DGRL_GEMM_DIAG_NOST32",
    "gemm_nost16": "-DGRL_GEMM_DIAG_NOST16",
    "gemm_noepi_io": "-DGRL_GEMM_DIAG_NORES -DGRL_GEMM_DIAG_NOCAB -DGRL_GEMM_DIAG_NOST32 -DGRL_GEMM_DIAG_NOST16",
}


def build(only=None):
    sys.path.insert(0, PKG)
    import build as b

    os.makedirs(AB, exist_ok=True)
    for name, defs in VARIANTS.items():
        if only and name not in only and name != "base":
            continue
        os.environ["GRL_NVCC_DEFINES"] = defs
        b.build(force=True)
        shutil.copy(LIB, os.path.join(AB, f"lib{name}.so"))
        print("built", name, defs)
    os.environ["GRL_NVCC_DEFINES"] = ""
    b.build(force=True)  # leave the production library in place


def run(batch, precision):
    keep = LIB + ".production"
    shutil.copy(LIB, keep)
    rows = []
    try:
        for name in VARIANTS:
            src = os.path.join(AB, f"lib{name}.so")
            if not os.path.exists(src):
                print("missing", src, "- run `kernel_diag.py build` first")
                continue
            shutil.copy(src, LIB)
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "time_model.py"), "--variant", "base", "--size",
                                  "256", "--batch", str(batch), "--precision", precision, "--style", "init"],
                                 capture_output=True, text=True, timeout=300)
            m = re.search(r"([0-9.]+) ms/forward", out.stdout)
            rows.append((name, float(m.group(1)) if m else float("nan")))
            print(f"{name:14s} {rows[-1][1]:8.1f} ms/forward", flush=True)
    finally:
        shutil.move(keep, LIB)
    base = dict(rows).get("base")
    if base:
        print("\n| removed ingredient | ms / forward | saved |\n|---|---:|---:|")
        for name, ms in rows:
            print(f"| {name} | {ms:.1f} | {base - ms:+.1f} ms ({100 * (base - ms) / base:+.1f} %) |")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["build", "run"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--only", default="", help="comma list of variants to build (base is always built)")
    a = ap.parse_args()
    build([v for v in a.only.split(",") if v]) if a.cmd == "build" else run(a.batch, a.precision)
