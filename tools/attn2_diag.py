"""Dev tool: differential timing of attn2.cu.  `build` (here): one libgrl_b200.so per GRL_A2_DIAG_* define (only attn2.cu
is recompiled, the other objects are reused) -> ab/lib<name>.so.  `run` (GPU box): times the three attention launches of
one GRL-Base block (tools/attn_debug.py) per variant.  Variants other than `base` compute garbage on purpose."""
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
STUB = "-DGRL_A2_DIAG_NOBIAS -DGRL_A2_DIAG_NOEXP -DGRL_A2_DIAG_NOLDTM -DGRL_A2_DIAG_NOSTTM -DGRL_A2_DIAG_NOMAX"
VARIANTS = {
    "base": "",
    # ingredients (results are garbage by construction)
    "noexp": "-DGRL_A2_DIAG_NOEXP",
    "nobias": "-DGRL_A2_DIAG_NOBIAS",
    "noldtm_nosttm": "-DGRL_A2_DIAG_NOLDTM -DGRL_A2_DIAG_NOSTTM",
    "stub": STUB,
    "stub_nomma": STUB + " -DGRL_A2_DIAG_NOPV -DGRL_A2_DIAG_NOQK",
    # structure (results stay correct)
    "nwg1": "-DGRL_A2_NWG=1",
    "nwg2": "-DGRL_A2_NWG=2",
    "multi_issuer": "-DGRL_A2_MULTI_ISSUER",
    "ilv": "-DGRL_A2_ILV",
    "multi_skew1200": "-DGRL_A2_MULTI_ISSUER -DGRL_A2_SKEW=1200",
}


def build():
    sys.path.insert(0, PKG)
    import build as b

    b.build()
    os.makedirs(AB, exist_ok=True)
    objdir = os.path.join(PKG, "build")
    objs = [os.path.join(objdir, os.path.basename(s)[:-3] + ".o") for s in b.sources()]
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    for name, defs in VARIANTS.items():
        out = os.path.join(AB, f"lib{name}.so")
        if not defs:
            shutil.copy(LIB, out)
            continue
        obj = os.path.join(AB, f"attn2_{name}.o")
        subprocess.run([nvcc] + b.NVCC_FLAGS + defs.split() + ["-c", os.path.join(PKG, "csrc", "attn2.cu"), "-o", obj],
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        subprocess.run([nvcc, "-shared", "-o", out] + [o if not o.endswith("attn2.o") else obj for o in objs], check=True)
        os.remove(obj)
        print("built", name, flush=True)


def run(batch):
    keep = LIB + ".production"
    shutil.copy(LIB, keep)
    try:
        for name in VARIANTS:
            src = os.path.join(AB, f"lib{name}.so")
            if not os.path.exists(src):
                continue
            shutil.copy(src, LIB)
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "attn_debug.py"), "--variants", "5", "--batch",
                                  str(batch), "--iters", "3"], capture_output=True, text=True, timeout=300)
            m = re.search(r"window ([0-9.]+) ms\s+stripe pass1 ([0-9.]+) ms\s+pass2 ([0-9.]+) ms", out.stdout)
            print(f"{name:14s} " + (f"window {m.group(1)}  pass1 {m.group(2)}  pass2 {m.group(3)} ms" if m else "FAILED " + out.stderr[-300:]), flush=True)
    finally:
        shutil.move(keep, LIB)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["build", "run"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--only", default="", help="comma-separated subset of the variants")
    a = ap.parse_args()
    if a.only:
        VARIANTS = {k: v for k, v in VARIANTS.items() if k in a.only.split(",")}
    build() if a.cmd == "build" else run(a.batch)
