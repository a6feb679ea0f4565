"""Static evidence that needs no GPU: per-kernel resource usage (ptxas -v logs written by build.py) and SASS mnemonic counts
(cuobjdump -sass of the built objects) for the tcgen05 / TMA / TMEM instructions the profiling recipe names.

    python tools/sass_summary.py > profiles/r2_sass_summary.md
"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "grl-image-restoration_b200", "build")
MNEMONICS = ["UTCHMMA", "UTMALDG", "UBLKCP", "UTCBAR", "LDTM", "STTM", "LDGSTS", "MUFU.EX2", "FADD2", "SYNCS", "ELECT", "LDL", "STL"]


def demangle(names):
    p = subprocess.run(["c++filt"], input="\n".join(names), stdout=subprocess.PIPE, text=True)
    return dict(zip(names, p.stdout.split("\n")))


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\((CUtensorMap|tc::|grl::|float|int|long|unsigned|const|void|__half|__nv).*$", "", name)
    return name if len(name) < 90 else name[:87] + "..."


def resources():
    """ptxas -v: registers, spills, static shared memory per kernel."""
    out = {}
    for log in sorted(glob.glob(os.path.join(BUILD, "*.ptxas.log"))):
        cur = None
        for line in open(log):
            m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
            if m:
                cur = m.group(1)
                out[cur] = {"file": os.path.basename(log)[:-10]}
                continue
            if cur is None:
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m:
                out[cur].update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
            m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?", line)
            if m:
                out[cur].update(regs=int(m.group(1)), bars=int(m.group(2) or 0))
                m2 = re.search(r"(\d+) bytes smem", line)
                out[cur]["smem"] = int(m2.group(1)) if m2 else 0
    return out


def sass_counts():
    out = {}
    for obj in sorted(glob.glob(os.path.join(BUILD, "*.o"))):
        txt = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
        cur = None
        for line in txt.split("\n"):
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = m.group(1)
                out[cur] = collections.Counter()
                continue
            if cur is None:
                continue
            m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            op = m.group(1)
            out[cur]["_total"] += 1
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + "."):
                    out[cur][mn] += 1
    return out


def main():
    res, sass = resources(), sass_counts()
    names = demangle(sorted(set(res) | set(sass)))
    print("# Static kernel summary of libgrl_b200.so (sm_100a): `ptxas -v` resources and `cuobjdump -sass` mnemonic counts\n")
    print("Produced by `tools/sass_summary.py` from the objects `__graft_entry__.build()` compiles (no GPU involved).  UTCHMMA = "
          "`tcgen05.mma`, UTMALDG = `cp.async.bulk.tensor` (TMA), UBLKCP = `cp.async.bulk`, UTCBAR = `tcgen05.commit`, LDTM / STTM = "
          "`tcgen05.ld / st`, LDGSTS = `cp.async`, SYNCS = mbarrier ops, LDL / STL = local-memory (spill) accesses.\n")
    hdr = ["kernel", "file", "regs", "spill st/ld B", "static smem B", "SASS instr"] + MNEMONICS
    print("| " + " | ".join(hdr) + " |")
    print("|" + "---|" * len(hdr))
    rows = []
    for k in sorted(set(res) | set(sass), key=lambda k: (res.get(k, {}).get("file", ""), names[k])):
        r, c = res.get(k, {}), sass.get(k, collections.Counter())
        tc = c["UTCHMMA"] + c["UTMALDG"] + c["LDTM"]
        rows.append((tc == 0, [f"`{short(names[k])}`", r.get("file", ""), str(r.get("regs", "")),
                               f"{r.get('spill_st', 0)}/{r.get('spill_ld', 0)}", str(r.get("smem", "")), str(c["_total"])]
                     + [str(c[m]) if c[m] else "" for m in MNEMONICS]))
    for only_simt in (False, True):
        for simt, row in rows:
            if simt == only_simt:
                print("| " + " | ".join(row) + " |")
    n_tc = sum(1 for s, _ in rows if not s)
    print(f"\n{len(rows)} kernels, {n_tc} of them tensor-core / TMA kernels (listed first).")


if __name__ == "__main__":
    sys.exit(main())
