"""Dev tool: markdown summary of an `ncu --set full --import-source on` report (read HERE with the ncu CLI):
per-launch duration, pipe utilisation, DRAM bytes, warp-stall breakdown, executed-instruction mix by opcode and the
hottest SASS lines of the first launch.   python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md"""
import csv
import io
import subprocess
import sys
from collections import Counter

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("smsp__warps_active.avg.per_cycle_active", "warps active / scheduler"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def ncu(rep, *args):
    out = subprocess.run(["ncu", "-i", rep, "--csv"] + list(args), capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep):
    rows = ncu(rep, "--page", "raw")
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep.split('/')[-1]}\n")
    print("| launch | " + " | ".join(n for m, n in METRICS if m in ix) + " |")
    print("|---|" + "---:|" * sum(m in ix for m, _ in METRICS))
    for r in body:
        name = r[ix["Kernel Name"]].split("(")[0][-48:]
        cells = []
        for m, _ in METRICS:
            if m in ix:
                v = r[ix[m]]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {units[ix[m]]}".strip())
        print(f"| `{name}` | " + " | ".join(cells) + " |")
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    print("\nWarp stalls per issued instruction (first launch): " + ", ".join(
        f"{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} {float(body[0][ix[h]]):.2f}"
        for h in sorted(stalls, key=lambda h: -float(body[0][ix[h]]))[:8]))

    src = ncu(rep, "--page", "source", "--print-source", "sass", "--kernel-id", ":::1")
    shdr = next(r for r in src if r and r[0] == "Address")
    six = {h: i for i, h in enumerate(shdr)}
    lines, seen = [], set()
    for r in src:
        if len(r) == len(shdr) and r[0].startswith("0x"):
            if r[0] in seen:
                break
            seen.add(r[0])
            lines.append(r)
    S, E = six["# Samples"], six["Instructions Executed"]
    tot_s = sum(int(r[S]) for r in lines) or 1
    tot_e = sum(int(r[E]) for r in lines) or 1
    mix, smp = Counter(), Counter()
    for r in lines:
        ops = [o for o in r[1].split() if not o.startswith("@")]
        op = ops[0] if ops else ""
        op = op if op.startswith("MUFU") else op.split(".")[0]
        mix[op] += int(r[E])
        smp[op] += int(r[S])
    print("\n| opcode | executed (warp instr) | share | stall samples | share |\n|---|---:|---:|---:|---:|")
    for op, n in mix.most_common(18):
        print(f"| {op} | {n} | {100 * n / tot_e:.1f} % | {smp[op]} | {100 * smp[op] / tot_s:.1f} % |")
    print("\nHottest SASS lines (stall samples):\n\n```")
    for r in sorted(lines, key=lambda r: -int(r[S]))[:14]:
        print(f"{int(r[S]):6d} samples  {int(r[E]):9d} exec  {' '.join(r[1].split())[:100]}")
    print("```")


if __name__ == "__main__":
    main(sys.argv[1])
