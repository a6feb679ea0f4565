"""Dev tool: profiles/traffic.json (DRAM bytes of the three attention launches of one block, read by bench.py's roofline
object) from an `ncu --set full` capture of tools/attn_debug.py.   python tools/make_traffic.py <rep> <batch>"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(rep, batch):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, m):  # bytes / microseconds whatever unit ncu picked
        v, u = float(r[ix[m]]), units[ix[m]].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
        return v * scale.get(u, 1.0)

    names = ["window", "stripe pass 1", "stripe pass 2"]
    per, tot = {}, 0.0
    for n, r in zip(names, body[:3]):
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        per[f"{n} {r[ix['Kernel Name']].split('(')[0].split('::')[-1]}"] = {
            "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "us": round(val(r, "gpu__time_duration.sum"), 1)}
        tot += rd + wr
    tj = {
        "source": f"ncu --set full --clock-control none, profiles/{os.path.basename(rep)} (GRL-Base x4 SR 256x256, B={batch}, one "
                  "block's three attention launches, tools/attn_debug.py)",
        "attention_dram_bytes_per_image_per_block": int(tot / batch),
        "launches_per_block": 3,
        f"per_launch_B{batch}": per,
        "algorithmic_bytes_per_image_per_block": 113000000,
        "note": "q / k / v slots (16-bit, written by the QKV GEMM just before) + anchors + merged output + the stripe X1 buffer; "
                "every K / V tile is fetched once per 384 queries and re-reads hit L2",
    }
    json.dump(tj, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(tj, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
