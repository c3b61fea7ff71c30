"""SASS basic blocks of one `ncu --set full --import-source on` capture, ranked by executed warp instructions
(run in the build container):  python tools/ncu_source_blocks.py gpurun_out/prof_pw1dw.ncu-rep "title" > profiles/..."""
import collections
import csv
import io
import re
import subprocess
import sys


def main(path, title):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, IndexError):
            return 0.0

    blocks, cur = [], None
    for r in data:
        n = int(f(r, "Instructions Executed"))
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
        op = m.group(2) if m else "?"
        if cur is None or cur["count"] != n:
            cur = dict(count=n, len=0, instr=0.0, samples=0.0, ops=collections.Counter())
            blocks.append(cur)
        cur["len"] += 1
        cur["instr"] += n
        cur["samples"] += f(r, "# Samples")
        cur["ops"][op] += 1
    tot_i = sum(b["instr"] for b in blocks)
    tot_s = sum(b["samples"] for b in blocks)
    stalls = collections.Counter()
    for r in data:
        for k in ix:
            if k.startswith("stall_") and "Not Issued" not in k:
                stalls[k] += f(r, k)
    print(title)
    print(f"total warp instructions {int(tot_i)}  stall samples {int(tot_s)}")
    print("stall samples by reason:", ", ".join(f"{k[6:]} {int(v)}" for k, v in stalls.most_common(8)))
    for b in sorted(blocks, key=lambda b: -b["instr"])[:16]:
        print(f"{100 * b['instr'] / tot_i:5.1f}% instr {100 * b['samples'] / max(tot_s, 1):5.1f}% samples  "
              f"count={b['count']:8d} len={b['len']:4d} {dict(b['ops'].most_common(7))}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
