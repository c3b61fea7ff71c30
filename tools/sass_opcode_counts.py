"""Per-kernel counts of the SASS opcodes that prove the tensor-core / TMA / TMEM / bulk-copy paths (run in the build
container):  python tools/sass_opcode_counts.py > profiles/r02_sass_tensor_tma_opcode_counts.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "stylish_tts_b200", "csrc", "libstylish_b200.so")
OPS = ("UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "SYNCS", "MUFU", "FFMA", "HMMA")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::|sty::|<unnamed>::", "", name)
            cur = per.setdefault(name.split("(")[0][:90], collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur["total"] += 1
            for op in OPS:
                if m.group(1).startswith(op):
                    cur[op] += 1
    print("SASS opcode counts per kernel of libstylish_b200.so (cuobjdump -sass, sm_100a), round 2 final build")
    print("UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, UTMALDG / UTMASTG = cp.async.bulk.tensor load / store (TMA),")
    print("UBLKCP = cp.async.bulk (1-D bulk copy), LDTM / STTM = tcgen05.ld / st (TMEM), SYNCS = mbarrier ops")
    print(f"{'kernel':92s} " + " ".join(f"{o:>8s}" for o in ("total",) + OPS))
    tot = collections.Counter()
    for name, c in per.items():
        if not any(c[o] for o in ("UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM")):
            continue
        print(f"{name:92s} " + " ".join(f"{c[o]:8d}" for o in ("total",) + OPS))
        tot.update(c)
    print(f"{'SUM over the tensor-core / TMA kernels':92s} " + " ".join(f"{tot[o]:8d}" for o in ("total",) + OPS))
    print(f"kernels in the library: {len(per)}; with tcgen05 / TMA / bulk-copy opcodes: "
          f"{sum(1 for c in per.values() if any(c[o] for o in ('UTCHMMA', 'UTMALDG', 'UBLKCP')))}")


if __name__ == "__main__":
    main()
