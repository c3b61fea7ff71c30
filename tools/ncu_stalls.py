"""Where the stall samples of one `ncu --set full --import-source on` capture sit, by SASS region:
    python tools/ncu_stalls.py file.ncu-rep [top]
Prints the instructions with the most samples (address, opcode, samples, top stall reason) and a coarse
split of samples by "role marker" (the nearest preceding UTCHMMA / LDTM / UTMALDG / SYNCS / STG / MUFU)."""
import collections, csv, io, subprocess, sys


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    def f(r, k):
        try: return float(r[ix[k]])
        except (ValueError, IndexError, KeyError): return 0.0
    stall_cols = [k for k in ix if k.startswith("stall_") and "Not Issued" not in k]
    tot = sum(f(r, "# Samples") for r in data)
    tot_i = sum(f(r, "Instructions Executed") for r in data)
    print(f"{path}: {len(data)} SASS lines, {int(tot)} samples, {int(tot_i)} warp instructions")
    agg = collections.Counter()
    for r in data:
        for k in stall_cols: agg[k] += f(r, k)
    print("stall reasons:", ", ".join(f"{k[6:]} {100*v/max(tot,1):.1f}%" for k, v in agg.most_common(8)))
    ranked = sorted(range(len(data)), key=lambda i: -f(data[i], "# Samples"))[:top]
    print("hottest instructions:")
    for i in sorted(ranked):
        r = data[i]
        why = max(stall_cols, key=lambda k: f(r, k))
        print(f"  line {i:5d} {100*f(r,'# Samples')/max(tot,1):5.2f}%  exec {int(f(r,'Instructions Executed')):9d}  {why[6:]:22s} {r[ix['Source']].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
