"""Per-source-line stall-sample summary of one kernel of an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py <file.ncu-rep> <kernel regex> [top N] [launch index]"""
import csv
import io
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    skip = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", "regex:" + rx,
                          "--launch-skip", str(skip), "--launch-count", "1"], capture_output=True, text=True).stdout
    hdr, agg, fname = None, {}, ""
    for r in csv.reader(io.StringIO(out)):
        if len(r) == 2 and r[0] == "File Name":
            fname = r[1].split("/")[-1]
            continue
        if len(r) > 6 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr) or r[2] != "-":   # keep the per-line rows (Address == "-")
            continue
        try:
            samp = int(r[hdr.index("# Samples")])
            inst = int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        stalls = {h[6:]: int(r[i]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit() and int(r[i])}
        key = (fname, int(r[0]))
        a = agg.setdefault(key, [0, 0, r[1].strip()[:100], {}])
        a[0] += samp
        a[1] += inst
        for k, v in stalls.items():
            a[3][k] = a[3].get(k, 0) + v
    tot = sum(a[0] for a in agg.values()) or 1
    print(f"# {rep} kernel~{rx}: {tot} samples")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        st = ",".join(f"{k}:{v}" for k, v in sorted(a[3].items(), key=lambda kv: -kv[1])[:3])
        print(f"{a[0]:6d} {100 * a[0] / tot:5.1f}% inst={a[1]:7d} {f}:{ln:<5d} {a[2]}  [{st}]")


if __name__ == "__main__":
    main()
