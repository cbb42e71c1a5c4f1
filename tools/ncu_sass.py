"""Per-SASS-instruction executed counts from an .ncu-rep (--import-source on): prints the instruction stream with
executed warp-instructions (in units of `unit`), active lanes and stall samples, so hot regions can be read directly.

    python tools/ncu_sass.py prof.ncu-rep [min_count_fraction_of_max]
"""
import csv, subprocess, sys

def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr)]

if __name__ == "__main__":
    recs = load(sys.argv[1])
    total = sum(int(r["Instructions Executed"]) for r in recs)
    print(f"total warp instructions {total:,}")
    cum = 0
    for i, r in enumerate(recs):
        n = int(r["Instructions Executed"]); cum += n
        lanes = float(r["Avg. Threads Executed"] or 0)
        print(f"{i:5d} {100*n/total:6.3f}% cum {100*cum/total:6.2f}% L{lanes:4.1f} s{int(r['# Samples']):6d}  {r['Source'].strip()}")
