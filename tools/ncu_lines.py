"""Per-source-line executed-instruction profile from an .ncu-rep captured with --import-source on
(kernels compiled with -lineinfo).  Shows where the warp instructions go.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    # the export is a sequence of per-file tables; each starts with "File Path", then "Function Name", then a header row
    per_line = defaultdict(lambda: [0, 0, 0])      # (file,line) -> [warp inst, thread inst, samples]
    src_text = {}
    cur_file, header = None, None
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == "File Path":
            cur_file, header = row[1], None
            continue
        if row[0] == "Function Name":
            continue
        if row[0] == "Line No":
            header = row
            continue
        if header is None:
            continue
        rec = dict(zip(header, row))
        try:
            line = int(rec["Line No"])
            inst = int(float(rec.get("Instructions Executed") or 0))
            tinst = int(float(rec.get("Thread Instructions Executed") or 0))
            samples = int(float(rec.get("# Samples") or 0))
        except ValueError:
            continue
        key = (cur_file.split("/")[-1], line)
        per_line[key][0] += inst
        per_line[key][1] += tinst
        per_line[key][2] += samples
        src_text.setdefault(key, row[1][:110])
    # every SASS instruction appears once per table it is attributed to; use the CUDA-C view only
    total = sum(v[0] for v in per_line.values())
    print(f"total warp instructions attributed: {total:,}")
    print(f"{'file:line':28s} {'warp inst':>14s} {'%':>6s} {'thr/inst':>8s} {'samples':>8s}  source")
    for key, v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        eff = v[1] / v[0] if v[0] else 0
        print(f"{key[0] + ':' + str(key[1]):28s} {v[0]:14,d} {100 * v[0] / max(total, 1):6.2f} {eff:8.1f} {v[2]:8d}  {src_text[key].strip()}")


if __name__ == "__main__":
    main()
