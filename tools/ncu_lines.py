"""Per-source-line instruction / stall-sample totals of one kernel from an .ncu-rep captured with
--import-source on (cuda,sass correlated view).  usage: python tools/ncu_lines.py <rep> [top_n]"""
import csv
import subprocess
import sys


def main(path, top=45):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname = ""
    hdr = None
    tot_i = tot_s = 0
    lines = []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(r) < len(hdr) or not r[0]:
            continue
        try:
            n_i, n_s = int(r[ii]), int(r[si])
        except ValueError:
            continue
        lines.append((n_i, n_s, fname, r[0], r[1].strip()))
        tot_i += n_i
        tot_s += n_s
    print(f"total warp instructions {tot_i}, samples {tot_s}")
    for n_i, n_s, f, ln, src in sorted(lines, reverse=True)[:top]:
        print(f"{100 * n_i / tot_i:5.1f}% inst {100 * n_s / max(1, tot_s):5.1f}% smp  {f}:{ln:>4s}  {src[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
