"""Turn the .ncu-rep files a tools/gpu_profile_all.sh run left in gpurun_out/ into the tracked summaries
under profiles/ (ncu_summary text per kernel + r1_traffic.json, which bench.py reads for roofline.traffic
and the issue-slot figure).   usage: python tools/refresh_profiles.py [tag]   (tag defaults to r1)"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
REPS = {"blend_bwd": "mobgs_blend_bwd", "blend_fwd": "mobgs_blend_fwd", "tile_sort": None,
        "synth_project_bwd": None, "synth_project_fwd": None, "fused_adam": None, "ssim": None}


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[2:]


def main(tag):
    traffic = {"c4_1M_1080p_K7": {}, "issue_slot_utilisation": {}, "lsu_pipe_utilisation": {},
               "source": f"profiles/{tag}_*_ncu.txt (ncu --set full --clock-control none, one launch each; traffic = "
                         "dram__bytes_read.sum + dram__bytes_write.sum)"}
    for stem, api in REPS.items():
        rep = os.path.join(OUT, stem + ".ncu-rep")
        if not os.path.exists(rep):
            continue
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep],
                             capture_output=True, text=True).stdout
        lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "25"],
                               capture_output=True, text=True).stdout
        with open(os.path.join(PROF, f"{tag}_{stem}_ncu.txt"), "w") as f:
            f.write(txt + "\nper-source-line share of executed warp instructions / stall samples (tools/ncu_lines.py):\n" + lines)
        if api:
            hdr, rows = raw(rep)
            r = rows[0]
            g = lambda k: float(r[hdr.index(k)].replace(",", ""))
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
            units = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()[1]
            units = next(csv.reader([units]))
            rd = g("dram__bytes_read.sum") * scale[units[hdr.index("dram__bytes_read.sum")]]
            wr = g("dram__bytes_write.sum") * scale[units[hdr.index("dram__bytes_write.sum")]]
            traffic["c4_1M_1080p_K7"][api] = int(rd + wr)
            traffic["issue_slot_utilisation"][api] = round(g("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100, 4)
            traffic["lsu_pipe_utilisation"][api] = round(g("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") / 100, 4)
    with open(os.path.join(PROF, "r1_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r1")
