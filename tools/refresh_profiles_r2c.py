"""Turn gpurun_out/r2c_*.ncu-rep (tools/gpu_r2c_ncu.sh) into the tracked summaries under profiles/: one ncu_summary +
ncu_lines text per kernel, r2c_traffic.json (dram bytes per launch, issue-slot / LSU-pipe utilisation: feeds bench.py's
roofline.traffic) and the launch list.   usage: python tools/refresh_profiles_r2c.py"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
REPS = {"blend_bwd_tr": "mobgs_blend_bwd", "blend_fwd": "mobgs_blend_fwd"}


def main():
    traffic = {"c4_1M_1080p_K7": {}, "issue_slot_utilisation": {}, "lsu_pipe_utilisation": {}, "shared_wavefronts": {},
               "source": "profiles/r2c_*_ncu.txt (ncu --set full --clock-control none, one launch each; traffic = "
                         "dram__bytes_read.sum + dram__bytes_write.sum)"}
    for stem, api in REPS.items():
        rep = os.path.join(OUT, f"r2c_{stem}.ncu-rep")
        if not os.path.exists(rep):
            continue
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
        lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "25"], capture_output=True, text=True).stdout
        with open(os.path.join(PROF, f"r2c_{stem}_ncu.txt"), "w") as f:
            f.write(txt + "\nper-source-line share of executed warp instructions / stall samples (tools/ncu_lines.py):\n" + lines)
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units, r = rows[0], rows[1], rows[2]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
        g = lambda k: float(r[hdr.index(k)].replace(",", ""))  # noqa: E731
        b = lambda k: g(k) * scale[units[hdr.index(k)]]  # noqa: E731
        traffic["c4_1M_1080p_K7"][api] = int(b("dram__bytes_read.sum") + b("dram__bytes_write.sum"))
        traffic["issue_slot_utilisation"][api] = round(g("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100, 4)
        traffic["lsu_pipe_utilisation"][api] = round(g("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") / 100, 4)
        for k in ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum"):
            if k in hdr:
                traffic["shared_wavefronts"].setdefault(api, {})[k] = g(k)
    with open(os.path.join(PROF, "r2c_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    print(json.dumps(traffic, indent=1))
    if os.path.exists(os.path.join(OUT, "r2c_launches.csv")):
        shutil.copy(os.path.join(OUT, "r2c_launches.csv"), os.path.join(PROF, "r2c_launches.csv"))


if __name__ == "__main__":
    main()
