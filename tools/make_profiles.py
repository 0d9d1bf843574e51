"""Turn the captures of tools/profile_round.sh <tag> (gpurun_out/) into the tracked summaries under profiles/:
   profiles/<tag>_launches.csv, <tag>_launches_summary.txt, <tag>_ncu_full_summary.txt and profiles/roofline_traffic.json
   (dram__bytes_read.sum + dram__bytes_write.sum per launch of the march / brightness kernels, which bench.py reports
   as roofline.traffic).   python tools/make_profiles.py <tag> ["note"]"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
note = sys.argv[2] if len(sys.argv) > 2 else ""
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
cmd = "python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras"

shutil.copy(os.path.join(G, f"launches_{tag}.csv"), os.path.join(P, f"{tag}_launches.csv"))
summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), os.path.join(G, f"launches_{tag}.csv")],
                      capture_output=True, text=True).stdout
open(os.path.join(P, f"{tag}_launches_summary.txt"), "w").write(
    f"ncu --metrics gpu__time_duration.sum --clock-control none, {cmd}  ({tag}; {note})\n"
    "3 passes of the hot path in this command: 1 device-timed + 2 through the host-buffer (e2e) calls\n" + summ)

full, traffic = "", {}
if os.path.exists(os.path.join(P, "roofline_traffic.json")):
    traffic = json.load(open(os.path.join(P, "roofline_traffic.json")))
for k in ("brightness_kernel", "march_kernel", "traverse_kernel", "traverse_los", "gemm128_kernel", "kry_loop"):
    rep = os.path.join(G, f"prof_{k}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    full += subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))

    def bytes_of(name):
        v = float(d[name].replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[name]]
    traffic[k] = {"dram_bytes_per_launch": bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum"),
                  "dram_read": bytes_of("dram__bytes_read.sum"), "dram_write": bytes_of("dram__bytes_write.sum"),
                  "kernel": d.get("Kernel Name", "")[:80], "capture": f"{tag}: ncu --set full --clock-control none -k regex:{k} -c 1 {cmd}"}
open(os.path.join(P, f"{tag}_ncu_full_summary.txt"), "w").write(
    f"ncu --set full --clock-control none --import-source on, first launch of each kernel, {cmd}  ({tag}; {note})\n" + full)
json.dump(traffic, open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
print(summ)
print(json.dumps(traffic, indent=1))
