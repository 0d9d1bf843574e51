"""Per-source-line cost of one kernel: joins the SASS rows of an .ncu-rep (samples, instructions executed) with the
line table of the object file (nvdisasm --print-line-info) by instruction order.
   python tools/ncu_lines.py <file.ncu-rep> <file.o> <mangled-name-substring> [top_n]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, obj, key = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
i_src, i_s, i_e = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
sass = []
for r in rows[2:]:
    try:
        sass.append((r[i_src].strip(), int(r[i_s]), int(r[i_e])))
    except (ValueError, IndexError):
        pass

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines, cur, inside = [], None, False
for ln in dis.splitlines():
    if ln.startswith("\t.section\t.text."):
        inside = key in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", ln)
    if m:
        lines.append((cur, m.group(1).strip()))
print(f"{len(sass)} SASS rows in the report, {len(lines)} in the object")
n = min(len(sass), len(lines))
by_line = collections.defaultdict(lambda: [0, 0])
tot_s = sum(s for _, s, _ in sass) or 1
tot_e = sum(e for _, _, e in sass) or 1
mism = 0
for k in range(n):
    if sass[k][0].split()[0].split(".")[0] != lines[k][1].split()[0].split(".")[0]:
        mism += 1
    by_line[lines[k][0]][0] += sass[k][1]
    by_line[lines[k][0]][1] += sass[k][2]
print(f"opcode mismatches in the join: {mism}")
src_cache = {}
for (key_, (s, e)) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if key_:
        f = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", key_[0])
        if os.path.exists(f):
            src_cache.setdefault(f, open(f).read().splitlines())
            if key_[1] - 1 < len(src_cache[f]):
                text = src_cache[f][key_[1] - 1].strip()[:100]
    print(f"{s / tot_s * 100:5.1f}% samples {e / tot_e * 100:5.1f}% inst  {key_}  {text}")
