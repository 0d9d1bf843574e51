"""Opcode mix of one kernel from an .ncu-rep captured with --import-source on: python tools/ncu_opmix.py file.ncu-rep"""
import collections
import csv
import subprocess
import sys


def main(path, top=22):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[1]
    ia, ii, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    tot = 0
    byop, samp = collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) <= ii:
            continue
        n = int(r[ii])
        tot += n
        f = r[ia].split()
        op = (f[1] if f[0].startswith("@") else f[0]).split(".")[0]
        byop[op] += n
        samp[op] += int(r[isamp])
    print("kernel:", rows[0][1][:100])
    print("warp instructions executed:", tot)
    for op, n in byop.most_common(top):
        print(f"  {op:10s} {n / tot * 100:6.2f}%   stall samples {samp[op]}")


if __name__ == "__main__":
    main(sys.argv[1])
