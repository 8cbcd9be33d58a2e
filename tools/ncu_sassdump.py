"""Dump the SASS of one kernel from an ncu report with executed counts per unit (and stall samples):
usage: ncu_sassdump.py REPORT KERNEL_REGEX DIVISOR > out.txt"""
import csv, subprocess, sys
rep, kern, div = sys.argv[1], sys.argv[2], float(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
for r in rows:
    if r and r[0] == "Address":
        hdr = r; iE = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); iSrc = hdr.index("Source"); continue
    if hdr and r and r[0].startswith("0x") and len(r) > iE:
        print(f"{int(r[0],16)&0xffff:5x} {int(r[iE])/div:8.2f} {int(r[iS]):6d}  {r[iSrc]}")
