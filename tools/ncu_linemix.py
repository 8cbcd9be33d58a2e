"""Per-source-line instruction / stall-sample totals of one kernel: joins the SASS page of an ncu report (--set full
--import-source on) with nvdisasm's line table of the object the report was captured from (same build!).
usage: ncu_linemix.py REPORT KERNEL_REGEX OBJECT MANGLED_SUBSTRING [DIVISOR]"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, kern, obj, mangled = sys.argv[1:5]
div = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
line_of, cur, on = {}, None, False
for l in dis:
    if l.startswith("//---") and ".text." in l:
        on = mangled in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        inl = re.findall(r'inlined at "[^"]+", line (\d+)', m.group(3)) if "inlined" in m.group(3) else []
        cur = (int(m.group(2)), int(inl[-1]) if inl else None)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).split()[1] if m.group(2).startswith("@") else m.group(2).split()[0])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, cb = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cb = {"name": r[1], "rows": []}; blocks.append(cb); continue
    if cb is not None:
        cb["rows"].append(r)
best = None
for b in blocks:
    hdr = b["rows"][0]; iE = hdr.index("Instructions Executed"); iS = hdr.index("# Samples")
    body = [r for r in b["rows"][1:] if len(r) > iE and r[0].startswith("0x")]
    tot = sum(int(r[iE]) for r in body)
    if best is None or tot > best[0]:
        best = (tot, b["name"], body, iE, iS)
tot, name, body, iE, iS = best
base = int(body[0][0], 16)
by_line, by_outer = collections.Counter(), collections.Counter()
samp_line, samp_outer = collections.Counter(), collections.Counter()
ops = collections.defaultdict(collections.Counter)
miss = 0
for r in body:
    off = int(r[0], 16) - base
    n, s = int(r[iE]), int(r[iS])
    if off not in line_of:
        miss += n; continue
    (ln, outer), op = line_of[off]
    by_line[ln] += n; samp_line[ln] += s
    o = outer if outer is not None else ln
    by_outer[o] += n; samp_outer[o] += s
    ops[o][op.split(".")[0]] += n
src = open([l for l in dis if "//## File" in l][0].split('"')[1]).read().splitlines() if dis else []
print(f"{name}: {tot} warp-instr, {tot/div:.2f} per unit; unmatched {miss}")
tot_s = sum(samp_outer.values()) or 1
print("-- by outermost source line (inlined callees attributed to the call site) --")
for ln, n in by_outer.most_common(40):
    top = ",".join(f"{k}{v/div:.2f}" for k, v in ops[ln].most_common(4))
    text = src[ln - 1].strip()[:80] if 0 < ln <= len(src) else ""
    print(f"  L{ln:4d} {n/div:8.3f} /unit  {100*samp_outer[ln]/tot_s:5.1f}% samples  [{top}]  {text}")
