"""Per-source-line share of executed warp instructions for one kernel of an .ncu-rep, by joining
`ncu --page source --csv` (per SASS address) with `nvdisasm -g` line info of the built library.
usage: python tools/ncu_lines.py report.ncu-rep kernel_regex mangled_prefix [top]"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, kre, mangled = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 35
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "harmonic_interpolation_b200", "libhgrid.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
i0 = [i for i, l in enumerate(sass) if l.startswith(".text." + mangled) and l.rstrip().endswith(":")][0]
cur, addr2line = None, {}
for l in sass[i0 + 1:]:
    if l.lstrip().startswith(".section"): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: addr2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hi = [i for i, r in enumerate(rows) if "Source" in r and "Address" in r][0]
hdr = rows[hi]; ai = hdr.index("Address"); ii = hdr.index("Instructions Executed")
base, per, tot = None, collections.Counter(), 0
for r in rows[hi + 1:]:
    try: n = int(r[ii].replace(",", "")); a = int(r[ai], 16)
    except (ValueError, IndexError): continue
    if base is None: base = a
    per[addr2line.get(a - base)] += n; tot += n
srcs = {}
print("total warp instructions:", tot)
for ln, n in per.most_common(top):
    text = ""
    if ln:
        p = os.path.join(root, "harmonic_interpolation_b200", "csrc", ln[0])
        if ln[0] not in srcs and os.path.isfile(p): srcs[ln[0]] = open(p).read().split("\n")
        if ln[0] in srcs and ln[1] - 1 < len(srcs[ln[0]]): text = srcs[ln[0]][ln[1] - 1].strip()[:105]
    print("%5.1f%%  %-22s %s" % (100.0 * n / tot, "%s:%d" % ln if ln else "?", text))
