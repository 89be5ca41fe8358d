"""Warp-stall samples of one kernel by CUDA source line: joins an `ncu --page source --csv` export (per SASS address; several
kernels may follow each other in the file) with the `nvdisasm -g` line table of the built library.
usage: python tools/ncu_stalls.py source.csv section_index mangled_prefix [top]"""
import collections, csv, os, re, subprocess, sys, tempfile
src_csv, sec, mangled = sys.argv[1], int(sys.argv[2]), sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "harmonic_interpolation_b200", "libhgrid.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
addr2line = {}
for cub in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
    sass = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
    hits = [i for i, l in enumerate(sass) if l.startswith(".text." + mangled) and l.rstrip().endswith(":")]
    if not hits: continue
    cur = None
    for l in sass[hits[0] + 1:]:
        if l.lstrip().startswith(".section"): break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m: addr2line[int(m.group(1), 16)] = cur
    break
rows = list(csv.reader(open(src_csv)))
heads = [i for i, r in enumerate(rows) if "Source" in r and "Address" in r]
hi = heads[sec]; end = heads[sec + 1] - 1 if sec + 1 < len(heads) else len(rows)
hdr = rows[hi]; ai = hdr.index("Address"); si = hdr.index("Warp Stall Sampling (All Samples)"); ii = hdr.index("Instructions Executed")
reasons = [(j, h) for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
per, ex, why, tot, base = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter), 0, None
whole = collections.Counter()
for r in rows[hi + 1:end]:
    try: a = int(r[ai], 16); s = int(r[si] or 0); n = int(r[ii] or 0)
    except (ValueError, IndexError): continue
    if base is None: base = a
    ln = addr2line.get(a - base)
    per[ln] += s; ex[ln] += n; tot += s
    for j, h in reasons:
        v = int(r[j] or 0)
        why[ln][h] += v; whole[h] += v
print("stall samples:", tot, " by reason:", ", ".join("%s %d" % (h[6:], v) for h, v in whole.most_common(6)))
srcs = {}
for ln, s in per.most_common(top):
    text = ""
    if ln:
        p = os.path.join(root, "harmonic_interpolation_b200", "csrc", ln[0])
        if ln[0] not in srcs and os.path.isfile(p): srcs[ln[0]] = open(p).read().split("\n")
        if ln[0] in srcs and ln[1] - 1 < len(srcs[ln[0]]): text = srcs[ln[0]][ln[1] - 1].strip()[:90]
    w = ",".join("%s %d" % (h[6:], v) for h, v in why[ln].most_common(2))
    print("%5.1f%% exec %9d %-24s [%s] %s" % (100.0 * s / max(tot, 1), ex[ln], "%s:%d" % ln if ln else "?", w, text))
