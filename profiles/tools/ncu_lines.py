"""Aggregate ncu warp-stall samples per CUDA source line.
usage: tools_ncu_lines.py <report.ncu-rep> <lib.so> <kernel-substring> [top]
Joins the SASS rows of `ncu --page source --csv` (in address order) with nvdisasm's line info for the same kernel."""
import csv, re, subprocess, sys, tempfile, os, collections, glob
rep, lib, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + os.environ.get("NCU_ARGS", "").split(), capture_output=True, text=True).stdout  # NCU_ARGS e.g. "-k regex:level -s 0 -c 1" selects one launch
rows = list(csv.reader(out.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
sass = rows[hdr_i + 1:]
# several launches in the selection: rows restart at a lower address; NCU_CHUNK picks one (default 0)
chunks, curc, prev = [], [], -1
for r in sass:
    try:
        ad = int(r[0], 16) if not r[0].isdigit() else int(r[0])
    except (ValueError, IndexError):
        continue
    if ad < prev:
        chunks.append(curc); curc = []
    curc.append(r); prev = ad
chunks.append(curc)
sass = chunks[int(os.environ.get("NCU_CHUNK", "0"))]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = glob.glob(tmp + "/*.cubin")[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
lines, cur, fn, infn = [], None, None, False
for l in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        fn = m.group(1); infn = kname in fn; continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if infn and re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
print(f"sass rows {len(sass)}  disasm instrs {len(lines)}")
agg = collections.Counter(); inst = collections.Counter()
for r, ln in zip(sass, lines):
    try:
        agg[ln] += int(r[col["# Samples"]]); inst[ln] += int(r[col["Instructions Executed"]])
    except (ValueError, KeyError):
        pass
tot = sum(agg.values()); toti = sum(inst.values())
src = {}
print(f"total samples {tot}, warp instructions {toti}")
for ln, c in agg.most_common(top):
    f, n = ln if ln else ("?", 0)
    if f not in src:
        p = [q for q in glob.glob("/root/repo/voxelis_b200/csrc/*") if os.path.basename(q) == f]
        src[f] = open(p[0]).read().split("\n") if p else []
    text = src[f][n - 1].strip()[:90] if src[f] and n - 1 < len(src[f]) else ""
    print(f"{100*c/tot:5.1f}% smp {100*inst[ln]/max(toti,1):5.1f}% ins  {f}:{n}  {text}")
