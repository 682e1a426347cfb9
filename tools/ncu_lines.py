#!/usr/bin/env python3
"""Aggregates the per-instruction samples of an .ncu-rep by CUDA source line (file:line from the -lineinfo of the matching object).
    python tools/ncu_lines.py <rep> <object.o> <kernel substring> [--top 40]"""
import csv, io, os, re, subprocess, sys, tempfile, collections

def main():
    rep, obj, kern = sys.argv[1:4]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
    line_of, sec, cur = {}, None, "?"
    for l in dis.splitlines():
        if l.startswith("\t.section") or ".text." in l and l.strip().startswith(".section"):
            sec = l
        m = re.search(r'//## File "([^"]*)", line (\d+)', l)
        if m: cur = f"{os.path.basename(m.group(1))}:{m.group(2)}"
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s", l)
        if m and sec and kern in sec: line_of[int(m.group(1), 16)] = cur
    src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
    h, rows = None, []
    for r in src:
        if len(r) > 5 and r[0] == "Address":
            if h: break
            h = {n: i for i, n in enumerate(r)}; continue
        if h and len(r) >= len(h) and r[h["# Samples"]].isdigit(): rows.append(r)
    base = int(rows[0][h["Address"]], 16)
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    agg = collections.defaultdict(lambda: collections.Counter())
    for r in rows:
        ln = line_of.get(int(r[h["Address"]], 16) - base, "?")
        a = agg[ln]
        a["samples"] += int(r[h["# Samples"]]); a["exec"] += int(r[h["Instructions Executed"]]); a["thr"] += int(r[h["Thread Instructions Executed"]]); a["n"] += 1
        for s in stalls: a[s] += int(r[h[s]] or 0)
    tot = sum(a["samples"] for a in agg.values()) or 1
    totex = sum(a["exec"] for a in agg.values()) or 1
    print(f"{'line':28s} {'samp%':>6s} {'exec%':>6s} {'lanes':>5s} {'sass':>5s}  top stalls")
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = sorted(((s[6:], a[s]) for s in stalls if a[s]), key=lambda kv: -kv[1])[:3]
        print(f"{ln:28s} {100 * a['samples'] / tot:6.2f} {100 * a['exec'] / totex:6.2f} {a['thr'] / max(1, a['exec']):5.1f} {a['n']:5d}  " + ", ".join(f"{k} {v}" for k, v in st))

if __name__ == "__main__":
    main()
