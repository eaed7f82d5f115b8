"""Opcode histogram per kernel of a cubin/object file: python tools/sass_hist.py file.o [name-substring]"""
import collections, re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
pat = sys.argv[2] if len(sys.argv) > 2 else ""
name = None
hist = collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        hist[name][m.group(1).split(".")[0]] += 1
for n, h in hist.items():
    if pat in n:
        tot = sum(h.values())
        print(n, "total", tot)
        print("   ", ", ".join(f"{k} {v}" for k, v in h.most_common(18)))
