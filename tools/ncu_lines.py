"""Per-source-line hot spots from an .ncu-rep captured with --import-source on (here, no GPU):
    python tools/ncu_lines.py <file.ncu-rep> <cubin> <kernel substring> [top N]
Joins ncu's SASS-level source page (samples / instructions per instruction) with nvdisasm's line info of the same
kernel by instruction ordinal, then aggregates by file:line (inlined call chains are attributed to the innermost line)."""
import csv, io, re, subprocess, sys
rep, cubin, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
h = rows[hi]
S, I, T, SRC = h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("Source")
inst = [(int(r[S]), int(r[I]), int(r[T]), r[SRC].strip()) for r in rows[hi + 1:] if len(r) > T and r[S].isdigit()]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# locate the function
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
lines, cur = [], "?"
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = f"{m.group(1).split('/')[-1]}:{m.group(2)}"
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append((cur, l.split("*/", 1)[1].strip()))
if len(lines) != len(inst):
    print(f"warning: {len(lines)} instructions in the cubin vs {len(inst)} in the report (different build?)")
n = min(len(lines), len(inst))
agg = {}
for k in range(n):
    a = agg.setdefault(lines[k][0], [0, 0, 0])
    a[0] += inst[k][0]; a[1] += inst[k][1]; a[2] += inst[k][2]
ts, ti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print(f"{n} instructions, {ts} samples, {ti} warp instructions")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * a[0] / ts:5.1f}% samples {100 * a[1] / ti:5.1f}% inst  lanes {a[2] / max(a[1], 1):4.1f}  {key}")
