"""Per-PHASE breakdown of k_fast_queue from an .ncu-rep captured with --import-source on (here, no GPU):
    python tools/ncu_phases.py <file.ncu-rep> <cubin> <kernel substring>
Like tools/ncu_lines.py (ncu's per-instruction samples joined with nvdisasm's line info by instruction ordinal), but the
innermost source line of every instruction is mapped to a phase of the path step through line ranges of tpt_fast.cu that
are looked up by marker text, so the table survives edits of the file."""
import csv, io, os, re, subprocess, sys
rep, cubin, kern = sys.argv[1], sys.argv[2], sys.argv[3]
src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "toypathtracer_b200", "csrc", "tpt_fast.cu")).read().splitlines()
def line_of(marker, start=0):
    return next(i + 1 for i, l in enumerate(src) if i >= start and marker in l)
L_sweep0 = line_of("f2_bcast(float x)")        # the packed-math helpers are pass 1's instructions
L_k2 = line_of("struct FastHitterK2")
L_k2_hit = line_of("int hit(const SceneView&", L_k2)
L_k2_pass2 = line_of("while (cand)", L_k2)
L_k2c = line_of("struct FastHitterK2C")
L_k2c_pass2 = line_of("while (cand)", L_k2c)
L_pairs = line_of("void build_sph_pairs_from_r2(")
L_step = line_of("bool path_step(")
L_light = line_of("if (wantLight)", L_step)
L_queue = line_of("k_fast_queue(DrawParams p")
L_regen = line_of("// ---- regeneration", L_queue)
L_call = line_of("const bool finished = KFORM == 3", L_queue)
L_end = line_of("// ---- variant 8")
L_gen = line_of("void generate_slab_rays(")
L_qpath = line_of("struct QPath")
def phase(f, ln):
    if f == "tpt_fast.cu":
        if L_sweep0 <= ln < L_k2: return "pass 1 (packed sweep)"
        if L_k2 <= ln < L_k2_hit: return "pass 2 (candidates)"              # the hitter's LDS.128 helper
        if L_k2_hit <= ln < L_k2_pass2: return "sweep setup (ray constants)"
        if L_k2_pass2 <= ln < L_k2c: return "pass 2 (candidates)"
        if L_k2c <= ln < L_k2c_pass2: return "sweep setup (ray constants)"
        if L_k2c_pass2 <= ln < L_pairs: return "pass 2 (candidates)"
        if L_step <= ln < L_light: return "shade: hit point, material branches"
        if L_light <= ln < L_queue: return "shade: light sample / continue"
        if L_gen <= ln < L_qpath: return "regeneration: slab ray generation"
        if L_queue <= ln < L_regen: return "kernel body around the step (staging, hoisted address math)"
        if L_regen <= ln < L_call: return "regeneration: dealing, pop"
        if L_call <= ln < L_end: return "finish (RED, counters)"
        return "other tpt_fast.cu (samplers, seeds)"
    if f == "tpt_integrator.cuh": return "shade: integrator helpers (RNG, vectors, scatter)"
    if f == "tpt_device_utils.cuh": return "kernel body around the step (staging, hoisted address math)"
    return "intrinsics headers (funnel shift, ballot, atomics)"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
h = rows[hi]
S, I, T = h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
inst = [(int(r[S]), int(r[I]), int(r[T])) for r in rows[hi + 1:] if len(r) > T and r[S].isdigit()]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
lines, cur = [], ("?", 0)
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lines.append(cur)
n = min(len(lines), len(inst))
if len(lines) != len(inst): print(f"warning: {len(lines)} instructions in the cubin vs {len(inst)} in the report (different build?)")
agg = {}
for k in range(n):
    f, ln = lines[k]
    # the funnel shifts of pass 1 live in sm_32_intrinsics.hpp: attribute header lines with full-warp execution to pass 1
    a = agg.setdefault(phase(f, ln), [0, 0, 0])
    a[0] += inst[k][0]; a[1] += inst[k][1]; a[2] += inst[k][2]
ts, ti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print(f"{kern}: {n} SASS instructions, {ti} warp instructions executed, {ts} stall samples")
print(f"{'phase':52s} {'% warp instr':>12s} {'% samples':>10s} {'active lanes':>13s}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{key:52s} {100 * a[1] / ti:12.1f} {100 * a[0] / ts:10.1f} {a[2] / max(a[1], 1):13.1f}")
