#!/usr/bin/env python
"""Map an ncu SASS source-page CSV onto CUDA source lines using nvdisasm line info.

usage: ncu_lines.py <report.ncu-rep> <lib.so> <mangled-substring e.g. k_stepILi64> [top_n]
Prints, per source line (file:line incl. the inlining function), the share of warp-stall samples, of executed
warp instructions, and the average active threads.  (ncu's own --print-source cuda CSV carries no metrics.)
"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, so, kern = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 50
# <lib.so> may also be the build directory: every *.o holds one cubin (the .so holds several with the same name)
inputs = [os.path.join(so, f) for f in sorted(os.listdir(so)) if f.endswith(".o")] if os.path.isdir(so) else [so]
line_of = {}  # (func, idx) -> "file:line"
op_of = {}
for inp in inputs:
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(inp)], cwd=tmp, check=False, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for cub in os.listdir(tmp):
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
        func, cur, idx = None, "?", 0
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                func, idx, cur = m.group(1), 0, "?"
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
            if m:
                cur = f"{os.path.basename(m.group(1))}:{m.group(2)}"
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*)", ln)
            if func and m:
                line_of[(func, idx)] = cur
                toks = [t for t in m.group(1).replace("{", " ").split() if not t.startswith("@")]
                op_of[(func, idx)] = toks[0].split(".")[0].rstrip(";") if toks else ""
                idx += 1
# 2. ncu sass page
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
agg = collections.defaultdict(lambda: [0, 0, 0])
hdr = None
func = None
funcs = sorted({f for f, _ in line_of})
k = 0
mismatch = 0
done_first = False
for r in rows:
    if r and r[0] == "Kernel Name":
        if done_first:
            break  # first MATCHING profiled launch only
        name = r[1]
        func = None
        want_v = re.search(r"ILi(\d+)", kern)  # k_stepILi16 -> launches of k_step<(int)16, ...> only
        if re.sub(r"I?Li\d+.*", "", kern) in name and (not want_v or (f"<(int){want_v.group(1)}," in name or f"<{want_v.group(1)}," in name)):
            cands = [f for f in funcs if kern in f]
            func = cands[0] if cands else None
        k = 0
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if func is None or hdr is None or len(r) < len(hdr) - 2:
        continue
    done_first = True
    i_s, i_i, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    toks = [t for t in r[1].replace("{", " ").split() if not t.startswith("@")]
    sass_op = toks[0].split(".")[0].rstrip(";") if toks else ""
    want_op = op_of.get((func, k), "")
    if want_op and sass_op and want_op != sass_op:
        mismatch += 1
    a = agg[line_of.get((func, k), "?")]
    a[0] += int(r[i_s] or 0); a[1] += int(r[i_i] or 0); a[2] += int(r[i_t] or 0)
    k += 1
ts = sum(a[0] for a in agg.values()) or 1
ti = sum(a[1] for a in agg.values()) or 1
print(f"kernel {kern}: {k} SASS instructions, {ti} warp-inst executed, {ts} samples")
if k == 0 or mismatch > k // 20:
    sys.exit(f"the library does not match the profiled binary ({mismatch} opcode mismatches over {k} instructions): rebuild the profiled revision first")
src_cache = {}
def src(loc):
    f, _, n = loc.partition(":")
    for d in ("topotrafficrl_b200/csrc", "include"):
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", d, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            try:
                return src_cache[p][int(n) - 1].strip()[:100]
            except Exception:
                return ""
    return ""
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{a[0]/ts*100:5.1f}% smp {a[1]/ti*100:5.1f}% inst thr/inst {a[2]/max(a[1],1):5.1f} | {loc:22s} {src(loc)}")
