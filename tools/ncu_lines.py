"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.
usage: ncu_lines.py <ncu-rep> <kernel-regex> [top]"""
import csv, subprocess, sys, io, collections
csv.field_size_limit(10**9)
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
sortkey = sys.argv[4] if len(sys.argv) > 4 else "samples"
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name-base", "mangled" if "ILi" in kre else "function",  # template instantiations: match the mangled name
                      "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname = None; hdr = None
agg = collections.defaultdict(lambda: collections.Counter())
src = {}
tot = collections.Counter()
kernels = 0
cur = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or fname is None: continue
    if r[0] != "":
        try: cur = (fname, int(r[0]))
        except ValueError: continue
        src[cur] = r[1].strip() if len(r) > 1 else ""
        continue
    if cur is None or len(r) < len(hdr) or not r[2].startswith("0x"): continue
    def g(name):
        try: return float(r[hdr.index(name)])
        except Exception: return 0.0
    a = agg[cur]
    a["inst"] += g("Instructions Executed"); a["tinst"] += g("Thread Instructions Executed"); a["samples"] += g("# Samples")
    for s_ in ("stall_barrier","stall_long_sb","stall_short_sb","stall_mio","stall_lg","stall_wait","stall_math","stall_branch_resolving","stall_no_inst","stall_not_selected","stall_dispatch","stall_selected"):
        a[s_] += g(s_)
    a["bankx"] += g("L1 Wavefronts Shared Excessive")
for k, a in agg.items():
    for n, v in a.items(): tot[n] += v
print(f"total warp-inst {tot['inst']:.3g} thread-inst/warp-inst {tot['tinst']/max(tot['inst'],1):.1f} samples {tot['samples']:.0f}")
stalls = {n: v for n, v in tot.items() if n.startswith("stall_")}
print("stall mix:", ", ".join(f"{n[6:]} {100*v/max(sum(stalls.values()),1):.0f}%" for n, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:7]))
print(f"{'file:line':28s} {'%inst':>6s} {'%smp':>6s} {'lanes':>5s} {'bankx':>8s}  top-stall  source")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][sortkey])[:top]:
    st = max(((n, v) for n, v in a.items() if n.startswith("stall_")), key=lambda kv: kv[1], default=("", 0))
    print(f"{key[0]+':'+str(key[1]):28s} {100*a['inst']/max(tot['inst'],1):6.1f} {100*a['samples']/max(tot['samples'],1):6.1f} {a['tinst']/max(a['inst'],1):5.1f} {a['bankx']:8.0f}  {st[0][6:]:10s} {src.get(key,'')[:90]}")
