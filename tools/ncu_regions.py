"""Instruction share of source-line ranges of one kernel in an ncu report.
usage: ncu_regions.py <ncu-rep> <kernel-regex> file:lo-hi[:label] ..."""
import csv, subprocess, sys, io, collections
csv.field_size_limit(10**9)
rep, kre = sys.argv[1], sys.argv[2]
regions = []
for a in sys.argv[3:]:
    parts = a.split(":")
    lo, hi = parts[1].split("-")
    regions.append((parts[0], int(lo), int(hi), parts[2] if len(parts) > 2 else a))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name-base", "mangled" if "ILi" in kre else "function",  # template instantiations: match the mangled name
                      "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname = None; hdr = None; cur = None
inst = collections.Counter(); tinst = collections.Counter(); smp = collections.Counter(); sass = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or fname is None: continue
    if r[0] != "":
        try: cur = (fname, int(r[0]))
        except ValueError: pass
        continue
    if cur is None or len(r) < len(hdr) or not r[2].startswith("0x"): continue
    def g(name):
        try: return float(r[hdr.index(name)])
        except Exception: return 0.0
    inst[cur] += g("Instructions Executed"); tinst[cur] += g("Thread Instructions Executed"); smp[cur] += g("# Samples"); sass[cur] += 1
tot = sum(inst.values()); ts = sum(smp.values())
print(f"total warp-inst {tot:.4g}, samples {ts:.0f}")
rest_i = tot; rest_s = ts
for f, lo, hi, label in regions:
    keys = [k for k in inst if k[0] == f and lo <= k[1] <= hi]
    i = sum(inst[k] for k in keys); t = sum(tinst[k] for k in keys); s_ = sum(smp[k] for k in keys); n = sum(sass[k] for k in keys)
    rest_i -= i; rest_s -= s_
    print(f"{label:40s} inst {100*i/tot:5.1f}%  samples {100*s_/max(ts,1):5.1f}%  lanes {t/max(i,1):5.1f}  sass {n}")
print(f"{'(other)':40s} inst {100*rest_i/tot:5.1f}%  samples {100*rest_s/max(ts,1):5.1f}%")
