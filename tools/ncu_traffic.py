"""profiles/traffic.json from an `ncu --set full` report of tools/prof_compress.py:
DRAM bytes (read + write) per launch of every pipeline kernel, keyed by bench.py's stage names.
usage: ncu_traffic.py <ncu-rep> <frames_per_launch> <out.json>"""
import csv, io, json, subprocess, sys
rep, frames, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, data = rows[0], rows[1], rows[2:]
stage = {"k_xdelta_planes": "transform", "k_hzr_hist": "hist", "k_hzr_tree": "tree", "k_scan_offsets": "layout",
         "k_hzr_encode": "encode", "k_frame_parse": "parse", "k_hzr_decode": "decode", "k_planes_to_samples": "inverse"}
def num(r, n):
    v = float(r[hdr.index(n)].replace(",", ""))
    u = units[hdr.index(n)]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
res = {}
for r in data:
    name = r[hdr.index("Kernel Name")]
    for k, st in stage.items():
        if k in name:
            kn = name.split("(")[0].replace("void ", "").replace("rspt::", "")
            e = {"kernel": kn, "dram_bytes_per_launch": num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum"),
                 "dram_read": num(r, "dram__bytes_read.sum"), "dram_write": num(r, "dram__bytes_write.sum"),
                 "duration_us": float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))}
            if st in res:
                # a stage made of several launches (encode = k_hzr_encode_sparse + k_hzr_encode; the decoder's three
                # payload classes; the two tree launches): add them up -- the capture holds ONE pass of the pipeline
                for key in ("dram_bytes_per_launch", "dram_read", "dram_write", "duration_us"):
                    e[key] += res[st][key]
                prev = res[st]["kernel"].split(" + ")
                e["kernel"] = res[st]["kernel"] if kn in prev else res[st]["kernel"] + " + " + kn
                e["launches"] = res[st].get("launches", 1) + 1
            res[st] = e
            break
json.dump({"source": rep.split("/")[-1], "frames_per_launch": frames, "kernels": res}, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
