"""profiles/r1_traffic.json from an ncu --set full report: DRAM bytes (read + write) per launch of every
hand-written kernel, and per C-ABI entry point (sum of its kernels, averaged over the captured layers).
usage: python scratch/make_traffic.py gpurun_out/prof.ncu-rep profiles/r1_traffic.json"""
import collections, csv, io, json, re, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
ri, wi, ti = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per = collections.OrderedDict()
for r in rows[2:]:
    name = re.sub(r"^void ", "", r[ki])
    name = re.split(r"[(<]", name)[0].replace("bliss::", "")
    b = float(r[ri]) * mult[units[ri]] + float(r[wi]) * mult[units[wi]]
    t = float(r[ti]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[units[ti]]
    d = per.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "time_us": 0.0})
    d["launches"] += 1
    d["dram_bytes"] += b
    d["time_us"] += t
ENTRY = {"bliss_frontier_prob": ["k_prob_pass1", "k_prob_pass2", "k_prob_pass3", "k_collect_candidates"],
         "bliss_block_count": ["k_block_count"], "bliss_block_index": ["k_block_index"],
         "bliss_block_fill": ["k_block_fill"], "bliss_spmm": ["k_spmm_item", "k_spmm_seg", "k_spmm_combine_items", "k_spmm_combine"]}
res = {"_source": f"ncu --set full --clock-control none, {rep} (scratch/exp_layer.py: Reddit-shape step, stage path; "
                  "cold-cache replays, so kernels that hit L2 in a real step show their inputs as DRAM reads here); "
                  "dram__bytes_read.sum + dram__bytes_write.sum per launch", "kernels": {}}
for k, d in per.items():
    res["kernels"][k] = {"launches": d["launches"], "dram_bytes_per_launch": d["dram_bytes"] / d["launches"],
                         "avg_us_under_ncu": d["time_us"] / d["launches"]}
for e, ks in ENTRY.items():
    have = [k for k in ks if k in per]
    if not have:
        continue
    calls = max(per[k]["launches"] for k in have)
    res[e] = {"dram_bytes_per_launch": sum(per[k]["dram_bytes"] for k in have) / calls, "launches": calls,
              "kernels": have}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k.startswith("bliss_")}, indent=1))
