"""Sums ncu dram bytes + durations over the decoder's umma_conv launches of one bench step.
Usage: python tools/decoder_traffic.py gpurun_out/traffic.csv  -> writes profiles/decoder_traffic.json"""
import csv, hashlib, json, os, sys


def sources_digest():
    """Same digest as bench.py: the traffic figure is only valid for the kernel sources it was measured on."""
    h = hashlib.sha1()
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sbv2-api_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.startswith("umma_") and f.endswith((".cu", ".cuh", ".h")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()


lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
per = {}
order = []
for row in csv.DictReader(lines):
    kid = row["ID"]
    if kid not in per:
        per[kid] = {"name": row["Kernel Name"]}
        order.append(kid)
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    name = row["Metric Name"]
    if name.startswith("dram__bytes"):
        mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        per[kid][name] = v * mul
    elif name == "gpu__time_duration.sum":
        per[kid]["us"] = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
# the decoder's tensor-core launches are the umma_conv / umma_pair launches after the step's last to_planar_kernel
names = [per[k]["name"] for k in order]
last_tp = max(i for i, n in enumerate(names) if "to_planar_kernel" in n)
dec = [per[k] for i, k in enumerate(order) if i > last_tp and ("umma_conv" in per[k]["name"] or "umma_pair" in per[k]["name"])]
rd = sum(k.get("dram__bytes_read.sum", 0) for k in dec)
wr = sum(k.get("dram__bytes_write.sum", 0) for k in dec)
us = sum(k.get("us", 0) for k in dec)
out = {"sources_sha1": sources_digest(), "launches": len(dec), "dram_bytes_read_per_step": rd, "dram_bytes_write_per_step": wr, "dram_bytes_per_step": rd + wr,
       "kernel_time_us_under_ncu": us, "note": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum, cold-cache serialised replays; bench workload (32 x ~8 s)"}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "decoder_traffic.json"), "w"), indent=1)
print(out)
