#!/usr/bin/env python
"""DRAM traffic per kernel class from an `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum` log of ONE
eager step (tools/profile_step.py).  Writes the JSON that bench.py reports as roofline.traffic.
usage: ncu_traffic.py launches.csv out.json "<workload string>" """
import csv
import json
import re
import sys

CLASS = [("conv_tc_kernel|conv_halo_kernel|conv_chain_kernel|strip_gemm_kernel|conv_simt|conv_direct", "fce_conv2d"),
         ("conv_dwpw_kernel", "fce_dwpw_conv"), ("conv_stem2_kernel", "fce_stem2_conv"),
         ("dwconv", "fce_dwconv3x3"), ("stem_fused", "fce_stem_conv"), ("decode_kernel", "fce_detect_decode"),
         ("gate_", "fce_gate_apply"), ("bifpn_kernel", "fce_bifpn_fuse"), ("coord_pool", "fce_coord_pool"),
         ("coordatt_mlp", "fce_coordatt_mlp"), ("psa_", "fce_psa_attention"), ("sppf", "fce_sppf_pool"),
         ("nms_", "fce_nms"), ("strip_attn", "fce_strip_attn"), ("upsample", "fce_upsample2x"), ("copy_kernel", "fce_copy_view")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

per_id = {}
for row in csv.reader(open(sys.argv[1], errors="replace")):
    if len(row) < 15 or not row[0].isdigit():
        continue
    kid, name, metric, unit, val = int(row[0]), row[4], row[12], row[13], row[14]
    if not metric.startswith("dram__bytes"):
        continue
    e = per_id.setdefault(kid, {"name": name, "bytes": 0.0})
    e["bytes"] += float(val.replace(",", "")) * UNIT.get(unit, 1.0)
classes = {}
for kid, e in sorted(per_id.items()):
    cls = next((c for pat, c in CLASS if re.search(pat, e["name"])), None)
    if cls is None:
        continue
    c = classes.setdefault(cls, {"dram_bytes_per_step": 0.0, "kernel_launches": 0})
    c["dram_bytes_per_step"] += e["bytes"]
    c["kernel_launches"] += 1
out = {"workload": sys.argv[3] if len(sys.argv) > 3 else "", "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, "
       "one eager step (tools/profile_step.py), cold-cache serialised launches", "classes": classes}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out))
