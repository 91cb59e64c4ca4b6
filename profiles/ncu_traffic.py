"""profiles/<tag>_full_sel.csv (ncu --set full, one row per launch; written by profiles/run_gpu.sh ncu) ->
profiles/r2_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per timed C-ABI call of bench.py's roofline
(all kernels of the call, per step), plus a per-kernel markdown table on stdout.

  python profiles/ncu_traffic.py profiles/r2_ncu_full_sel.csv profiles/r2_traffic.json
"""
import csv
import json
import re
import sys

GROUPS = {                                  # bench.py roofline key -> kernels the call launches (C2: layer 2 = <NCH 2, H 1>)
    "c2:edge_attn_fwd": r"edge_fwd_stream_kernel|edge_fwd_hub_finalize",
    "c2:edge_attn_bwd_split:cols": r"split_cols_(tasks_)?kernel<2, 1",       # (layer 1 runs the <1, 2, ..> instances)
    "c2:edge_attn_bwd_split:rels": r"split_rels_(tasks_)?kernel<2, 1",
    "c2:edge_attn_bwd_split:node": r"bwd_node_kernel",
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        i = col[name]
        try:
            return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
        except ValueError:
            return 0.0

    per = {}
    for r in data:
        name = r[col["Kernel Name"]]
        short = name.split("(")[0]
        d = per.setdefault(short, {"n": 0, "bytes": 0.0, "ns": 0.0})
        d["n"] += 1
        d["bytes"] += val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        tu = units[col["gpu__time_duration.sum"]]
        d["ns"] += float(r[col["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(tu, 1)
    out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per timed call (all its kernels, per step), from " + src}
    for key, pat in GROUPS.items():
        ks = [k for k in per if re.search(pat, k)]
        if not ks:
            continue
        main_k = max(ks, key=lambda k: per[k]["bytes"])
        steps = max(1, per[main_k]["n"])
        out[key] = int(sum(per[k]["bytes"] for k in ks) / steps)
    json.dump(out, open(dst, "w"), indent=1)
    print("| kernel | launches | avg ms | DRAM rd+wr GB per launch | GB/s |")
    print("|---|---:|---:|---:|---:|")
    for k, d in sorted(per.items(), key=lambda kv: -kv[1]["ns"]):
        ms = d["ns"] / d["n"] / 1e6
        gb = d["bytes"] / d["n"] / 1e9
        print(f"| `{k[:70]}` | {d['n']} | {ms:.3f} | {gb:.2f} | {gb / ms * 1e3 if ms else 0:.0f} |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
