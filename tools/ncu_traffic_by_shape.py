#!/usr/bin/env python
"""Join an ncu metrics pass of ONE resident bench step with the library's own per-launch shape log, so that DRAM traffic
is reported PER SHAPE next to that shape's algorithmic bytes (bench.py `roofline.traffic_by_shape`).

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      --profile-from-start off --csv --log-file gpurun_out/traffic_TAG.csv python bench.py ... --dump-shapes gpurun_out/shapes_TAG.json
  python tools/ncu_traffic_by_shape.py gpurun_out/traffic_TAG.csv gpurun_out/shapes_TAG.json profiles/r02_gemm_traffic_by_shape.json

The shape log lists, in launch order, every tensor-core launch of one step (class, M, N, K) as recorded by
ir_profile_records; the ncu log lists every kernel of the profiled step in launch order. Both are filtered to the same
kernel families (gemm_tc_kernel = classes gemm + conv, attn_tc_kernel = attention) and zipped; a count mismatch aborts."""
import csv
import json
import sys
from collections import OrderedDict


def read_ncu(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
    by_id = OrderedDict()
    for r in rd:
        d = by_id.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]]})
        try:
            v = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        d[r[ix["Metric Name"]]] = v * scale.get(r[ix["Metric Unit"]], 1.0)
    for d in by_id.values():
        rows.append(d)
    return rows


def alg_bytes(k, m, n, kk):
    if k == "gemm":
        return 2.0 * m * kk + 2.0 * n * kk + 2.0 * m * n
    if k == "conv":
        cin = kk // 9 if kk % 9 == 0 else kk // 4
        return 2.0 * m * cin + 2.0 * n * kk + 2.0 * m * n
    if k == "attention":
        return 4 * 2.0 * m * n * kk
    return None


def main():
    ncu_csv, shapes_json, out = sys.argv[1:4]
    ncu = read_ncu(ncu_csv)
    shapes = json.loads(open(shapes_json).read())
    fam = {"gemm": "gemm_tc_kernel", "conv": "gemm_tc_kernel", "attention": "attn_tc_kernel"}
    s_rows = [s for s in shapes if s["class"] in fam]
    n_rows = [r for r in ncu if ("gemm_tc_kernel" in r["name"] or ("attn_tc_kernel" in r["name"] and "xattn" not in r["name"]))]
    if len(s_rows) != len(n_rows):
        sys.exit(f"launch count mismatch: shape log {len(s_rows)} vs ncu {len(n_rows)}")
    agg = OrderedDict()
    tot_b, tot_n = 0.0, 0
    for s, r in zip(s_rows, n_rows):
        assert fam[s["class"]] in r["name"], (s, r["name"])
        b = r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0)
        a = agg.setdefault((s["class"], s["M"], s["N"], s["K"]), {"launches": 0, "dram": 0.0, "us": 0.0, "kernel": r["name"].split("(")[0].replace("void ", "")})
        a["launches"] += 1
        a["dram"] += b
        a["us"] += r.get("gpu__time_duration.sum", 0.0)
        if s["class"] in ("gemm", "conv"):
            tot_b += b
            tot_n += 1
    rows = []
    for (k, m, n, kk), a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        alg = alg_bytes(k, m, n, kk)
        rows.append({"class": k, "M": m, "N": n, "K": kk, "kernel": a["kernel"], "launches_per_step": a["launches"],
                     "dram_bytes_per_launch": a["dram"] / a["launches"], "algorithmic_bytes": alg,
                     "traffic_over_algorithmic": (a["dram"] / a["launches"] / alg) if alg else None,
                     "ncu_us_per_launch": a["us"] / a["launches"]})
    doc = {"source": f"{out}: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum (cold L2 per launch, --clock-control none) of one "
                     "resident 1024x1024 step, joined in launch order with the library's per-launch shape log (tools/ncu_traffic_by_shape.py)",
           "mean_dram_bytes_per_launch": tot_b / max(1, tot_n), "gemm_conv_launches": tot_n, "shapes": rows}
    open(out, "w").write(json.dumps(doc, indent=1))
    print(f"{len(rows)} shapes, {tot_n} gemm/conv launches -> {out}")


if __name__ == "__main__":
    main()
