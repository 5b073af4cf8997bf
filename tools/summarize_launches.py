"""Turn an ncu `--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,
sm__pipe_tensor_cycles_active...` launch list (csv) into a per-kernel table and, for the RetinaFace forward, the sums
`bench.py` cites as `roofline.traffic`:

  python tools/summarize_launches.py profiles/r02_fwd_glue_launches_v2.csv --forward 71 --json profiles/r02_fwd_traffic.json
"""
import argparse
import collections
import csv
import io
import json


def load(path):
    rows = [l for l in open(path) if l.startswith('"')]
    by = collections.OrderedDict()
    for x in csv.DictReader(io.StringIO("".join(rows))):
        d = by.setdefault(int(x["ID"]), {"name": x["Kernel Name"].split("(")[0].replace("void ", "").replace("frb::", ""),
                                          "grid": x["Grid Size"], "block": x["Block Size"]})
        d[x["Metric Name"]] = float(x["Metric Value"].replace(",", ""))
    return by


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--forward", type=int, default=0, help="the first N conv / pool launches after --skip are one forward")
    ap.add_argument("--skip", type=int, default=0, help="conv/maxpool launches to skip (warm-up forward)")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    by = load(a.csv)
    fwd = [d for d in by.values() if d["name"].startswith(("conv_", "maxpool", "pool_border"))][a.skip:]
    if a.forward:
        fwd = fwd[:a.forward]
    print(f"{'kernel':42s} {'grid':>14s} {'us':>9s} {'dramR MB':>9s} {'dramW MB':>9s} {'GB/s':>8s} {'L2 MB':>9s} {'tensor%':>8s}")
    for i, d in by.items():
        t = d["gpu__time_duration.sum"] / 1e3
        r, w = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
        print(f"{d['name'][:42]:42s} {d['grid']:>14s} {t:9.1f} {r / 1e6:9.1f} {w / 1e6:9.1f} {(r + w) / t / 1e3:8.1f} "
              f"{d.get('lts__t_bytes.sum', 0.0) / 1e6:9.1f} {d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0.0):8.1f}")
    if a.forward:
        t = sum(d["gpu__time_duration.sum"] for d in fwd) / 1e6
        tw = sum(d["gpu__time_duration.sum"] * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) for d in fwd) / \
             sum(d["gpu__time_duration.sum"] for d in fwd)
        out = {"source": a.csv, "launches": len(fwd), "serialized_ms": t,
               "dram_read_bytes": sum(d.get("dram__bytes_read.sum", 0.0) for d in fwd),
               "dram_write_bytes": sum(d.get("dram__bytes_write.sum", 0.0) for d in fwd),
               "l2_bytes": sum(d.get("lts__t_bytes.sum", 0.0) for d in fwd),
               "tensor_pipe_active_pct_time_weighted": tw,
               "pair_launches": sum(1 for d in fwd if d["name"].startswith("conv_tma_kernel<") and
                                    [x.strip() for x in d["name"].split("<")[1].rstrip(">").split(",")][4:5] == ["1"]),
               "kernels": dict(collections.Counter(d["name"].split("<")[0] for d in fwd))}
        out["dram_bytes"] = out["dram_read_bytes"] + out["dram_write_bytes"]
        print(json.dumps(out))
        if a.json:
            json.dump(out, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
