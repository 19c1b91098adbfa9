"""Turns `ncu -i <rep> --page raw --csv` into the per-kernel summary committed under profiles/ and the traffic
file bench.py reads (profiles/r2_traffic_<what>.json).

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_summary.py /tmp/raw.csv profiles/r2_ncu_loss_a --images 16 --anchors 22400 [--json profiles/r2_traffic_loss.json]
"""
import argparse
import csv
import json
import re

COLS = [("gpu__time_duration.sum", "us", "time_us"),
        ("dram__bytes_read.sum", "byte", "dram_read_bytes"),
        ("dram__bytes_write.sum", "byte", "dram_write_bytes"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "%", "dram_throughput_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "%", "sm_throughput_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "%", "warps_active_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "%", "issue_active_pct"),
        ("smsp__inst_executed.sum", "inst", "warp_instructions"),
        ("launch__registers_per_thread", "", "registers"),
        ("launch__grid_size", "", "grid"),
        ("launch__block_size", "", "block"),
        ("launch__waves_per_multiprocessor", "", "waves"),
        ("lts__t_sector_hit_rate.pct", "%", "l2_hit_pct"),
        ("l1tex__t_sector_hit_rate.pct", "%", "l1_hit_pct")]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"paa::", "", name)
    return name.strip()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw_csv")
    ap.add_argument("out_prefix")
    ap.add_argument("--images", type=int, default=None)
    ap.add_argument("--anchors", type=int, default=None)
    ap.add_argument("--json", default=None)
    ap.add_argument("--title", default="")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.raw_csv)))
    # ncu prints a header row with metric names and a second row with units
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
    col = {n: i for i, n in enumerate(names)}
    per = {}
    order = []
    for r in data:
        if len(r) < len(names):
            continue
        k = short(r[col["Kernel Name"]])
        rec = per.setdefault(k, {"launches": 0})
        if k not in order:
            order.append(k)
        rec["launches"] += 1
        for metric, _, key in COLS:
            if metric not in col:
                continue
            try:
                v = float(r[col[metric]].replace(",", ""))
            except ValueError:
                continue
            u = units[col[metric]]
            v *= UNIT.get(u, 1.0)
            rec.setdefault("_" + key, []).append(v)
    out = {}
    for k in order:
        rec = per[k]
        out[k] = {"launches": rec["launches"]}
        for _, _, key in COLS:
            vs = rec.get("_" + key)
            if vs:
                out[k][key] = sum(vs) / len(vs)
    with open(a.out_prefix + "_raw.md", "w") as f:
        f.write("# %s\n\nper launch (mean over the captured launches); `ncu --set full --clock-control none`\n\n" % a.title)
        keys = [c[2] for c in COLS]
        f.write("| kernel | launches | " + " | ".join(keys) + " |\n|---|---|" + "---|" * len(keys) + "\n")
        for k in order:
            f.write("| %s | %d | " % (k, out[k]["launches"]) +
                    " | ".join(("%.4g" % out[k][x]) if x in out[k] else "" for x in keys) + " |\n")
    if a.json:
        # keyed by the kernel name without template arguments, the way bench.py reports kernels
        kernels = {}
        for k in order:
            base = re.sub(r"<.*$", "", k)
            d = kernels.setdefault(base, {})
            for key in ("dram_read_bytes", "dram_write_bytes", "time_us", "issue_active_pct", "warps_active_pct",
                        "dram_throughput_pct"):
                if key in out[k]:
                    d[key] = out[k][key]
        json.dump({"source": a.title, "images": a.images, "anchors_per_image": a.anchors, "kernels": kernels},
                  open(a.json, "w"), indent=1)
    print(open(a.out_prefix + "_raw.md").read())


if __name__ == "__main__":
    main()
