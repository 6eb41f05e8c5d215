"""profiles/*_traffic_*.json from `ncu -i X.ncu-rep --page raw --csv`: DRAM bytes per launch of K1, K2, K3 (last
launch of each kernel in the capture).  usage: traffic_from_ncu.py raw.csv out.json "<command that was profiled>" """
import csv, json, sys

KEYS = {"fp": "fp_kernel", "bp": "bp_kernel", "update": "barrier_update_kernel"}
COLS = {"dram_bytes_read": "dram__bytes_read.sum", "dram_bytes_write": "dram__bytes_write.sum",
        "duration_ms_under_ncu": "gpu__time_duration.sum", "l2_hit_pct": "lts__t_sector_hit_rate.pct",
        "l1tex_throughput_pct": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "%": 1.0}


def main(raw, out, command):
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        for key, pat in KEYS.items():
            if pat in name:
                e = {"kernel": name}
                for k, col in COLS.items():
                    e[k] = float(r[idx[col]].replace(",", "")) * SCALE.get(units[idx[col]], 1.0)
                e["traffic_bytes"] = e["dram_bytes_read"] + e["dram_bytes_write"]
                res[key] = e
    res["command"] = command
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:4])
