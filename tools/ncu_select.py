"""Cut an `ncu --page raw --csv` export down to the metrics the round summaries quote (metric,unit,launch0), keeping
the metric list of the existing profiles/r01_*_ncu_full_raw_selected.csv files, and refresh profiles/ncu_traffic.json
(dram bytes read + written of the dominant kernel, per launch).

    python tools/ncu_select.py gpurun_out profiles
"""
import csv
import json
import os
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    traffic_path = os.path.join(dst, "ncu_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for w in ("ovm", "superpose", "allpairs"):
        raw = os.path.join(src, f"prof_{w}_raw.csv")
        sel = os.path.join(dst, f"r01_{w}_ncu_full_raw_selected.csv")
        if not os.path.exists(raw) or not os.path.exists(sel):
            continue
        keep = [r[0] for r in csv.reader(open(sel))][1:]
        rows = list(csv.reader(open(raw)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        table = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        extra = ["smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
                 "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__inst_executed.sum",
                 "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
                 "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "Kernel Name"]
        with open(sel, "w", newline="") as f:
            wr = csv.writer(f)
            wr.writerow(["metric", "unit", "launch0"])
            for k in keep + [e for e in extra if e not in keep]:
                if k in table:
                    wr.writerow([k, table[k][0], table[k][1]])
        try:
            traffic[w] = int(float(table["dram__bytes_read.sum"][1].replace(",", "")) *
                             {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[table["dram__bytes_read.sum"][0]] +
                             float(table["dram__bytes_write.sum"][1].replace(",", "")) *
                             {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[table["dram__bytes_write.sum"][0]])
        except (KeyError, ValueError) as e:
            print("traffic not updated for", w, e)
    json.dump(traffic, open(traffic_path, "w"))
    print(traffic)


if __name__ == "__main__":
    main()
