"""Cut an `ncu --page raw --csv` export down to the metrics the round summaries quote (metric,unit,launch0), keeping
the metric list of the existing profiles/r01_*_ncu_full_raw_selected.csv files, and refresh profiles/ncu_traffic.json
(dram bytes read + written of the dominant kernel, per launch).

    python tools/ncu_select.py gpurun_out profiles [round prefix, default r02]

Reads <src>/<prefix>_prof_<workload>_raw.csv (tools/collect_profiles.sh), writes
<dst>/<prefix>_<workload>_ncu_full_raw_selected.csv; ncu_traffic.json also gets the all-pairs kernel's tensor-pipe
activity and L2 -> SM operand traffic, which bench.py quotes next to its own timing.
"""
import csv
import json
import os
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rnd = sys.argv[3] if len(sys.argv) > 3 else "r02"
    traffic_path = os.path.join(dst, "ncu_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for w, src_name in (("ovm", "ovm"), ("superpose", "superpose"), ("allpairs", "allpairs_20k")):
        raw = os.path.join(src, f"{rnd}_prof_{src_name}_raw.csv")
        lst = os.path.join(dst, f"r01_{w}_ncu_full_raw_selected.csv")   # the metric list of round 1 is kept
        sel = os.path.join(dst, f"{rnd}_{w}_ncu_full_raw_selected.csv")
        if not os.path.exists(raw) or not os.path.exists(lst):
            continue
        keep = [r[0] for r in csv.reader(open(lst))][1:]
        rows = list(csv.reader(open(raw)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        table = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        extra = ["smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
                 "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__inst_executed.sum",
                 "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
                 "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "Kernel Name",
                 "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                 "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
                 "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum", "gpu__time_duration.sum",
                 "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                 "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
                 "smsp__cycles_active.avg", "sm__cycles_active.avg"]
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
        if w == "allpairs":
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
            for key, metric in (("allpairs_tensor_pipe_active_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                                ("allpairs_l2_to_sm_bytes", "l1tex__m_xbar2l1tex_read_bytes.sum")):
                if metric in table:
                    u, v = table[metric]
                    traffic[key] = float(v.replace(",", "")) * scale.get(u, 1)
    json.dump(traffic, open(traffic_path, "w"))
    print(traffic)


if __name__ == "__main__":
    main()
