#!/usr/bin/env python
"""Summary of one `ncu --set full` capture for profiles/: the metrics the roofline discussion uses, plus (optionally) an
entry of profiles/ncu_traffic.json keyed to the SHA-1 of the kernel's source file, which is what bench.py checks before
it prints `roofline.traffic`.

    python tools/ncu_summary.py gpurun_out/X.ncu-rep "title / command" [--traffic-key KEY --src-file mm2d3d_b200/csrc/F.cu] > profiles/X.txt
"""
import argparse
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__cycles_elapsed.max", "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ldgsts.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("title")
    ap.add_argument("--traffic-key")
    ap.add_argument("--src-file")
    ap.add_argument("--alg-bytes", type=float, default=0.0)
    ap.add_argument("--out-name", help="file name under profiles/ this summary is saved as (recorded in ncu_traffic.json)")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    get = {}
    for h, u, v in zip(hdr, units, vals):
        get[h] = (v, u)
    print(f"# {a.title}")
    print(f"# source: ncu --set full --clock-control none --import-source on, one launch ({os.path.basename(a.rep)}); kernel: "
          f"{get['Kernel Name'][0][:70]}  grid {get['Grid Size'][0]} block {get['Block Size'][0]}")
    for k in WANT:
        for h in hdr:
            if h == k or h.endswith("." + k) and k.startswith("sm__pipe_tensor_subpipe"):
                v, u = get[h]
                if v not in ("", "n/a"):
                    print(f"{h[:96]:96s} {v:>18s} {u}")
                break
    def num(k):
        v, u = get[k]
        x = float(v.replace(",", ""))
        return x * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9}.get(u, 1.0)
    dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    dur = num("gpu__time_duration.sum")
    print(f"# derived: DRAM traffic {dram / 1e6:.2f} MB per launch = {dram / dur / 1e9:.0f} GB/s over the launch"
          + (f"; algorithmic bytes {a.alg_bytes / 1e6:.2f} MB -> traffic / algorithmic = {dram / a.alg_bytes:.2f}, "
             f"algorithmic {a.alg_bytes / dur / 1e9:.0f} GB/s under ncu (cold caches, serialised)" if a.alg_bytes else ""))
    if a.traffic_key and a.src_file:
        path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        tr = json.load(open(path)) if os.path.exists(path) else {}
        sha = hashlib.sha1(b"".join(open(os.path.join(ROOT, f), "rb").read() for f in
                                     (a.src_file, "mm2d3d_b200/csrc/tc_common.cuh", "mm2d3d_b200/csrc/plan.cuh"))).hexdigest()
        tr[a.traffic_key] = {"dram_bytes": int(dram), "src_file": a.src_file, "src_sha1": sha,
                             "source": f"profiles/{a.out_name or os.path.basename(a.rep).replace('.ncu-rep', '.txt')} (ncu --set full: "
                                       "dram__bytes_read.sum + dram__bytes_write.sum, one launch)"}
        json.dump(tr, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
