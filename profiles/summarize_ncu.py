"""Turn the raw ncu CSV exports of profiles/run_final_r01.sh into the committed text summaries.

  python profiles/summarize_ncu.py gpurun_out  ->  profiles/launches_*_r01_summary.txt, profiles/ncu_hot_r01_summary.txt,
                                                    profiles/ncu_rollout_r01_summary.txt, profiles/roofline_traffic.json
"""
import collections
import csv
import json
import os
import re
import sys

src = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "gpurun_out"
out = os.path.dirname(os.path.abspath(__file__))


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("void ", "").replace("hb::", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("(anonymous namespace)::", "")
    return name.strip()[:70]


def launch_summary(csv_path, txt_path, title):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 10]
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    d = collections.OrderedDict()
    for r in rows[1:]:
        d.setdefault(short(r[ki]), []).append(float(r[vi].replace(",", "")) / 1e3)
    tot = sum(sum(v) for v in d.values())
    with open(txt_path, "w") as f:
        f.write(f"# {title}\n# ncu --metrics gpu__time_duration.sum --clock-control none (serialised, cold caches: compare SHARES)\n")
        f.write(f"# {sum(len(v) for v in d.values())} launches, {tot / 1e3:.3f} ms total\n")
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k:72s} n={len(v):4d} total={sum(v) / 1e3:9.3f} ms avg={sum(v) / len(v):9.2f} us share={sum(v) / tot:6.3f}\n")


WANT = [("gpu__time_duration.sum", "duration_us", 1.0), ("launch__grid_size", "grid", 1), ("launch__block_size", "block", 1),
        ("launch__registers_per_thread", "regs", 1), ("dram__bytes_read.sum", "dram_read", 1), ("dram__bytes_write.sum", "dram_write", 1),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak", 1), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct", 1),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex_pct", 1), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct", 1), ("smsp__inst_executed.sum", "warp_instructions", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct", 1),
        ("sm__inst_executed_pipe_tc.sum", "tensor_pipe_instructions", 1), ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct", 1),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard", 1),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_scoreboard", 1),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier", 1),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait", 1),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio_throttle", 1),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_instruction", 1)]


def unit_scale(unit):
    return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(unit, 1.0)


def full_summary(csv_path, txt_path, title):
    rows = list(csv.reader(open(csv_path)))
    hdr, units = rows[0], rows[1]
    traffic = {}
    # one block per kernel template: the longest instance (e.g. the K=128 layer of the forward kernel, not the K=20 one)
    best = {}
    di = hdr.index("gpu__time_duration.sum")
    for r in rows[2:]:
        name = short(r[hdr.index("Kernel Name")])
        dur = float(r[di].replace(",", "")) * unit_scale(units[di])
        if name not in best or dur > best[name][0]:
            best[name] = (dur, r)
    with open(txt_path, "w") as f:
        f.write(f"# {title}\n# ncu --set full --clock-control none; the longest instance of each kernel template\n")
        for name, (_, r) in best.items():
            f.write(f"\n== {name}\n")
            vals = {}
            for metric, label, _ in WANT:
                if metric in hdr:
                    i = hdr.index(metric)
                    try:
                        v = float(r[i].replace(",", "")) * unit_scale(units[i])
                    except ValueError:
                        continue
                    vals[label] = v
                    f.write(f"   {label:28s} {v:16.3f}   ({metric})\n")
            if "dram_read" in vals and "duration_us" in vals:
                tb = vals["dram_read"] + vals.get("dram_write", 0.0)
                f.write(f"   {'dram_traffic_bytes':28s} {tb:16.0f}\n   {'dram_GB_per_s':28s} {tb / vals['duration_us'] / 1e3:16.1f}\n")
                traffic[name] = tb
    return traffic


def main_r01():
    launch_summary(os.path.join(src, "launches_update_r01.csv"), os.path.join(out, "launches_update_r01_summary.txt"),
                   "update phase of one C2 iteration (bench.py --steps 1 --warmup 1), library kernels only")
    launch_summary(os.path.join(src, "launches_rollout_r01.csv"), os.path.join(out, "launches_rollout_r01_summary.txt"),
                   "rollout steps of one C2 iteration (CUDA-graph replay) + the GAE launch")
    t1 = full_summary(os.path.join(src, "ncu_hot_r01_raw.csv"), os.path.join(out, "ncu_hot_r01_summary.txt"),
                      "update-phase kernels at the C2 launch size (819200 rows, profiles/ncu_target.py)")
    t2 = full_summary(os.path.join(src, "ncu_rollout_r01_raw.csv"), os.path.join(out, "ncu_rollout_r01_summary.txt"),
                      "rollout inference kernel and GAE kernel inside bench.py")
    label_of = {"tc_linear_ln_fwd_kernel<128, 0, 3>": "tc_linear_ln_fwd_3xtf32", "tc_dx_ln_bwd_kernel<128, 0, 3>": "tc_dx_ln_bwd_3xtf32",
                "tc_dw_accum_kernel<128, 3>": "tc_dw_accum_3xtf32", "discrete_rows_kernel<8, 16, 5, 2, 0>": "policy_head_grad",
                "discrete_rows_kernel<8, 16, 5, 1, 0>": "policy_head_eval", "value_rows_grad_kernel<8, 16, 0>": "value_head_grad"}
    traffic = {label_of[k]: v for k, v in {**t1, **t2}.items() if k in label_of}
    json.dump(traffic, open(os.path.join(out, "roofline_traffic.json"), "w"), indent=1)
    print(traffic)


if __name__ == "__main__":
    if len(sys.argv) >= 5 and sys.argv[1] == "--full":   # python profiles/summarize_ncu.py --full raw.csv out.txt "title"
        print(full_summary(sys.argv[2], sys.argv[3], sys.argv[4]))
    else:
        main_r01()
