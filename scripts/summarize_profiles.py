"""Turn the raw ncu outputs of scripts/gpu_final.sh (gpurun_out/) into the small tracked summaries under profiles/.

    python scripts/summarize_profiles.py r01c

  launches_<tag>.csv        -> profiles/<tag>_launches_summary_note_encoder.csv   (per kernel: launches, total ns, share)
  launches_train_<tag>.csv  -> profiles/<tag>_launches_summary_train_step.csv     (one step, no CUDA graph)
  prof_{gemm,attn,ln}_<tag>.ncu-rep -> profiles/<tag>_ncu_<kernel>_full_summary.csv (selected metrics, one column per
                                        captured launch) and profiles/roofline_traffic.json (dram bytes per launch)
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__block_size",
        "launch__grid_size", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__pcsamp_sample_count",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_wait", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def launch_summary(src, dst, header, last_step_marker=None):
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
    if not rows:
        return
    ix = {h: i for i, h in enumerate(rows[0])}
    body = rows[1:]
    if last_step_marker:
        ends = [i for i, r in enumerate(body) if last_step_marker in r[ix["Kernel Name"]]]
        if len(ends) >= 2:
            body = body[ends[-2] + 1:ends[-1] + 1]
    agg = collections.OrderedDict()
    for r in body:
        d = agg.setdefault(r[ix["Kernel Name"]][:90], [0, 0.0])
        d[0] += 1
        d[1] += float(r[ix["Metric Value"]].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {header}\nkernel,launches,total_ns,share\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{v[0]},{v[1]:.0f},{v[1] / tot:.4f}\n")
    print("wrote", dst, f"({len(body)} launches, {tot / 1e6:.3f} ms)")


def full_summary(rep, dst):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        print("empty report", rep)
        return None
    hdr, units, launches = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write("metric,unit," + ",".join(f"launch{i}" for i in range(len(launches))) + "\n")
        f.write("Kernel Name,," + ",".join('"' + l[ix["Kernel Name"]][:100] + '"' for l in launches) + "\n")
        for m in KEEP:
            if m in ix:
                f.write(f"{m},{units[ix[m]]}," + ",".join(l[ix[m]].replace(",", "") for l in launches) + "\n")
    print("wrote", dst)
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    tr = []
    for l in launches:
        b = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(l[ix[m]].replace(",", "")) * scale.get(units[ix[m]], 1.0)
        tr.append(b)
    return tr


if __name__ == "__main__":
    tag = sys.argv[1]
    # ---- round-2 layout (scripts/gpu_final.sh): train-step and note-encoder commands profiled separately
    tcmd = "python bench.py --steps 2 --warmup 3 --skip-note-encoder --skip-eager --cpu-train-steps 0"
    ncmd = "python bench.py --config 2 --steps 2 --warmup 3 --cpu-chunks 0"
    p = os.path.join(OUT, f"launches_note_{tag}.csv")
    if not os.path.exists(p):
        p = os.path.join(OUT, f"launches_{tag}.csv")
    if os.path.exists(p):
        launch_summary(p, os.path.join(PROF, f"{tag}_launches_summary_note_encoder.csv"),
                       f"ncu --metrics gpu__time_duration.sum --clock-control none -c 1000: {ncmd}  (all launches of the run; "
                       "cold-cache serialised launches: shares, not absolute times)")
    p = os.path.join(OUT, f"launches_train_{tag}.csv")
    if os.path.exists(p):
        launch_summary(p, os.path.join(PROF, f"{tag}_launches_summary_train_step.csv"),
                       f"ncu --metrics gpu__time_duration.sum --clock-control none -c 6000: {tcmd}  (last step only; "
                       "cold-cache serialised launches: shares, not absolute times)", "clip_adamw")
    traffic_path = os.path.join(PROF, "roofline_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    jobs = (("gemm_train", "train_step:gemm_bf16_tcgen05_kernel", tcmd), ("gemm_note", "note_encoder:gemm_bf16_tcgen05_kernel", ncmd),
            ("attn_bwd", "train_step:attn_bwd_fused_kernel", tcmd), ("attn_fwd", "note_encoder:attn_fwd_pair_kernel", ncmd),
            ("gemm", "fame_gemm_bias_act", ncmd), ("attn", "fame_attn_fwd", ncmd), ("ln", "fame_layernorm", ncmd))
    for kern, key, cmd in jobs:
        rep = os.path.join(OUT, f"prof_{kern}_{tag}.ncu-rep")
        if os.path.exists(rep):
            tr = full_summary(rep, os.path.join(PROF, f"{tag}_ncu_{kern}_full_summary.csv"))
            if tr:
                traffic[key] = int(sum(tr) / len(tr))
                traffic[f"_source_{key}"] = (f"profiles/{tag}_ncu_{kern}_full_summary.csv: dram__bytes_read.sum + "
                                             f"dram__bytes_write.sum, mean over the {len(tr)} launches captured with "
                                             f"ncu --set full --clock-control none under: {cmd}")
    json.dump(traffic, open(traffic_path, "w"), indent=1)
    print("traffic", {k: v for k, v in traffic.items() if not k.startswith("_")})
