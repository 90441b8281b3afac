"""Turn the outputs of tools/gpu_round.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/.

    python tools/save_profiles.py r1e r01e

Writes profiles/<name>_bench_n1.json, _bench_reference_arm.json, _launches_{common,allsnp}_wave_K4096.csv,
_sweep_{common,allsnp}_K4096_ncu_full_summary.csv, _source_hotspots.txt and profiles/sweep_traffic.json."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, name = sys.argv[1], sys.argv[2]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__cycles_active.avg", "lts__t_sector_hit_rate.pct"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


shutil.copy(os.path.join(G, f"{tag}_bench.json"), os.path.join(P, f"{name}_bench_n1.json"))
shutil.copy(os.path.join(G, f"{tag}_bench_reference.json"), os.path.join(P, f"{name}_bench_reference_arm.json"))
shutil.copy(os.path.join(G, f"{tag}_launches.csv"), os.path.join(P, f"{name}_launches_common_wave_K4096.csv"))
shutil.copy(os.path.join(G, f"{tag}_launches_all.csv"), os.path.join(P, f"{name}_launches_allsnp_wave_K4096.csv"))
out = {}
for rep, kind, T in ((f"{tag}_sweep.ncu-rep", "common", 1000), (f"{tag}_sweep_all.ncu-rep", "allsnp", 3000)):
    path = os.path.join(G, rep)
    rows = list(csv.reader(ncu(path, "--page", "raw", "--csv").splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    with open(os.path.join(P, f"{name}_sweep_{kind}_K4096_ncu_full_summary.csv"), "w") as fh:
        w = csv.writer(fh)
        w.writerow(["metric", "unit", "launch0"])
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                w.writerow([k, units[i], r[i]])
    d = dict(zip(hdr, r))

    def val(k):
        v = float(d[k].replace(",", ""))
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(units[hdr.index(k)], 1)

    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    out[kind] = {"dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
                 "algorithmic_bytes_of_captured_launch": 148 * 8 * 4096 * (5 * 2 * T + 20000.0),
                 "duration_ms_under_ncu": float(d["gpu__time_duration.sum"]),
                 "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                 "smem_wavefront_pct": float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]),
                 "dram_pct": float(d["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"])}
    src = ncu(path, "--page", "source", "--csv", "--print-source", "cuda,sass")
    tmp = os.path.join("/tmp", f"{name}_{kind}_src.csv")
    open(tmp, "w").write(src)
    hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), tmp, "30"], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{name}_sweep_{kind}_K4096_source_hotspots.txt"), "w").write(hot)
mix = (126 * out["common"]["dram_bytes_per_launch"] + 42 * out["allsnp"]["dram_bytes_per_launch"]) / 168
json.dump({"kernel": "k_sweep<256,16,2,1>",
           "capture": f"profiles/{name}_sweep_{{common,allsnp}}_K4096_ncu_full_summary.csv (ncu --set full, tools/prof_sweep.py --K 4096 --jobs 148 --its 6 [--all-snps]; burn-in sweeps)",
           "dram_bytes_per_launch": mix,
           "note": "average over the launch mix of a benchmark step (126 common-SNP T=1000 launches, 42 all-SNP T=3000 launches)",
           "common": out["common"], "allsnp": out["allsnp"]}, open(os.path.join(P, "sweep_traffic.json"), "w"), indent=1)
b = json.load(open(os.path.join(P, f"{name}_bench_n1.json")))
print("value", b["value"], "e2e", b["e2e"]["value"], "cpu", b["cpu_baseline"]["value"], "roofline", b["roofline"]["achieved"], b["roofline"]["frac"],
      "share", b["roofline"]["share_of_step"], "avg launch ms", b["roofline"]["avg_launch_ms"])
print(json.dumps(out, indent=1))
