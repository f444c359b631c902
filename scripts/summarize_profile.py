"""Summarise ncu captures (gpurun_out/*.ncu-rep) into tracked text/JSON files under profiles/.
Usage: python scripts/summarize_profile.py <tag>   (expects the files written by scripts/profile_round.sh <tag>)"""
import csv, io, json, os, subprocess, sys, shutil

TAG = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(PR, exist_ok=True)
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    return d


def to_bytes(v, u):
    f = float(v)
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


traffic = {}
tpath = os.path.join(PR, "terrain_traffic.json")
if os.path.exists(tpath):
    traffic = json.load(open(tpath))
lines = [f"# ncu summaries, tag {TAG} (source: gpurun_out/*.ncu-rep of scripts/profile_round.sh {TAG}; --set full, --clock-control none)\n"]
for name, key in ((f"prof_{TAG}_florinsky4_32768", "florinsky_32768"), (f"prof_{TAG}_zt4_32768", "zevenbergthorne_32768")):
    rep = os.path.join(GO, name + ".ncu-rep")
    if not os.path.exists(rep):
        continue
    d = raw(rep)
    lines.append(f"\n## {name}  --  {d.get('Kernel Name', ('?',''))[0]}\n")
    for w in WANT:
        if w in d:
            lines.append(f"{w:90s} {d[w][0]:>20s} {d[w][1]}")
    rd, wr = to_bytes(*d["dram__bytes_read.sum"]), to_bytes(*d["dram__bytes_write.sum"])
    dur = float(d["gpu__time_duration.sum"][0]) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}[d["gpu__time_duration.sum"][1]]
    px = 32768 * 32768
    lines.append(f"-> DRAM traffic per launch {rd + wr:.4e} B (read {rd:.4e} + write {wr:.4e}); algorithmic 20 B/px x {px} px = {20*px:.4e} B; ratio {(rd+wr)/(20*px):.3f}")
    lines.append(f"-> under ncu (cold, serialised): {dur*1e3:.3f} ms, {(rd+wr)/dur/1e9:.0f} GB/s DRAM; thread-instructions/pixel = {float(d['smsp__inst_executed.sum'][0])*32/px:.1f}")
    traffic[key] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "source": f"profiles/ncu_{TAG}.txt ({name})"}
# K2 / K3 captures: one block per kernel launch in the report
for name in (f"prof_{TAG}_variogram", f"prof_{TAG}_nuthkaab"):
    rep = os.path.join(GO, name + ".ncu-rep")
    if not os.path.exists(rep):
        continue
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    lines.append(f"\n## {name}\n")
    short = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
             "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
             "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
             "launch__registers_per_thread", "launch__grid_size"]
    for r in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, r)}
        lines.append("kernel " + d.get("Kernel Name", ("?", ""))[0][:110])
        for w in short:
            if w in d:
                lines.append(f"    {w:80s} {d[w][0]:>18s} {d[w][1]}")
open(os.path.join(PR, f"ncu_{TAG}.txt"), "w").write("\n".join(lines) + "\n")
json.dump(traffic, open(tpath, "w"), indent=1)
for f in (f"launches_{TAG}.csv", f"bench_{TAG}.json", f"bench_{TAG}_zt.json", f"bench_{TAG}_horn.json",
          f"bench_{TAG}_reference.json", f"perf_probe_{TAG}.txt", f"perf_vario_nk_{TAG}.txt",
          f"bench_{TAG}_variogram.json", f"bench_{TAG}_nuthkaab.json"):
    src = os.path.join(GO, f)
    if os.path.exists(src):
        shutil.copy(src, os.path.join(PR, f))
# launch-list shares
src = os.path.join(GO, f"launches_{TAG}.csv")
if os.path.exists(src):
    rows = list(csv.reader(open(src)))
    hdr = None; tot = {}; n = {}
    for r in rows:
        if r and r[0] == "ID": hdr = r; continue
        if hdr and len(r) == len(hdr):
            k = r[hdr.index("Kernel Name")][:100]; v = float(r[hdr.index("Metric Value")])
            tot[k] = tot.get(k, 0) + v; n[k] = n.get(k, 0) + 1
    s = sum(tot.values())
    with open(os.path.join(PR, f"launch_shares_{TAG}.txt"), "w") as f:
        f.write(f"# device-time shares of every kernel of `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu` under ncu ({TAG}).\n"
                "# The torch kernels (randn / cumsum / fill) build the synthetic DEM BEFORE the timed region; inside the timed\n"
                "# region only xbt::terrain_fused_kernel launches (gpu_launches == steps).\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"{v/1e6:10.3f} ms {100*v/s:6.2f}% x{n[k]:3d} {k}\n")
print(open(os.path.join(PR, f"ncu_{TAG}.txt")).read())
