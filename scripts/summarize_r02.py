"""Copy the round-2 evidence written by scripts/profile_r02_final.sh <tag> from gpurun_out/ into profiles/ (tracked):
ncu brief of every captured kernel, launch list + device-time shares of the bench command, bench lines, and the DRAM
traffic of the headline kernel for bench.py's roofline.traffic.   Usage: python scripts/summarize_r02.py r02f"""
import csv, io, json, os, shutil, subprocess, sys

TAG = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}

for f in (f"ncu_{TAG}.txt", f"launches_{TAG}.csv", f"bench_{TAG}.json", f"bench_{TAG}_reference.json", f"perf_probe_{TAG}.txt"):
    if os.path.exists(os.path.join(GO, f)):
        shutil.copy(os.path.join(GO, f), os.path.join(PR, f))

rep = os.path.join(GO, f"prof_{TAG}_florinsky4_32768.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    d = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}
    rd = float(d["dram__bytes_read.sum"][0]) * UNIT[d["dram__bytes_read.sum"][1]]
    wr = float(d["dram__bytes_write.sum"][0]) * UNIT[d["dram__bytes_write.sum"][1]]
    tpath = os.path.join(PR, "terrain_traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    traffic["florinsky_32768"] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                                  "source": f"profiles/ncu_{TAG}.txt (prof_{TAG}_florinsky4_32768, ncu --set full)"}
    json.dump(traffic, open(tpath, "w"), indent=1)
    print("headline traffic", rd + wr, "ratio to 20 B/px", (rd + wr) / (20 * 32768 * 32768))

src = os.path.join(GO, f"launches_{TAG}.csv")
if os.path.exists(src):
    rows = list(csv.reader(open(src)))
    hdr = None; tot = {}; n = {}
    for r in rows:
        if r and r[0] == "ID": hdr = r; continue
        if hdr and len(r) == len(hdr):
            k = r[hdr.index("Kernel Name")][:110]; v = float(r[hdr.index("Metric Value")].replace(",", ""))
            u = r[hdr.index("Metric Unit")]
            v *= {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(u, 1)
            tot[k] = tot.get(k, 0) + v; n[k] = n.get(k, 0) + 1
    s = sum(tot.values())
    with open(os.path.join(PR, f"launch_shares_{TAG}.txt"), "w") as f:
        f.write(f"# device-time shares of the first 600 kernel launches of `python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1`\n"
                f"# under ncu ({TAG}; cold cache, serialised: compare shares).  torch kernels (randn / cumsum / fill / copy) build the\n"
                "# synthetic inputs BEFORE the timed regions; inside the headline's timed steps only\n"
                "# xbt::florinsky_sliding_kernel launches (gpu_launches == steps), the e2e leg launches it once per row block.\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"{v/1e6:10.3f} ms {100*v/s:6.2f}% x{n[k]:4d} {k}\n")
    print(open(os.path.join(PR, f"launch_shares_{TAG}.txt")).read()[:3000])
