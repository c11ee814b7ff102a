"""Copy the judged artefacts of the last scripts/gpu_full.sh run from gpurun_out/ (scratch) into profiles/ (tracked).
    python scripts/collect_profiles.py [round-prefix, default r01]"""
import csv, collections, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
R = sys.argv[1] if len(sys.argv) > 1 else "r01"

for src, dst in [("bench.json", f"{R}_bench_1gpu.json"), ("bench_ref.json", f"{R}_bench_reference_arm.json"),
                 ("ref_cuda.json", f"{R}_ref_cuda_build.json")]:
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))

# ---- launch list
rows = list(csv.reader(open(os.path.join(G, "launches_bench.csv"), errors="ignore")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum": continue
    v = float(r[ix["Metric Value"]].replace(",", "")); u = r[ix["Metric Unit"]]
    v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
    a = agg.setdefault(r[ix["Kernel Name"]].split("(")[0].replace("void ", ""), [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values()); n = sum(a[0] for a in agg.values())
with open(os.path.join(P, f"{R}_launches_bench.md"), "w") as f:
    f.write(f"# Round 1 -- launch list of `python bench.py --steps 2 --warmup 3 --no-cpu`\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none -s 25 -c 40` (cold-cache, serialised: compare shares).\n\n"
            "| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|\n")
    for k, a in agg.items(): f.write(f"| {k} | {a[0]} | {a[1]:.1f} | {a[1]/a[0]:.1f} | {a[1]/tot*100:.1f}% |\n")
    f.write(f"\nTotal {tot:.1f} us over {n} launches.  One refiner step = vertex, cluster span, bin scan, cluster fill, raster tile, tile scan, "
            "cloud offsets, cloud fill (tiles), scene pack, ICP plan, ICP persistent (+ 1 memset); the capture window also holds the "
            "per-stage timing calls of bench.py, which go through the public depth2cloud entry points (cloud_count / cloud_scan / cloud_fill).\n")

# ---- ncu summary + traffic of the ICP launch
rep = os.path.join(G, "step_kernels.ncu-rep")
txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
open(os.path.join(P, f"{R}_ncu_step_kernels.txt"), "w").write(txt)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
for r in rr[2:]:
    if "icp_persistent" in r[h.index("Kernel Name")]:
        def val(name):
            i = next(j for j, x in enumerate(h) if x.endswith(name)); v = float(r[i].replace(",", "")); u = rr[1][i]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        b = json.load(open(os.path.join(G, "bench.json")))
        json.dump({"kernel": "icp_persistent_kernel<PackedScene>", "source": f"ncu --set full --clock-control none, one launch of the C2 batch (profiles/{R}_ncu_step_kernels.txt)",
                   "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes_per_launch": rd + wr,
                   "algorithmic_bytes_per_launch": b["roofline"]["algorithmic_bytes_per_launch"]},
                  open(os.path.join(P, f"{R}_icp_traffic.json"), "w"), indent=1)
        break
print(open(os.path.join(P, f"{R}_launches_bench.md")).read())
print(open(os.path.join(P, f"{R}_icp_traffic.json")).read())
