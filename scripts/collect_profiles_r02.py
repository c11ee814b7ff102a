"""Round 2: turn the ncu captures of scripts/gpu_profiles.sh (gpurun_out/, scratch) into the tracked summaries under profiles/.
    python scripts/collect_profiles_r02.py"""
import csv, collections, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

def summary(rep, out, title):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), os.path.join(G, rep),
                          "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum", "launch__cluster", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
                          "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__inst_executed_pipe_lsu.avg.pct"], capture_output=True, text=True).stdout
    # per-segment instruction / stall-sample shares from the source page
    src = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    seg_txt = ""
    try:
        hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
        hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
        iS, iN, iSm, iT = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        tot = sum(int(r[iN]) for r in data); totS = sum(int(r[iSm]) for r in data)
        prev = None; segs = []
        for k, r in enumerate(data):
            n = int(r[iN])
            if prev is None: prev, start, acc, accs, acct = n, k, 0, 0, 0
            if abs(n - prev) > 0.25 * max(prev, 1) and n != prev:
                segs.append((start, k - 1, prev, acc, accs, acct)); start, acc, accs, acct = k, 0, 0, 0
            prev = n; acc += n; accs += int(r[iSm]); acct += int(r[iT])
        segs.append((start, len(data) - 1, prev, acc, accs, acct))
        seg_txt = f"\n---- code regions (consecutive SASS instructions with the same execution count), {tot} warp-instructions, {totS} stall samples\n"
        for (a, b, cnt, acc, accs, acct) in segs:
            if acc > 0.01 * tot or accs > 0.02 * totS:
                seg_txt += (f"  SASS [{a:5d}-{b:5d}] {b - a + 1:4d} instr x {cnt:9d} executions = {100 * acc / tot:5.1f}% of instructions, "
                            f"{100 * accs / totS:5.1f}% of stall samples, {acct / max(acc, 1):4.1f} active lanes; first: {data[a][iS].strip()[:48]}\n")
        cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
        ci = [hdr.index(c) for c in cols]
        top = sorted(range(len(data)), key=lambda k: -int(data[k][iSm]))[:12]
        seg_txt += "---- the 12 instructions with the most stall samples\n"
        for k in sorted(top):
            r = data[k]; st = sorted(((c[6:], int(r[i])) for c, i in zip(cols, ci) if int(r[i]) > 0), key=lambda x: -x[1])[:2]
            seg_txt += f"  SASS {k:5d} {100 * int(r[iSm]) / totS:5.2f}%  {r[iS].strip()[:64]:64s} {st}\n"
    except StopIteration:
        pass
    open(os.path.join(P, out), "w").write(f"# {title}\n# ncu --set full --clock-control none --import-source on (scripts/gpu_profiles.sh); B200, one launch\n\n" + txt + seg_txt)
    print("wrote", out)

def main():
    summary("r02_icp_hyp.ncu-rep", "r02_ncu_icp_hyp.txt", "icp_hyp_kernel<PackedScene>: C2, 512 hypotheses x 31 passes, cluster size 2 (scripts/time_icp.py)")
    summary("r02_raster_tile.ncu-rep", "r02_ncu_raster_tile.txt", "raster_tile_kernel: C2, 512 poses of obj_06, clustered path (scripts/time_step.py)")
    summary("r02_icp_nn.ncu-rep", "r02_ncu_icp_nn.txt", "icp_hyp_kernel<PackedNnScene>: C3, 512 hypotheses x 31 passes against the 99k-point kd-tree (scripts/time_nn.py)")

    # ---- launch list of the bench command
    rows = list(csv.reader(open(os.path.join(G, "r02_launches_bench.csv"), errors="ignore")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum": continue
        v = float(r[ix["Metric Value"]].replace(",", "")); u = r[ix["Metric Unit"]]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        launches.append((r[ix["Kernel Name"]].split("(")[0].replace("void ", ""), v))
    # the refiner steps: vertex_kernel ... icp_hyp_kernel groups; take the LAST complete step before the stage timings
    names = [n for n, _ in launches]
    step_end = max(i for i, n in enumerate(names) if "icp_hyp_kernel" in n and i > 1 and "icp_order_kernel" in names[i - 1] and "cloud_fill_tiles" in names[i - 2])
    step_start = max(i for i in range(step_end) if "vertex_kernel" in names[i])
    step = launches[step_start:step_end + 1]
    tot = sum(v for _, v in step)
    with open(os.path.join(P, "r02_launches_bench.md"), "w") as f:
        f.write("# Round 2 -- launch list of `python bench.py --steps 2 --warmup 3 --no-cpu --no-configs`\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` (cold-cache, serialised: compare SHARES, not absolutes).\n\n"
                "## one refiner step (pr_refiner_run_device), in launch order\n\n| kernel | us | share of the step |\n|---|---:|---:|\n")
        for n, v in step: f.write(f"| {n} | {v:.1f} | {100 * v / tot:.1f}% |\n")
        live = ""
        try:
            import json
            bj = json.loads(open(os.path.join(P, "r02_bench_1gpu.json")).read().strip().splitlines()[-1])
            live = (f"  Live in bench.py (CUDA events inside the timed steps): ICP {bj['stages']['icp_ms']:.2f} ms of a {bj['ms_per_step']:.2f} ms step = "
                    f"{100 * bj['stages']['icp_ms'] / bj['ms_per_step']:.0f} %; here {100 * step[-1][1] / tot:.1f} %.")
        except Exception:
            pass
        f.write(f"\nTotal {tot:.1f} us in {len(step)} launches (+ 1 memset).{live}\n\n## all launches of the capture, aggregated\n\n| kernel | launches | total us | mean us |\n|---|---:|---:|---:|\n")
        agg = collections.OrderedDict()
        for n, v in launches:
            a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
        for k, a in agg.items(): f.write(f"| {k} | {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.1f} |\n")
    print(open(os.path.join(P, "r02_launches_bench.md")).read()[:1800])


if __name__ == "__main__":
    if len(sys.argv) == 4:      # python scripts/collect_profiles_r02.py <rep in gpurun_out> <out file in profiles/> <title>
        summary(sys.argv[1], sys.argv[2], sys.argv[3])
    else:
        main()
