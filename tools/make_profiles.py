"""Turn the files tools/gpu_profiles.sh brought back (gpurun_out/) into the tracked summaries under profiles/:
the launch list with each kernel's share of a step, the headline ncu metrics of the four hot kernels, and
profiles/attn_traffic.json (DRAM bytes per launch of the attention kernels, read by bench.py)."""
import csv, collections, json, subprocess, sys, os

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
out = "profiles"
rows = [r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r) > 14 and r[0].isdigit()]
# steps = warmup 3 + timed 3; keep the launches of the last step: find the last 'fold_kernel' occurrence pattern
names = [r[4] for r in rows]
times = [float(r[14]) for r in rows]       # gpu__time_duration.sum, ns or us depending on unit column
unit = rows[0][13]
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(unit, 1e-6)
short = lambda n: n.split("(")[0].replace("void ", "").replace("spotv2::", "").replace("<unnamed>::", "")[:70]
# last step = launches after the last launch of the first kernel name of a step (fold_kernel)
idx = [i for i, n in enumerate(names) if "fold_all_kernel" in n] or \
      [i for i, n in enumerate(names) if "fold_copy_padded_kernel" in n] or \
      [i for i, n in enumerate(names) if "fold_kernel<8>" in n or "fold_kernel<(int)8>" in n]
first = idx[-1] if idx else 0                    # a step starts with the fold (one launch since r2c; W_aug copy | fold_kernel<8> before)
step = list(zip(names[first:], times[first:]))
agg = collections.OrderedDict()
for n, t in step:
    k = short(n)
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += t * scale
tot = sum(v[1] for v in agg.values())
with open(f"{out}/{tag}_launch_list_summary.txt", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-structured --no-graph\n")
    f.write("# launches of the LAST step (cold-cache, serialised: the SHARE is what carries over to the bench)\n")
    f.write(f"# {len(step)} launches, {tot:.3f} ms\n")
    for k, (c, t) in agg.items():
        f.write(f"{k:72s} x{c:<3d} {t:8.3f} ms  {100 * t / tot:5.1f}%\n")
with open(f"{out}/{tag}_launches.csv", "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "block", "grid", "gpu__time_duration", unit])
    for r in rows[first:]:
        w.writerow([short(r[4]), r[7], r[8], r[14], r[13]])

want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
traffic = {}
with open(f"{out}/{tag}_ncu_hot_kernels.txt", "w") as f:
    f.write("# ncu --set full --clock-control none (one launch each, warm-up skipped), python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-structured --no-graph\n")
    for rep in ("fwd_full", "bwd2_full", "gemm_full"):
        path = f"gpurun_out/{rep}.ncu-rep"
        if not os.path.exists(path):
            continue
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(txt.splitlines()))
        hdr, units = rr[0], rr[1]
        for vals in rr[2:]:
            d = dict(zip(hdr, vals))
            u = dict(zip(hdr, units))
            f.write("-----\nKernel Name = " + short(d.get("Kernel Name", "?")) + "\n")
            for k in want:
                if k in d:
                    f.write(f"{k} = {d[k]} {u[k]}\n")
            if rep in ("fwd_full", "bwd2_full"):
                conv = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                b = float(d["dram__bytes_read.sum"]) * conv[u["dram__bytes_read.sum"]] + \
                    float(d["dram__bytes_write.sum"]) * conv[u["dram__bytes_write.sum"]]
                traffic["fwd" if rep == "fwd_full" else "bwd"] = b
if traffic:
    traffic["sum"] = sum(traffic.values())
    traffic["note"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch at B=4096 (ncu --set full, one capture each)"
    json.dump(traffic, open(f"{out}/attn_traffic.json", "w"), indent=1)
print(open(f"{out}/{tag}_launch_list_summary.txt").read())
print(open(f"{out}/{tag}_ncu_hot_kernels.txt").read())
print(traffic)
