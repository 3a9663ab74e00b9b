"""Summarise .ncu-rep captures (and the per-launch csv) into small text files under profiles/."""
import csv, subprocess, sys, os, io, collections

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum"]


def rep(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        out.write(f"kernel: {r[hdr.index('Kernel Name')][:100]}\n")
        for w in WANT:
            if w in hdr:
                out.write(f"    {w:75s} {r[hdr.index(w)]} {units[hdr.index(w)]}\n")


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0][:60]
        ns = float(r[-1].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ns
    tot = sum(a[1] for a in agg.values())
    out.write(f"{'kernel':62s} {'launches':>8s} {'total_us':>10s} {'share':>7s}\n")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write(f"{k:62s} {n:8d} {ns/1e3:10.1f} {ns/tot:7.3f}\n")


if __name__ == "__main__":
    tag = sys.argv[1]
    src = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out"
    os.makedirs("profiles", exist_ok=True)
    with open(f"profiles/{tag}_ncu_summary.txt", "w") as out:
        out.write("# ncu --set full --clock-control none captures (python bench.py, T170 L40), per launch\n")
        for f in sorted(os.listdir(src)):
            if f.endswith(".ncu-rep"):
                rep(os.path.join(src, f), out)
    if os.path.exists(os.path.join(src, "launches.csv")):
        with open(f"profiles/{tag}_launches.txt", "w") as out:
            out.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n")
            launches(os.path.join(src, "launches.csv"), out)
