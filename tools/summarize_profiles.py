"""Turn the ncu outputs a GPU call left in gpurun_out/ into the small tracked summaries under profiles/.

  python tools/summarize_profiles.py launches gpurun_out/launches_r01.csv profiles/r01_launches   -> .csv (slim per-launch list) + .md (per-kernel shares)
  python tools/summarize_profiles.py full gpurun_out/prof_conv_tc.ncu-rep profiles/r01_conv_tc_full.md
"""
import collections, csv, io, re, subprocess, sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "smsp__cycles_active.avg", "sm__inst_executed.sum",
]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("b200::", "")


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    agg = collections.defaultdict(lambda: [0, 0.0])
    with open(dst + ".csv", "w") as f:
        f.write("id,kernel,grid,block,duration_ns\n")
        for r in rows:
            v = float(r["Metric Value"].replace(",", ""))
            v *= {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r["Metric Unit"], 1.0)
            k = short(r["Kernel Name"])
            f.write('%s,"%s","%s","%s",%d\n' % (r["ID"], k, r["Grid Size"], r["Block Size"], v))
            agg[k][0] += 1
            agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst + ".md", "w") as f:
        f.write("# ncu launch list summary (`--metrics gpu__time_duration.sum --clock-control none`), source %s\n\n" % src)
        f.write("%d launches, %.2f ms summed device time (cold-cache, serialised: compare shares)\n\n" % (len(rows), tot / 1e6))
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f%% |\n" % (k, v[0], v[1] / 1e6, 100 * v[1] / tot))


def full(rep, dst):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none, source %s\n\n" % rep)
        for r in rows[2:]:
            f.write("## `%s` grid %s block %s\n\n| metric | value | unit |\n|---|---:|---|\n" %
                    (short(r[hdr.index("Kernel Name")]), r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
            for m in FULL_METRICS:
                if m in hdr:
                    f.write("| %s | %s | %s |\n" % (m, r[hdr.index(m)], units[hdr.index(m)]))
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
