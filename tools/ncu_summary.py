"""Summarise ncu artefacts (run in the build container, no GPU needed):
  python tools/ncu_summary.py rep  <file.ncu-rep>      -> key metrics + top stall sites
  python tools/ncu_summary.py list <launches.csv>      -> per-kernel totals and shares
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]


def ncu(path, page):
    out = subprocess.run(["ncu", "-i", path, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def rep(path):
    rows = ncu(path, "raw")
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print(f"kernel: {d.get('Kernel Name')}")
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"  {k:86s} {d[k]:>16s} {u.get(k, '')}")
    rows = ncu(path, "source")
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
    agg = collections.Counter()
    for r in data:
        for k in hdr:
            if k.startswith("stall_") and "Not Issued" not in k:
                agg[k] += int(r[ix[k]] or 0)
    print(f"\nwarp-state samples: {tot} over {len(data)} SASS instructions")
    print("  " + ", ".join(f"{k}={v}" for k, v in agg.most_common(8)))
    print("top stall sites (samples, executions, SASS, dominant reasons):")
    data.sort(key=lambda r: -int(r[ix["# Samples"]] or 0))
    for r in data[:12]:
        st = {k: int(r[ix[k]] or 0) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
        top = ", ".join(f"{k}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2])
        print(f"  {r[ix['# Samples']]:>6s} {r[ix['Instructions Executed']]:>9s}  {r[ix['Source']][:70]:70s} {top}")


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ix["Kernel Name"]].split("(")[0][:90]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        v = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if "plnlp" in k)
    print(f"launches {sum(a[0] for a in agg.values())}, total {tot:.0f} us (serialised, cold cache: compare SHARES); "
          f"plnlp_b200 kernels {ours / tot * 100:.1f}% of device time")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"{a[1]:10.1f} us {a[1] / tot * 100:5.1f}% n={a[0]:4d} {k}")


if __name__ == "__main__":
    {"rep": rep, "list": launches}[sys.argv[1]](sys.argv[2])
