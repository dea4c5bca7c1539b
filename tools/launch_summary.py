"""Aggregate an ncu launch list (csv from `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]
--clock-control none --csv --log-file X python bench.py --no-graph ...`) into a per-kernel table of ONE training step
(the last complete one: launches after the previous step's AdamW up to this step's AdamW).
Usage: python tools/launch_summary.py launches.csv > profiles/rN_launch_summary.txt"""
import collections
import csv
import re
import sys


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    per = collections.OrderedDict()
    for r in rd:
        k = int(r["ID"])
        e = per.setdefault(k, {"name": r["Kernel Name"]})
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"].lower()
        m = r["Metric Name"]
        if m.startswith("gpu__time_duration"):
            e["us"] = val * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1.0)
        elif m.startswith("dram__bytes"):
            scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
            e["rd" if "read" in m else "wr"] = val * scale
    launches = list(per.values())
    ends = [i for i, e in enumerate(launches) if "adamw" in e["name"]]
    # AdamW launches come in runs (one per optimiser group): a step ends at the last launch of a run
    step_ends = [i for j, i in enumerate(ends) if j + 1 == len(ends) or ends[j + 1] != i + 1]
    if len(step_ends) >= 2:
        lo, hi = step_ends[-2] + 1, step_ends[-1] + 1
    else:
        lo, hi = 0, len(launches)
    step = launches[lo:hi]
    agg = collections.OrderedDict()
    for e in step:
        name = re.sub(r"^void ", "", e["name"])
        name = re.sub(r"cgpt::\(anonymous namespace\)::|cgpt::<unnamed>::|cgpt::", "", name)
        name = re.sub(r"\(.*$", "", name)
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += e.get("us", 0.0)
        a[2] += e.get("rd", 0.0)
        a[3] += e.get("wr", 0.0)
    tot = sum(a[1] for a in agg.values())
    has_dram = any(a[2] or a[3] for a in agg.values())
    print(f"# one eager training step: {len(step)} launches, sum of kernel durations {tot / 1e3:.2f} ms "
          "(serialised, cold-cache per-launch times: compare shares, not absolutes)")
    hdr = f"{'kernel':<84}{'n':>5}{'ms':>9}{'share':>7}"
    if has_dram:
        hdr += f"{'dram rd MB':>12}{'dram wr MB':>12}"
    print(hdr)
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        line = f"{name[:82]:<84}{a[0]:>5}{a[1] / 1e3:>9.3f}{100 * a[1] / tot:>6.1f}%"
        if has_dram:
            line += f"{a[2] / 1e6:>12.1f}{a[3] / 1e6:>12.1f}"
        print(line)
    fam = [a for n, a in agg.items() if n.startswith("gemm_bf16_kernel")]
    if fam:
        n = sum(a[0] for a in fam)
        print(f"# gemm_bf16_kernel family: {n} launches, {sum(a[1] for a in fam) / 1e3:.3f} ms "
              f"({100 * sum(a[1] for a in fam) / tot:.1f}% of the step)"
              + (f", DRAM traffic per launch {sum(a[2] + a[3] for a in fam) / n / 1e6:.1f} MB" if has_dram else ""))


if __name__ == "__main__":
    main(sys.argv[1])
