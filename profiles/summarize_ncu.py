"""Summaries of the ncu captures that bench.py's roofline numbers are checked against.

  python profiles/summarize_ncu.py launches gpurun_out/launches.csv  > profiles/rNN_launches_summary.txt
      (ncu --metrics gpu__time_duration.sum --clock-control none --csv: share of each kernel in the step)
  python profiles/summarize_ncu.py full gpurun_out/full.ncu-rep WINDOWS_PER_LAUNCH profiles/rNN_ncu_traffic.json > profiles/rNN_ncu_summary.txt
      (ncu --set full: pipe utilisation, DRAM traffic per launch and per window of every kernel)
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

SLOT_OF = [  # kernel-name pattern -> profile slot name (mcd_profile_slot_name); first match wins
    (r"TcCfg<(?:\(int\))?\d+, (?:\(int\))?17, (?:\(int\))?16, (?:\(int\))?32", "st_gcnnsd1.0"), (r"TcCfg<(?:\(int\))?\d+, (?:\(int\))?17, (?:\(int\))?32, (?:\(int\))?32", "st_gcnnsd1.1|st_gcnnsu3.0"),
    (r"TcCfg<(?:\(int\))?\d+, (?:\(int\))?12, (?:\(int\))?32, (?:\(int\))?64", "st_gcnnsd2.0"), (r"TcCfg<(?:\(int\))?\d+, (?:\(int\))?12, (?:\(int\))?64, (?:\(int\))?64", "st_gcnnsd2.1|st_gcnnsu4.0"),
    (r"TcCfg<(?:\(int\))?\d+, (?:\(int\))?10, (?:\(int\))?64, (?:\(int\))?128", "st_gcnnsd3.0"), (r"(?:Tc|Cf)Cfg<(?:\(int\))?\d+, (?:\(int\))?10, (?:\(int\))?128, (?:\(int\))?64", "st_gcnnsd3.1"),
    (r"(?:Tc|Cf)Cfg<(?:\(int\))?\d+, (?:\(int\))?12, (?:\(int\))?64, (?:\(int\))?32", "st_gcnnsu4.1"),
    (r"EdgeCfg<(?:\(int\))?\d+, (?:\(int\))?17, (?:\(int\))?\d+, (?:\(bool\))?1", "st_gcnnsp1a.0"), (r"EdgeCfg<(?:\(int\))?\d+, (?:\(int\))?17, (?:\(int\))?\d+, (?:\(bool\))?0", "st_gcnnsu3.1"),
    (r"joint_resample_kernel<(?:\(int\))?17, (?:\(int\))?12", "down1"), (r"joint_resample_kernel<(?:\(int\))?12, (?:\(int\))?10", "down2"),
    (r"joint_resample_kernel<(?:\(int\))?10, (?:\(int\))?12", "up3"), (r"joint_resample_kernel<(?:\(int\))?12, (?:\(int\))?17", "up2"),
]


def short(name: str) -> str:
    name = re.sub(r"\(int\)|\(bool\)|mcd::|void ", "", name)
    name = re.sub(r"\(BlockWeights, BlockIO\)|\(.*\)$", "", name)
    return name[:70]


def launches(path: str) -> None:
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ci = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows:
        if r is hdr or len(r) < len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        a = agg.setdefault(short(r[ci["Kernel Name"]]), [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * us / total:6.2f}%  {n:5d} launches  {us / n:9.1f} us/launch  {k}")


def full(rep: str, windows: int, traffic_json: str) -> None:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "sm__cycles_elapsed.avg"]
    traffic = {}
    for r in rows[2:]:
        name = r[ci["Kernel Name"]]
        print(short(name))
        for w in want:
            if w in ci:
                print(f"   {w:75s} {r[ci[w]]:>16s} {rows[1][ci[w]]}")

        def num(key):
            v = float(r[ci[key]].replace(",", ""))
            u = rows[1][ci[key]]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        if "dram__bytes_read.sum" in ci:
            per_window = (num("dram__bytes_read.sum") + num("dram__bytes_write.sum")) / windows
            for pat, slot in SLOT_OF:
                if re.search(pat, name):
                    for s in slot.split("|"):
                        traffic.setdefault(s, per_window)
                    break
    json.dump({"source": f"{rep} (ncu --set full --clock-control none, {windows} windows per launch): "
                         "(dram__bytes_read.sum + dram__bytes_write.sum) / windows", "dram_bytes_per_window": traffic},
              open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], int(sys.argv[3]), sys.argv[4])
