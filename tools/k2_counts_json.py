"""Turns the ncu captures of tools/r02_counts.sh (one full-size launch of the pair kernel per workload) into
profiles/k2_inst_per_pair.json and profiles/k2_dram_traffic.json, and copies the CSVs to profiles/.
Usage: python tools/k2_counts_json.py gpurun_out/r02m_counts_ "round 2 final K2" """
import csv, json, os, shutil, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icikendalltau_b200 import synth

prefix, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
inst_path, dram_path = os.path.join(root, "profiles", "k2_inst_per_pair.json"), os.path.join(root, "profiles", "k2_dram_traffic.json")
inst, dram = json.load(open(inst_path)), json.load(open(dram_path))
for wl, (n, C, *_ ) in synth.WORKLOADS.items():
    path = f"{prefix}{wl}.csv"
    if not os.path.exists(path):
        continue
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    by_id = {}
    for r in rows:
        by_id.setdefault(r[0], {})[r[12]] = float(r[14].replace(",", ""))
    # the launch that ran (the other tiers exit at once): the longest one
    best = max(by_id.values(), key=lambda m: m.get("gpu__time_duration.sum", 0.0))
    pairs = C * (C - 1) // 2
    inst[wl] = {"warp_inst_per_pair": round(best["smsp__inst_executed.sum"] / pairs),
                "source": f"ncu smsp__inst_executed.sum {int(best['smsp__inst_executed.sum'])} for the full-size launch of {pairs} pairs "
                          f"(tools/r02_counts.sh, {path}, {note})"}
    dram[wl] = best["dram__bytes_read.sum"] + best["dram__bytes_write.sum"]
    shutil.copy(path, os.path.join(root, "profiles", f"r02_k2_counts_{wl}.csv"))
    print(wl, inst[wl]["warp_inst_per_pair"], "warp inst/pair", dram[wl], "DRAM bytes", round(best["gpu__time_duration.sum"] / 1e6, 3), "ms under ncu")
json.dump(inst, open(inst_path, "w"), indent=1)
json.dump(dram, open(dram_path, "w"), indent=1)
