"""Summarises an ncu launch list (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv)
of `bench.py --steps 1 --warmup 0`: per kernel the launches, device time and DRAM bytes; for the FIRST solve of the
command (one step = one 65 536-pose batch) the generation kernels' share of the step and their DRAM traffic.
usage: python profiles/launches_summary.py gpurun_out/r02_launches.csv [traffic.json]"""
import csv
import json
import re
import sys
from collections import defaultdict


def short(name):
    for key in ("memetic_generation_kernel", "memetic_init_kernel", "gd_local_kernel", "eval_cost_kernel", "species_pick_kernel",
                "pack_results_kernel", "sm_discover_kernel", "fp64_peak_kernel"):
        if key in name:
            # the Wide template flag: StaticSpec<n, kinds, tip, WIDE, ...>, PatternSpec<o, t, WIDE>, TreeSpec<WIDE>, GenericSpecW
            m = (re.search(r"StaticSpec<\d+, \d+, \w+, (\w+)", name) or re.search(r"PatternSpec<\d+, \d+, (\w+)", name)
                 or re.search(r"TreeSpec<(\w+)", name))
            is_wide = (m and m.group(1) in ("1", "true", "(bool)1")) or "Lb1ELb1E" in name or "GenericSpec" in name
            wide = "wide" if is_wide else "throughput"
            return key + (" [" + wide + "]" if key == "memetic_generation_kernel" else "")
    return "(library) " + name[:60]


def main(path, out_json=None):
    rows = list(csv.reader(line for line in open(path) if not line.startswith("==")))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    per = defaultdict(lambda: defaultdict(float))  # launch id -> metric -> value
    names = {}
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        lid = int(r[col["ID"]])
        names[lid] = r[col["Kernel Name"]]
        val = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3, "byte": 1.0, "Kbyte": 1e3,
                 "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        per[lid][r[col["Metric Name"]]] = val * scale
    ids = sorted(per)
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for lid in ids:
        k = short(names[lid])
        agg[k][0] += 1
        agg[k][1] += per[lid]["gpu__time_duration.sum"]
        agg[k][2] += per[lid]["dram__bytes_read.sum"] + per[lid]["dram__bytes_write.sum"]
    total = sum(v[1] for v in agg.values())
    print(f"{len(ids)} launches, {total:.2f} ms of device time under ncu (serialised, cold caches)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:45s} {v[0]:5d} launches  {v[1]:9.3f} ms  {100 * v[1] / total:5.1f} %  dram {v[2] / 1e9:8.3f} GB")
    # the first solve: from the first memetic_init_kernel to the launch before the second one
    inits = [lid for lid in ids if "memetic_init_kernel" in names[lid]]
    if len(inits) >= 1:
        lo = inits[0]
        hi = inits[1] if len(inits) > 1 else ids[-1] + 1
        step = [lid for lid in ids if lo <= lid < hi]
        gen = [lid for lid in step if "memetic_generation_kernel" in names[lid]]
        t_step = sum(per[l]["gpu__time_duration.sum"] for l in step)
        t_gen = sum(per[l]["gpu__time_duration.sum"] for l in gen)
        dram = sum(per[l]["dram__bytes_read.sum"] + per[l]["dram__bytes_write.sum"] for l in gen)
        rd = sum(per[l]["dram__bytes_read.sum"] for l in gen)
        busy = [l for l in gen if per[l]["gpu__time_duration.sum"] > 0.02]
        print(f"first solve: {len(step)} launches, {t_step:.2f} ms; generation kernels {len(gen)} launches ({len(busy)} with work), "
              f"{t_gen:.2f} ms = {100 * t_gen / t_step:.1f} % of the step; DRAM {dram / 1e9:.3f} GB ({rd / 1e9:.3f} read)")
        if out_json:
            json.dump({"source": path, "kernel": "memetic_generation_kernel (both flavours), all launches of one 65 536-pose batch",
                       "dram_bytes_per_step": dram, "dram_bytes_read_per_step": rd, "generation_launches_per_step": len(gen),
                       "generation_launches_with_work": len(busy), "generation_kernel_ms_under_ncu": t_gen,
                       "kernel_share_of_step_under_ncu": t_gen / t_step,
                       "note": "sum of dram__bytes_read.sum + dram__bytes_write.sum over every generation-kernel launch of one step "
                               "(ncu, serialised); BELOW the algorithmic figure (2 * P * (2n+2) * 8 B per problem-generation): only "
                               "the E elites are read back; the path is FP64-issue-bound, not HBM-bound"}, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:])
