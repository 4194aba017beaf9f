#!/usr/bin/env python
"""Per-kernel constants for bench.py's roofline fields, taken from an `ncu --set full` capture:

    python tools/ncu_facts.py gpurun_out/prof.ncu-rep gpurun_out/plain.log [--out profiles/kernel_facts.json]

plain.log is the output of the SAME bench.py command run without ncu just before the capture (its JSON line gives the
beam evaluations per scan and the particle count, which turn the capture's per-launch totals into per-unit figures).
Writes, per kernel: warp instructions and DRAM bytes per launch, per beam evaluation (k_scanmatch) or per particle
(k_register), issue-slot utilisation, and where the numbers came from."""
import csv
import io
import json
import subprocess
import sys


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:]]


def main():
    rep, plain = sys.argv[1], sys.argv[2]
    out_path = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else "profiles/kernel_facts.json"
    line = [json.loads(l) for l in open(plain).read().splitlines() if l.startswith("{")][-1]
    n = line["config"]["particles"] // line["n_gpus"]
    beam_evals = line["beam_evals_per_scan"] / line["n_gpus"]
    facts = {}
    for r in raw_rows(rep):
        name = r["Kernel Name"]
        key = "k_scanmatch" if "k_scanmatch" in name else "k_register" if "k_register" in name else None
        if key is None or key in facts:
            continue
        inst = float(r["smsp__inst_executed.sum"])
        cycles = float(r["sm__cycles_elapsed.max"])
        unit = {"dram__bytes_read.sum": r.get("dram__bytes_read.sum", "0"), "dram__bytes_write.sum": r.get("dram__bytes_write.sum", "0")}
        # ncu prints these two in the unit of its header row (Mbyte / Gbyte ...): re-read them in bytes through --page raw units
        dram = None
        f = {"kernel": name, "source": "%s (ncu --set full --clock-control none, one launch; bench line of the same command: %s)" % (rep, plain),
             "duration_ms": float(r["gpu__time_duration.sum"]), "registers_per_thread": int(float(r["launch__registers_per_thread"])),
             "warp_instructions_per_launch": inst, "issue_slots_busy_pct": 100.0 * inst / (4 * 148 * cycles),
             "warps_active_pct": float(r["sm__warps_active.avg.pct_of_peak_sustained_active"]),
             "l1_hit_pct": float(r["l1tex__t_sector_hit_rate.pct"]), "l2_hit_pct": float(r["lts__t_sector_hit_rate.pct"]),
             "fp64_pipe_pct": float(r["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]), "particles_in_launch": n}
        facts[key] = f
    # DRAM bytes with explicit units
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        key = "k_scanmatch" if "k_scanmatch" in d["Kernel Name"] else "k_register" if "k_register" in d["Kernel Name"] else None
        if key and "dram_bytes_per_launch" not in facts[key]:
            facts[key]["dram_bytes_per_launch"] = float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"])
    if "k_scanmatch" in facts:
        f = facts["k_scanmatch"]
        f["beam_evals_in_launch"] = beam_evals
        f["warp_instructions_per_beam_eval"] = f["warp_instructions_per_launch"] * 32.0 / beam_evals  # per thread-level beam evaluation
        f["dram_bytes_per_beam_eval"] = f["dram_bytes_per_launch"] / beam_evals
    if "k_register" in facts:
        f = facts["k_register"]
        f["dram_bytes_per_particle"] = f["dram_bytes_per_launch"] / n
        f["warp_instructions_per_particle"] = f["warp_instructions_per_launch"] / n
    json.dump(facts, open(out_path, "w"), indent=1)
    print(json.dumps(facts, indent=1))


if __name__ == "__main__":
    main()
