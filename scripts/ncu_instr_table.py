"""profiles/instr_table.json from ncu captures of the dominant kernels (run here, on the CPU box, on what gpurun brought back).

    ncu -i gpurun_out/prof_<wl>.ncu-rep --page raw --csv > gpurun_out/prof_<wl>.csv     (done by this script)
    python scripts/ncu_instr_table.py <wl>:<kernel regex>:<skipped launches> ...  [--commit HASH]

For every workload: executed thread-instructions of the captured launch / units that launch processed
(gpurun_out/ncu_units_<wl>.json, written by scripts/ncu_target.py), issue-slot and pipe utilisation, DRAM bytes of the launch."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PIPES = {"fmaheavy (IMAD.WIDE of Philox)": "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
         "fma": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
         "alu": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
         "xu (MUFU)": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
         "fp64": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
         "lsu": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"}


def f(x):
    return float(x.replace(",", "")) if x not in ("", "n/a") else None


def main():
    commit = sys.argv[sys.argv.index("--commit") + 1] if "--commit" in sys.argv else \
        subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    path = os.path.join(ROOT, "profiles", "instr_table.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    for spec in [a for a in sys.argv[1:] if ":" in a]:
        wl, kre, skipped = spec.split(":")
        rep = os.path.join(ROOT, "gpurun_out", f"prof_{wl}.ncu-rep")
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, unit_row = rows[0], rows[1]
        ki = hdr.index("Kernel Name")
        row = next(r for r in rows[2:] if re.search(kre, r[ki]))
        g = lambda name: f(row[hdr.index(name)]) if name in hdr else None  # noqa: E731
        units = json.load(open(os.path.join(ROOT, "gpurun_out", f"ncu_units_{wl}.json")))
        it = units["units"][int(skipped) // units["units"][0]["launches_per_step"]]
        n_units = it["events"] if wl == "lv_smc" else it["evals"] / it["launches_per_step"]
        tinst = g("smsp__thread_inst_executed.sum") or g("sass__thread_inst_executed_true_per_opcode") or \
            (g("smsp__inst_executed.sum") or 0) * 32.0
        dr, dw = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        dru, dwu = unit_row[hdr.index("dram__bytes_read.sum")], unit_row[hdr.index("dram__bytes_write.sum")]
        pipes = {k: g(v) for k, v in PIPES.items()}
        top = max((v, k) for k, v in pipes.items() if v is not None)
        table[f"{wl}/{units['precision']}"] = {
            "kernel": row[ki][:120], "thread_inst_per_unit": tinst / n_units, "units_in_launch": n_units,
            "thread_inst_executed": tinst, "warp_inst_executed": g("smsp__inst_executed.sum"),
            "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "binding_pipe": {"name": top[1], "pct": top[0]}, "pipes_pct": pipes,
            "registers": g("launch__registers_per_thread"), "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "duration_us_under_ncu": (g("gpu__time_duration.sum") or 0.0) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(
                unit_row[hdr.index("gpu__time_duration.sum")], 1.0),
            "dram_bytes_per_launch": dr * scale.get(dru, 1.0) + dw * scale.get(dwu, 1.0),
            "source": f"ncu --set full, launch {int(skipped) + 1} of {kre} (smsp__thread_inst_executed.sum / units of that launch), commit {commit}",
            "commit": commit}
        print(wl, json.dumps(table[f"{wl}/{units['precision']}"], indent=1))
    json.dump(table, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
