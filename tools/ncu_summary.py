"""Condense `ncu --set full` captures (.ncu-rep) into the small JSON summaries kept under profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep "what this capture is" > profiles/r02_x_ncu_full.json

Reads the report with `ncu -i ... --page raw --csv` (ncu is in the build container; no GPU needed) and keeps, per kernel
launch in the report: duration, DRAM bytes, pipe utilisation (FMA / ALU / FP64 / XU / LSU / tensor), issue slots, occupancy,
registers, launch shape, executed instructions, threads per instruction and the stall reasons per issue."""
import csv
import io
import json
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "duration_ms",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "pipe_fma_cycles_active_pct",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "pipe_alu_cycles_active_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "pipe_fp64_cycles_active_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "pipe_fp64_inst_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_inst_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_inst_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "pipe_tensor_cycles_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_per_block",
    "launch__occupancy_limit_registers": "occupancy_limit_registers_ctas",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_shared_mem_ctas",
    "smsp__inst_executed.sum": "warp_instructions_executed",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_instruction",
    "smsp__warps_eligible.avg.per_cycle_active": "eligible_warps_per_cycle",
}


def main():
    rep, what = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = {"what": what, "report": rep.split("/")[-1], "kernels": []}
    for r in data:
        k = {"kernel": r[hdr.index("Kernel Name")]}
        stalls = {}
        for h, un, v in zip(hdr, units, r):
            if h in KEEP:
                k[KEEP[h]] = f"{v} {un}".strip()
            elif h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                name = h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
                if float(v) >= 0.05:
                    stalls[name] = round(float(v), 3)
        k["stalls_per_issue_active"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
        out["kernels"].append(k)
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
