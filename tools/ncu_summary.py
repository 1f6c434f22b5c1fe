#!/usr/bin/env python
"""Turns an ncu capture (`ncu --set full ... -o X`) into the small per-kernel JSON summaries under profiles/ that bench.py folds
into its `roofline` object (run here, where ncu can read reports without a GPU):

    python tools/ncu_summary.py gpurun_out/r2_trace_prof.ncu-rep --match 'k_trace_persistent<0' --out profiles/r2_extrays_ncu_summary.json
"""
import argparse
import csv
import io
import json
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--match", required=True, help="substring of the kernel name (first matching launch is summarised)")
    ap.add_argument("--out", required=True)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    name_i = hdr.index("Kernel Name")
    row = next(r for r in rows[2:] if a.match in r[name_i])

    def val(key, default=None):
        if key not in hdr:
            return default
        try:
            return float(row[hdr.index(key)].replace(",", ""))
        except ValueError:
            return default

    def unit(key):
        return rows[1][hdr.index(key)] if key in hdr else ""

    def bytes_of(key):
        v, u = val(key, 0.0), unit(key).lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    cycles = val("sm__cycles_elapsed.avg")
    sms = val("launch__sm_count") or 148
    wavefronts = val("l1tex__data_pipe_lsu_wavefronts.sum")
    pipe_pct = val("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")
    if wavefronts is None and pipe_pct is not None and cycles:  # the pipe moves one wavefront per clock per SM at its peak
        wavefronts = pipe_pct / 100.0 * cycles * sms
    out = {
        "kernel": row[name_i],
        "source": "%s (ncu --set full --clock-control none; serialised, cold-cache replays)" % a.report.split("/")[-1],
        "duration_us": val("gpu__time_duration.sum"),
        "dram_bytes_read": bytes_of("dram__bytes_read.sum"),
        "dram_bytes_write": bytes_of("dram__bytes_write.sum"),
        "dram_bytes_per_launch": bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum"),
        "dram_throughput_pct": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "threads_per_instruction": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "warp_instructions": val("smsp__inst_executed.sum"),
        "issue_slot_utilization_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "l1tex_data_pipe_pct": val("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "l1tex_wavefronts": wavefronts,
        "l1tex_wavefronts_per_clk_per_sm": round(wavefronts / (cycles * sms), 4) if wavefronts and cycles else None,
        "l1_hit_pct": val("l1tex__t_sector_hit_rate.pct"),
        "l2_hit_pct": val("lts__t_sector_hit_rate.pct"),
        "registers": val("launch__registers_per_thread"),
        "achieved_occupancy_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "local_memory_requests": (val("l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", 0) or 0) + (val("l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", 0) or 0),
        "sm_cycles": cycles,
    }
    if a.note:
        out["note"] = a.note
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
