"""Summarise an .ncu-rep (one row per captured launch) into the handful of metrics the roofline discussion needs.
    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [out.json]"""
import csv, json, subprocess, sys

WANT = {
    "gpu__time_duration.sum": "time_us", "dram__bytes_read.sum": "dram_read_MB", "dram__bytes_write.sum": "dram_write_MB",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "lts__t_bytes.sum": "l2_bytes_MB", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__pipe_tensor_op_umma_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_umma_pct",
    "sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_umma_subpipe_pct",
    "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active": "tensor_hmma_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs", "launch__grid_size": "grid", "launch__block_size": "block",
    "launch__cluster_size": "cluster", "launch__shared_mem_per_block_dynamic": "smem_dyn_KB",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts", "smsp__inst_executed.sum": "warp_inst",
}


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:100]}
        for h, u, v in zip(hdr, units, r):
            if h in WANT:
                try:
                    x = float(v.replace(",", ""))
                except ValueError:
                    continue
                if u == "byte": x /= 1e6
                if u == "Kbyte": x /= 1e3 if "KB" not in WANT[h] else 1
                if u == "Gbyte": x *= 1e3
                if u == "ns": x /= 1e3
                if u == "ms": x *= 1e3
                d[WANT[h]] = round(x, 3)
        for h, v in zip(hdr, r):     # tensor-pipe activity under whatever name this ncu version reports it
            if h in ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
                     "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum",
                     "smsp__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_tensor.sum"):
                try:
                    d[h] = round(float(v.replace(",", "")), 2)
                except ValueError:
                    pass
        res.append(d)
    js = json.dumps(res, indent=1)
    print(js)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(js + "\n")


if __name__ == "__main__":
    main()
