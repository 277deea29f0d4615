"""Summarise an .ncu-rep (ncu -i ... --page raw --csv) into a small JSON for profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--extra key=value ...] > profiles/x.json
"""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "gpu_time_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_throughput_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefronts_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct_of_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_memory_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic",
    "sm__cycles_elapsed.avg": "sm_cycles",
}


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")], "source": "ncu --set full --clock-control none: " + rep}
        for i, h in enumerate(hdr):
            if h in KEYS:
                try:
                    v = float(vals[i].replace(",", ""))
                except ValueError:
                    continue
                if units[i] == "byte":
                    v /= 1e6
                if units[i] == "ns":
                    v /= 1e3
                if units[i] == "ms":
                    v *= 1e3
                d[KEYS[h]] = round(v, 3)
        res.append(d)
    extra = dict(a.split("=", 1) for a in sys.argv[2:] if "=" in a)
    print(json.dumps({"launches": res, **extra}, indent=1))


if __name__ == "__main__":
    main()
