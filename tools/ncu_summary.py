"""Summarise an .ncu-rep (read here with `ncu -i`, no GPU needed) into a small JSON + text table for profiles/.
Usage: python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep profiles/r01_ncu_X"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed": "lsu_wavefront_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_KB",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__cycles_elapsed.max": "cycles",
}


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {WANT[h]: i for i, h in enumerate(hdr) if h in WANT}
    iname = hdr.index("Kernel Name")
    res = []
    for r in data:
        d = {"kernel": r[iname]}
        for k, i in idx.items():
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                v = r[i]
            u = units[i]
            if k == "time_us" and u == "ns":
                v /= 1e3
            if k.endswith("_MB") and u == "byte":
                v /= 1e6
            if k.endswith("_MB") and u == "Kbyte":
                v /= 1e3
            if k.endswith("_MB") and u == "Gbyte":
                v *= 1e3
            d[k] = v
        res.append(d)
    json.dump(res, open(out + ".json", "w"), indent=1)
    cols = ["time_us", "tensor_pipe_active_pct", "dram_read_MB", "dram_write_MB", "dram_pct", "l2_pct", "lsu_wavefront_pct", "regs", "grid", "dyn_smem_KB"]
    with open(out + ".txt", "w") as f:
        f.write(f"# {rep}: one row per captured launch (ncu --set full --clock-control none; cold-cache, serialised)\n")
        f.write(" | ".join(["kernel"] + cols) + "\n")
        for d in res:
            f.write(" | ".join([d["kernel"][-70:]] + [f"{d.get(c, float('nan')):.2f}" if isinstance(d.get(c), float) else str(d.get(c)) for c in cols]) + "\n")
    print(open(out + ".txt").read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
