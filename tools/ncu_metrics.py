"""Trim an .ncu-rep to the per-launch metrics the roofline needs (run where `ncu` is on PATH; no GPU needed).

    python tools/ncu_metrics.py gpurun_out/prof.ncu-rep profiles/r01b_ncu_gemm.csv
Writes one CSV row per profiled launch: kernel, grid, duration, DRAM bytes read/written, tensor-pipe %, L2 hit %,
DRAM %, registers; prints the same as a table.
"""
import csv
import io
import re
import subprocess
import sys

COLS = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "dur_us"),
        ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct_active"),
        ("sm__cycles_active.avg", "sm_cycles_active"), ("sm__cycles_elapsed.avg", "sm_cycles_elapsed"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
        ("sm__inst_executed_pipe_tensor_op_hmma.sum", "hmma_inst")]


def to_mb(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {}
        for src, dst in COLS:
            if src not in hdr:
                continue
            i = hdr.index(src)
            v = r[i]
            if dst == "kernel":
                v = re.sub(r"\(CUtensorMap.*|\(const .*|\(float.*|\(.*", "", v.replace("void ", "").replace("(int)", "").replace("(bool)", ""))
            elif dst in ("dram_rd_MB", "dram_wr_MB"):
                v = "%.3f" % to_mb(v, units[i])
            elif dst == "dur_us":
                f = float(v.replace(",", ""))
                v = "%.2f" % (f / 1000.0 if units[i] in ("ns", "nsecond") else f)
            d[dst] = v
        res.append(d)
    keys = [dst for _, dst in COLS if any(dst in d for d in res)]
    with open(out, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        w.writerows(res)
    for d in res:
        print("  ".join("%s=%s" % (k, d.get(k, "")) for k in keys))


if __name__ == "__main__":
    main()
