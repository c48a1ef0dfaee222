"""profiles/r2_ncu_traffic.csv from the capture of r2_prof_traffic.py: one row per tagged launch with
the per-launch DRAM bytes (dram__bytes_read.sum, dram__bytes_write.sum), duration and throughput.

    python profiles/r2_make_traffic_csv.py gpurun_out/r2_traffic.ncu-rep > profiles/r2_ncu_traffic.csv
"""
import csv
import io
import subprocess
import sys

ORDER = [("colsumsq_partial", "stats_4096"), ("colsumsq_partial", "stats_11008"),
         ("fakequant_row", "fq_fp32"), ("fakequant_row", "fq128_fp32"), ("ste_bwd", "ste_fp32"),
         ("fakequant_row", "fq_bf16"), ("fakequant_row", "fq128_bf16"), ("ste_bwd", "ste_bf16"),
         ("gemv_m", "gemv_4096x4096"), ("gemm_mxq_pair", "gemm_4096x4096"),
         ("gemv_m", "gemv_11008x4096"), ("gemm_mxq_pair", "gemm_11008x4096"),
         ("gemv_m", "gemv_4096x11008"), ("gemm_mxq_pair", "gemm_4096x11008"),
         ("ptq_tile16", "ptq_4096x4096"), ("ptq_tile16", "ptq_4096x11008")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "msecond": 1e3, "ms": 1e3}


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        return float(r[ix[name]].replace(",", "")) * UNIT.get(units[ix[name]], 1.0)
    w = csv.writer(sys.stdout)
    w.writerow(["tag", "kernel", "duration_us", "dram_read_bytes", "dram_write_bytes", "dram_pct_of_peak", "sm_pct_of_peak", "registers"])
    k = 0
    for r in data:
        name = r[ix["Kernel Name"]]
        if k < len(ORDER) and ORDER[k][0] in name:
            w.writerow([ORDER[k][1], name, f"{val(r, 'gpu__time_duration.sum'):.2f}", f"{val(r, 'dram__bytes_read.sum'):.0f}",
                        f"{val(r, 'dram__bytes_write.sum'):.0f}", r[ix["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]],
                        r[ix["sm__throughput.avg.pct_of_peak_sustained_elapsed"]], r[ix["launch__registers_per_thread"]]])
            k += 1
    if k != len(ORDER):
        print(f"warning: matched {k} of {len(ORDER)} expected launches", file=sys.stderr)


if __name__ == "__main__":
    main()
