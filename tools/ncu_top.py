"""Summarise an .ncu-rep: headline raw metrics + the most-sampled SASS instructions with their stall reasons.
   python tools/ncu_top.py gpurun_out/prof.ncu-rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second", "sm__cycles_active.avg",
        "launch__registers_per_thread", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum"]
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; idx = {k: i for i, k in enumerate(h)}
data = [r for r in rows[2:] if len(r) > 4]
tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
print("total samples", tot)
stall_cols = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]] or 0))[:topn]:
    st = {k[6:]: r[idx[k]] for k in stall_cols if r[idx[k]] not in ("0", "")}
    print(r[0][-5:], r[1][:72].ljust(72), r[idx["# Samples"]].rjust(5), r[idx["Instructions Executed"]].rjust(8), st)
