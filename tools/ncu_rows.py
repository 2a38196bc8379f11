"""One line per profiled launch of an .ncu-rep (headline raw metrics), optionally the top stall sites of launch K.
   python tools/ncu_rows.py rep.ncu-rep [K [topn]]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]
idx = [hdr.index(w) for w in want if w in hdr]
for n, r in enumerate(rows[2:]):
    print(n, " | ".join(f"{hdr[i].split('.')[0][-24:]}={r[i][:40]}" for i in idx))
if len(sys.argv) > 2:
    k = int(sys.argv[2]); topn = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    for i, r in enumerate(rows):
        if "# Samples" in r:
            h = r; start = i + 1; break
    ix = {c: i for i, c in enumerate(h)}
    data = [r for r in rows[start:] if len(r) > 4 and r[ix["# Samples"]].isdigit()]
    seen, uniq = set(), []
    for r in data:
        if r[0] not in seen:
            seen.add(r[0]); uniq.append(r)
    tot = sum(int(r[ix["# Samples"]]) for r in uniq)
    print("total samples", tot)
    stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    for r in sorted(uniq, key=lambda r: -int(r[ix["# Samples"]]))[:topn]:
        st = {c[6:]: r[ix[c]] for c in stall_cols if r[ix[c]] not in ("0", "")}
        print(r[0][-5:], r[1][:70].ljust(70), r[ix["# Samples"]].rjust(6), r[ix["Instructions Executed"]].rjust(9), st)
