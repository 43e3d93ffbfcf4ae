"""profiles/*traffic*.csv from an `ncu --set full` report: per kernel (first launch of each name) the DRAM bytes read and
written, duration, and the batch the capture ran at.  bench.py's roofline.traffic reads the newest such file by kernel
name (so the number is never a constant in bench.py).   usage: ncu_traffic.py report.ncu-rep out.csv batch"""
import csv, io, subprocess, sys
rep, out, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def val(r, name):
    v = float(r[ix[name]].replace(",", ""))
    u = units[ix[name]].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1)
seen, lines = {}, []
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    name = r[ix["Kernel Name"]]
    key = name.split("(")[0]
    seen.setdefault(key, []).append((val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum"), val(r, "gpu__time_duration.sum"),
                                     r[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]] if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in ix else "",
                                     r[ix["sm__warps_active.avg.pct_of_peak_sustained_active"]] if "sm__warps_active.avg.pct_of_peak_sustained_active" in ix else ""))
with open(out, "w") as f:
    f.write("kernel,launch,dram_read_bytes,dram_write_bytes,duration_us,tensor_pipe_pct,warps_active_pct,batch\n")
    for key, ls in seen.items():
        for i, (rd, wr, du, tp, wa) in enumerate(ls):
            f.write(f"\"{key}\",{i},{rd:.0f},{wr:.0f},{du:.2f},{tp},{wa},{batch}\n")
print("wrote", out, {k: len(v) for k, v in seen.items()})
