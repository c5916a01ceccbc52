"""profiles/dgemm_traffic.json from an `ncu --set full` capture of dgemm_tma_dmma_kernel (raw CSV page), stamped with the
sha256 of nalgebra_b200/csrc/dgemm.cu so that bench.py drops it when the kernel source changes.
Usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; python tools/traffic_json.py raw.csv"""
import csv, hashlib, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[-1]
d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
def bytes_of(k):
    v = float(d[k].replace(",", "")); unit = u[k].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12, "b": 1, "kb": 1e3, "mb": 1e6, "gb": 1e9, "tb": 1e12}[unit]
rd, wr = bytes_of("dram__bytes_read.sum"), bytes_of("dram__bytes_write.sum")
out = {"kernel": d.get("Kernel Name", "dgemm_tma_dmma_kernel"), "workload": "16384x16384x16384 f64, 1 GPU",
       "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes_per_launch": rd + wr,
       "algorithmic_bytes_per_launch": 3 * 16384 * 16384 * 8,
       "duration_ms": float(d["gpu__time_duration.sum"].replace(",", "")) * {"nsecond": 1e-6, "ns": 1e-6, "usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0, "second": 1e3, "s": 1e3}[u["gpu__time_duration.sum"].lower()],
       "kernel_source_sha256": hashlib.sha256(open("nalgebra_b200/csrc/dgemm.cu", "rb").read()).hexdigest(),
       "source": "ncu --set full --clock-control none -k regex:dgemm_tma -c 1 python tools/gemm_once.py (dram__bytes_read.sum + dram__bytes_write.sum)"}
json.dump(out, open("profiles/dgemm_traffic.json", "w"), indent=1)
print(out)
