"""ncu raw CSV (ncu -i X.ncu-rep --page raw --csv) -> {kernel: {dram_read_bytes, dram_write_bytes, duration_us}} (first launch
of each kernel name).  usage: python tools/ncu_traffic.py raw.csv out.json"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def val(r, name):
    v = r[col[name]].replace(",", "")
    u = units[col[name]]
    x = float(v)
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ns": 1e-3, "ms": 1e3}.get(u, 1)
    return x * mult
out = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    short = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    if short in out:
        continue
    out[short] = {"dram_read_bytes": val(r, "dram__bytes_read.sum"), "dram_write_bytes": val(r, "dram__bytes_write.sum"),
                  "duration_us": val(r, "gpu__time_duration.sum")}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
