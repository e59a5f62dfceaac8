"""profiles/sweep_traffic.json <- DRAM bytes per launch of an ncu --set full capture of the config-2 sweep kernel taken inside
bench.py (the figure bench.py reports as roofline.traffic).  usage: update_sweep_traffic.py report.ncu-rep "<source note>" """
import csv, json, os, subprocess, sys
rep, note = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr, units = rows[0], rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot, n, name = 0.0, 0, ""
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[hdr.index(m)]) * scale[units[hdr.index(m)]]
    n += 1
d = {"kernel": name[:120], "source": note, "dram_bytes_per_launch": tot / n, "launches": n}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(root, "profiles", "sweep_traffic.json"), "w") as f:
    json.dump(d, f, indent=1)
print(d)
