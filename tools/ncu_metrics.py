"""tools/ncu_metrics.py X.ncu-rep [template.csv] > out.csv
Extracts the metrics named in the first column of `template.csv` (default: profiles/r01_sweep_kernel_ncu.csv) from the
raw page of an ncu report (last profiled launch) as `metric,unit,value` lines — the summary that lives under profiles/."""
import csv, os, subprocess, sys
rep = sys.argv[1]
tpl = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r01_sweep_kernel_ncu.csv")
want = [l.split(",")[0] for l in open(tpl).read().splitlines()[1:] if l.strip()]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[-1]
ix = {n: i for i, n in enumerate(hdr)}
print("metric,unit,value")
for m in want:
    if m in ix:
        print("%s,%s,%s" % (m, units[ix[m]], vals[ix[m]].replace(",", "")))
