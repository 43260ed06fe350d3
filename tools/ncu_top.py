"""Top stall sites of one kernel from an ncu report's SASS source page.
usage: python tools/ncu_top.py report.ncu-rep kernel_regex [n]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr) and r[ix["# Samples"]].isdigit()]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in body) for s in stalls}
print("total samples", tot)
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
body.sort(key=lambda r: -int(r[ix["# Samples"]] or 0))
for r in body[:n]:
    st = {s[6:]: int(r[ix[s]] or 0) for s in stalls if int(r[ix[s]] or 0)}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f'{int(r[ix["# Samples"]]):7d} {100 * int(r[ix["# Samples"]]) / tot:5.1f}%  {r[ix["Address"]][-5:]}  {r[ix["Source"]][:70]:70s} {top}')
