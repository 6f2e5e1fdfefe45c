"""Top stall lines of the N-th kernel in an ncu report's source page (SASS view)."""
import csv, subprocess, sys
rep, which = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
blocks, cur = [], []
for l in out:
    if l.startswith('"Kernel Name"'):
        if cur: blocks.append(cur)
        cur = []
    else:
        cur.append(l)
if cur: blocks.append(cur)
rows = list(csv.reader(blocks[which]))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = rows[1:]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("kernel", which, "total samples", tot)
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))
for i in order[:top]:
    r = data[i]
    st = {c: int(r[ix[c]]) for c in stall_cols if int(r[ix[c]]) > 0}
    st = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("%5d %6s %5.1f%%  %-70s %s" % (i, r[ix["# Samples"]], 100.0 * int(r[ix["# Samples"]]) / max(tot, 1), r[ix["Source"]].strip()[:70], st))
