"""List the hottest SASS instructions (by warp-stall samples) of one kernel from `ncu --page source --csv` output."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if r and r[0] == "Address")
idx = {h: i for i, h in enumerate(hdr)}
S = idx["# Samples"]
data = [r for r in rows if len(r) > S and r[S].isdigit() and r[0].startswith("0x")]
tot = sum(int(r[S]) for r in data)
print("total samples", tot, "instructions", len(data))
top = sorted(range(len(data)), key=lambda i: -int(data[i][S]))[:ntop]
for i in sorted(top):
    r = data[i]
    print(i, r[S], r[idx["Source"]].strip()[:110], "| smem wf", r[idx["L1 Wavefronts Shared"]], "ideal",
          r[idx["L1 Wavefronts Shared Ideal"]])
h = collections.Counter()
for i, r in enumerate(data):
    h[i // 250] += int(r[S])
print("samples per 250-instruction bucket:", sorted(h.items()))
