"""Helper (not a test): aggregate ncu warp-stall samples of a report by source file:line (CUDA view)."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None
agg = collections.Counter(); inst = collections.Counter(); src = {}
hdr = None
for row in csv.reader(txt.splitlines()):
    if not row:
        continue
    if row[0] == "File Path":
        cur = row[1].split("/")[-1]; continue
    if row[0] == "Function Name":
        continue
    if row[0] == "Line No":
        hdr = row; continue
    if hdr is None or cur is None:
        continue
    try:
        ln = int(row[0])
    except ValueError:
        continue
    d = dict(zip(hdr, row))
    def num(x):
        try:
            return int(float(x))
        except (TypeError, ValueError):
            return 0
    s = num(d.get("Warp Stall Sampling (All Samples)"))
    e = num(d.get("Instructions Executed"))
    agg[(cur, ln)] += s; inst[(cur, ln)] += e; src[(cur, ln)] = row[1].strip()[:90]
tot = sum(agg.values()); ti = sum(inst.values())
print(f"total samples {tot}, warp instructions {ti}")
byfile = collections.Counter()
for (f, l), v in agg.items():
    byfile[f] += v
print({k: f"{100*v/tot:.1f}%" for k, v in byfile.most_common()})
for (f, l), v in agg.most_common(top):
    print(f"{100*v/tot:5.1f}% samples {100*inst[(f,l)]/ti:5.1f}% inst  {f}:{l:<5d} {src[(f,l)]}")
