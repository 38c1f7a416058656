"""Reduce an ncu source-page CSV to a compact per-instruction table: address, SASS text, executed count, samples, stall columns, source line."""
import csv, gzip, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
keep = [i for i, h in enumerate(hdr) if h in ("Address", "Source", "Instructions Executed", "# Samples") or h.startswith("stall_")]
extra = [i for i, h in enumerate(hdr) if h in ("Source Location", "Location", "File", "Line")]
with gzip.open(sys.argv[2], "wt", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in keep + extra])
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[0] == "Address":
            continue
        w.writerow([r[i] for i in keep + extra])
