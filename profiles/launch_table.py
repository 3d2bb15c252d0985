"""Prints the per-kernel table of the LAST step in an ncu launch-list CSV (gpu__time_duration.sum)."""
import csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
names = [(r[4][:72], float(r[-1]), r[8]) for r in rows[1:]]
idx = [i for i, (k, v, g) in enumerate(names) if 'cigar_scan' in k or 'cigar_stream' in k]
s = idx[-1]; tot = sum(v for k, v, g in names[s:])
for k, v, g in names[s:]:
    print(f"{v/1000:9.1f} us {100*v/tot:5.1f}% {g:>14} {k}")
print(f"{tot/1e6:.3f} ms in {len(names)-s} launches")
