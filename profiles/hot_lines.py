"""Maps the SASS-level samples of an ncu report back to source lines: zips `ncu --page source --csv` (SASS order) with
`nvdisasm --print-line-info` of the same kernel.  usage: hot_lines.py rep.ncu-rep kernel_regex cubin mangled_substr [top]"""
import csv, io, re, subprocess, sys, collections
rep, kre, cubin, sub = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# first kernel instance only
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[start]; body = []
for r in rows[start + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"): break
    body.append(r)
ix = {h: i for i, h in enumerate(hdr)}
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
# find function
lines = []; cur = None; infn = False
for l in dis:
    m = re.match(r"\s*\.text\.(\S+):", l)
    if l.startswith("//--------------------- .text."):
        infn = sub in l; continue
    if not infn: continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
    if m: cur = (m.group(1), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l): lines.append(cur)
print(len(body), "sass rows;", len(lines), "disasm instrs")
n = min(len(body), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0])
for i in range(n):
    r = body[i]; k = lines[i]
    agg[k][0] += int(r[ix["# Samples"]] or 0); agg[k][1] += int(r[ix["Instructions Executed"]] or 0); agg[k][2] += int(r[ix["Thread Instructions Executed"]] or 0)
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
print(f"total samples {ts}, warp instr {ti}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k}: samples {100*v[0]/ts:5.1f}%  inst {100*v[1]/ti:5.1f}%  thr/inst {v[2]/max(v[1],1):5.1f}")
