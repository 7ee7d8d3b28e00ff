"""Summarise an .ncu-rep per CUDA source line and per opcode (read offline with `ncu -i`):
    python tools/ncu_lines.py gpurun_out/x.ncu-rep [min_percent]"""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur, hdr, agg, ops, ops_s, stalls = None, None, {}, Counter(), Counter(), Counter()
seen = set()
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; iI = hdr.index('Instructions Executed'); iS = hdr.index('# Samples'); continue
    if hdr is None: continue
    if r[0] != '':
        try: n = int(r[iI]); s = int(r[iS])
        except ValueError: continue
        a = agg.setdefault((cur, int(r[0]), r[1][:110]), [0, 0]); a[0] += n; a[1] += s
    else:
        key = (r[2], r[3])
        if key in seen: continue
        seen.add(key)
        try: n = int(r[iI]); s = int(r[iS])
        except ValueError: continue
        op = [o for o in r[3].split() if not o.startswith('@')][0].rstrip(';')
        ops[op] += n; ops_s[op] += s
        for i, h in enumerate(hdr):
            if h.startswith('stall_') and 'Not Issued' not in h:
                try: stalls[h] += int(r[i])
                except ValueError: pass
tot = sum(v[0] for v in agg.values()); st = sum(v[1] for v in agg.values())
print("total warp instructions %d, samples %d, distinct SASS %d" % (tot, st, len(seen)))
print("-- stalls (share of samples)")
for h, v in stalls.most_common(10): print("  %-28s %5.1f%%" % (h, 100.0 * v / max(1, sum(ops_s.values()))))
print("-- opcodes")
for op, v in ops.most_common(24): print("  %-26s %6.2f%%  samples %5.2f%%" % (op, 100.0 * v / sum(ops.values()), 100.0 * ops_s[op] / max(1, sum(ops_s.values()))))
print("-- source lines >= %.1f%% of instructions or samples" % thr)
for k, v in sorted(agg.items(), key=lambda x: (x[0][0], x[0][1])):
    if v[0] >= tot * thr / 100 or v[1] >= st * thr / 100:
        print("  %-12s %4d %6.2f%% s=%5.2f%%  %s" % (k[0][:12], k[1], 100.0 * v[0] / tot, 100.0 * v[1] / st, k[2]))
