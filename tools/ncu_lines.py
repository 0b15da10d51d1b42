"""Summarise an ncu report per CUDA source line (stall samples), for one kernel.
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [top]"""
import csv, subprocess, sys, collections, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur, hdr = None, None
agg = collections.OrderedDict()
stall_names = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ix = {h: i for i, h in enumerate(hdr)}; stall_names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[2] != "-": continue            # SASS rows have an address; source rows carry the per-line totals
    try: s = int(r[ix["# Samples"]])
    except ValueError: continue
    key = (cur, int(r[0]))
    st = {n: int(r[ix[n]] or 0) for n in stall_names}
    inst = int(r[ix["Instructions Executed"]] or 0)
    if key in agg:
        agg[key][0] += s; agg[key][2] += inst
        for n in stall_names: agg[key][3][n] += st[n]
    else:
        agg[key] = [s, r[1].strip()[:90], inst, st]
tot = sum(v[0] for v in agg.values())
print("total samples", tot)
byfile = collections.Counter()
for (f, l), v in agg.items(): byfile[f] += v[0]
print({f: round(100 * c / tot, 1) for f, c in byfile.items()})
allst = collections.Counter()
for v in agg.values():
    for n, c in v[3].items(): allst[n] += c
print("stalls:", {n: round(100 * c / max(1, sum(allst.values())), 1) for n, c in allst.most_common(8)})
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    ts = sorted(v[3].items(), key=lambda kv: -kv[1])[:2]
    print(f"{100*v[0]/tot:5.1f}%  {f}:{l:<4d} inst={v[2]:>9d} {ts[0][0][6:]}={ts[0][1]} {ts[1][0][6:]}={ts[1][1]} | {v[1]}")
