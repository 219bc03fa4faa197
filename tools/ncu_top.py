"""Top SASS instructions by stall samples from an `ncu --page source --csv` export. usage: ncu_top.py file.csv [n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = rows[2:]
def f(r, h):
    try: return float(r[ix[h]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
tot_inst = sum(f(r, "Instructions Executed") for r in data)
print(f"instructions(SASS lines)={len(data)} samples={tot:.0f} warp-instructions executed={tot_inst:.0f}")
agg = collections.Counter()
for r in data:
    for h in stall_cols: agg[h] += f(r, h)
print("stall totals:", ", ".join(f"{h[6:]}={v:.0f}" for h, v in agg.most_common(8)))
op = collections.Counter(); opi = collections.Counter()
for r in data:
    src = r[ix["Source"]].strip()
    toks = src.split()
    m = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    m = m.split(".")[0]
    op[m] += f(r, "# Samples"); opi[m] += f(r, "Instructions Executed")
print("by opcode (samples | warp-instr):", ", ".join(f"{k}={v:.0f}|{opi[k]:.0f}" for k, v in op.most_common(18)))
order = sorted(range(len(data)), key=lambda i: -f(data[i], "# Samples"))[:n]
for i in sorted(order):
    r = data[i]
    top = sorted(((f(r, h), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {f(r,'# Samples'):7.0f} {f(r,'Instructions Executed'):9.0f}  {r[ix['Source']][:70]:70s} {top[0][1]}={top[0][0]:.0f} {top[1][1]}={top[1][0]:.0f}")
