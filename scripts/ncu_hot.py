"""Top SASS lines by stall samples from `ncu -i rep --page source --csv`.   python scripts/ncu_hot.py rep [N] [stall column]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 30; col = sys.argv[3] if len(sys.argv) > 3 else "# Samples"
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in body)
print("total samples", tot, "rows", len(body))
for k in ["stall_long_sb", "stall_wait", "stall_selected", "stall_membar", "stall_branch_resolving", "stall_no_inst", "stall_short_sb", "stall_not_selected", "stall_math", "stall_dispatch", "stall_sleep", "stall_lg", "stall_mio"]:
    print(f"  {k:26s} {sum(f(r,k) for r in body)/tot*100:6.2f}%")
order = sorted(range(len(body)), key=lambda i: -f(body[i], col))[:N]
for i in sorted(order):
    r = body[i]
    print(f"{i:5d} {r[ix['Address']][-5:]} {f(r,'# Samples')/tot*100:5.2f}% lsb={f(r,'stall_long_sb'):6.0f} wait={f(r,'stall_wait'):6.0f} mb={f(r,'stall_membar'):5.0f} ssb={f(r,'stall_short_sb'):5.0f} exec={f(r,'Instructions Executed'):10.0f}  {r[ix['Source']][:90]}")
# cumulative share by 256-row buckets (where in the listing the samples fall)
if len(sys.argv) > 4:
    B = int(sys.argv[4]); acc = {}
    for i, r in enumerate(body): acc[i // B] = acc.get(i // B, 0) + f(r, "# Samples")
    for k in sorted(acc): print(f"rows {k*B:5d}-{k*B+B-1:5d}: {acc[k]/tot*100:6.2f}%  first: {body[k*B][ix['Source']][:60]}")
