"""Summarise an `ncu --page source --csv` export: share of the warp instructions by number of active lanes, and the
largest contiguous SASS regions of one lane-count class."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
print(rows[0][1][:110])
h = rows[1]; data = rows[2:]
iS, iI, iT, iAv, iSm = (h.index(k) for k in ('Source', 'Instructions Executed', 'Thread Instructions Executed', 'Avg. Threads Executed', '# Samples'))
def f(x):
    try: return float(x)
    except ValueError: return 0.0
tot = sum(f(r[iI]) for r in data); tt = sum(f(r[iT]) for r in data); ts = sum(f(r[iSm]) for r in data)
print('warp instructions %.4g  thread instructions %.4g  average active lanes %.2f' % (tot, tt, tt / tot))
def cls(a): return 1 if a < 1.5 else 4 if a < 6 else 12 if a < 14 else 20 if a < 22.5 else 25 if a < 26 else 32
by = {}
for r in data:
    k = cls(f(r[iAv])); by[k] = by.get(k, 0) + f(r[iI])
print('share of warp instructions by active lanes:', {k: round(v / tot, 3) for k, v in sorted(by.items())})
seg = []; cur = None
for i, r in enumerate(data):
    n = f(r[iI])
    if n == 0: continue
    k = cls(f(r[iAv]))
    if cur and cur[0] == k and i - cur[2] < 4: cur[2] = i; cur[3] += n; cur[4] += f(r[iSm])
    else:
        if cur: seg.append(cur)
        cur = [k, i, i, n, f(r[iSm])]
seg.append(cur); seg.sort(key=lambda s: -s[3])
for s in seg[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print('lanes~%-2d rows %5d-%-5d instructions %.3f  samples %.3f' % (s[0], s[1], s[2], s[3] / tot, s[4] / ts))
