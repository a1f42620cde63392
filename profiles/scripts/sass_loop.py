"""Counted SASS listing of a kernel's hot loop from an `ncu --page source --csv` export.

    sass_loop.py <source.csv> <steps> [min_share]

Prints every SASS instruction whose executed warp-instruction count is at least `min_share` (default 0.2) of the Metropolis
steps' warp count (steps / 32), with warp executions per warp-step, thread executions per step and the average active lanes,
followed by the totals: issue slots per step (warp instructions x 32 / steps) and thread instructions per step, for the
listed instructions and for the whole kernel."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
steps = float(sys.argv[2])
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.2
print(rows[0][1])
h, data = rows[1], rows[2:]
iS, iI, iT, iP = (h.index(k) for k in ("Source", "Instructions Executed", "Thread Instructions Executed", "Predicated-On Thread Instructions Executed"))


def f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


wsteps = steps / 32.0
tot_w = sum(f(r[iI]) for r in data)
tot_t = sum(f(r[iT]) for r in data)
tot_p = sum(f(r[iP]) for r in data)
sel_w = sel_t = sel_p = 0.0
n = 0
print("%-5s %-72s %9s %9s %6s" % ("row", "SASS", "warp/step", "thr/step", "lanes"))
for i, r in enumerate(data):
    w, t, pr = f(r[iI]), f(r[iT]), f(r[iP])
    if w < min_share * wsteps:
        continue
    n += 1
    sel_w += w
    sel_t += t
    sel_p += pr
    print("%-5d %-72s %9.3f %9.3f %6.1f" % (i, r[iS].strip()[:72], w / wsteps, pr / steps, t / max(w, 1.0)))
print()
print("listed instructions: %d; issue slots per step %.2f; thread instructions per step %.2f (predicated-on %.2f)" %
      (n, sel_w * 32 / steps, sel_t / steps, sel_p / steps))
print("whole kernel:        issue slots per step %.2f; thread instructions per step %.2f (predicated-on %.2f); average active lanes %.2f" %
      (tot_w * 32 / steps, tot_t / steps, tot_p / steps, tot_t / tot_w))
