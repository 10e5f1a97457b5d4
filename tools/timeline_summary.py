"""Text summary of a step timeline written by `CATGEN_BENCH_TIMELINE=<file> python bench.py ...` (cg_profile_timeline: one eager step with the
concurrent lanes on and an event pair around every launch).  Usage: python tools/timeline_summary.py timeline.json [--rows]"""
import collections, json, sys

def main():
    rows = json.load(open(sys.argv[1]))
    ev = sorted(rows, key=lambda r: r[2])
    busy, cs, ce = 0.0, ev[0][2], ev[0][3]
    for r in ev[1:]:
        if r[2] > ce: busy += ce - cs; cs, ce = r[2], r[3]
        else: ce = max(ce, r[3])
    busy += ce - cs
    span = max(r[3] for r in rows)
    print("%d launches, span %.2f ms, some kernel running %.2f ms (%.1f %%), sum of kernel durations %.2f ms" % (len(rows), span / 1e3, busy / 1e3, 100 * busy / span, sum(r[3] - r[2] for r in rows) / 1e3))
    per = collections.defaultdict(lambda: [0, 0.0])
    for r in rows: per[r[0]][0] += 1; per[r[0]][1] += r[3] - r[2]
    print("%-28s %5s %9s" % ("kernel", "n", "sum us"))
    for k, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1])[:25]: print("%-28s %5d %9.0f" % (k, n, t))
    lanes = collections.defaultdict(float)
    for r in rows: lanes[r[1]] += r[3] - r[2]
    print("kernel time per lane (us; -1 = main stream):", {k: round(v) for k, v in sorted(lanes.items())})
    if "--rows" in sys.argv:
        for r in ev: print("%8.0f %8.0f %6.0f lane%2d %s" % (r[2], r[3], r[3] - r[2], r[1], r[0]))

if __name__ == "__main__":
    main()
