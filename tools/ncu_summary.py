"""Summaries of ncu output for profiles/ (run HERE on files brought back in gpurun_out/).
  launch list:  python tools/ncu_summary.py launches gpurun_out/launches.csv          (ncu --metrics gpu__time_duration.sum --csv --log-file ...)
  full capture: ncu -i x.ncu-rep --page raw --csv > x.csv ; python tools/ncu_summary.py raw x.csv
"""
import collections, csv, re, sys

KEEP = ("gpu__time_duration.sum", "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__registers_per_thread",
        "sm__cycles_elapsed.max", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")

def rows(path):
    lines = [l for l in open(path, errors="replace") if not l.startswith("==")]
    return list(csv.reader(lines))

def launches(path):
    r = rows(path)
    hdr = r[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
    per = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
    for x in r[1:]:
        if len(x) <= vi: continue
        name = re.sub(r"\(.*", "", x[ki]).replace("cg::", "").strip()
        v = float(x[vi].replace(",", "")); v = v / 1e3 if x[ui] in ("ns", "nsecond") else v   # -> us
        per[name][0] += 1; per[name][1] += v; tot += v
    print("%d launches, %.2f ms in total (cold-cache, serialised under ncu: compare SHARES, not absolutes)" % (sum(v[0] for v in per.values()), tot / 1e3))
    for k, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1])[:30]:
        print("   %-34s %5d launches %10.1f us %5.1f%%   avg %7.1f us" % (k[:34], n, t, 100 * t / tot, t / n))

def raw(path):
    r = rows(path)
    hdr, units = r[0], r[1]
    ki = hdr.index("Kernel Name")
    for x in r[2:]:
        print(re.sub(r"\(.*", "", x[ki]).replace("cg::", ""), " grid", x[hdr.index("Grid Size")] if "Grid Size" in hdr else "")
        for m in KEEP:
            if m in hdr: print("   %-100s %s %s" % (m, x[hdr.index(m)], units[hdr.index(m)]))

if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
