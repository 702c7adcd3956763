"""Condense an .ncu-rep (ncu --set full) into the text summary kept under profiles/, and update
profiles/traffic.json (dram bytes per launch of the scan kernel, read by bench.py for roofline.traffic).
usage: python tools/ncu_summary.py <report.ncu-rep> <mode> <out.txt> "<note>" """
import csv, json, os, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_requests_srcunit_tex_op_red.sum',
        'lts__t_requests_srcunit_tex_op_atom.sum']


def main():
    rep, mode, out, note = sys.argv[1:5]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    get = lambda k: r[hdr.index(k)] if k in hdr else None
    lines = [note, "kernel: " + r[hdr.index('Kernel Name')], ""]
    for k in KEYS:
        if k in hdr:
            lines.append("%-72s %s %s" % (k, get(k), units[hdr.index(k)]))
    stalls = []
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and 'ratio' in h and 'not_issued' not in h:
            try:
                stalls.append((float(r[i]), h))
            except ValueError:
                pass
    lines.append("")
    lines.append("top warp stall reasons (warps per issue-active cycle):")
    for v, h in sorted(stalls, reverse=True)[:8]:
        lines.append("  %-88s %.3f" % (h, v))
    # hottest source lines
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    if len(srows) > 3:
        sh = srows[2]
        iE, iS = sh.index('Instructions Executed'), sh.index('Warp Stall Sampling (All Samples)')
        d = {}
        for x in srows[3:]:
            if len(x) > iE and x[2] == '-' and x[0].isdigit():
                d.setdefault((int(x[0]), x[1].strip()[:100]), [int(x[iE]), int(x[iS])])
        te, ts = sum(v[0] for v in d.values()) or 1, sum(v[1] for v in d.values()) or 1
        lines.append("")
        lines.append("hottest source lines (kta_kernels.cuh: line, % instructions executed, % stall samples):")
        for (l, s), (e, st) in sorted(d.items(), key=lambda kv: -kv[1][0])[:14]:
            lines.append("  %5d %6.2f%% %6.2f%%  %s" % (l, 100 * e / te, 100 * st / ts, s))
    open(out, "w").write("\n".join(lines) + "\n")
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    rd = float(get('dram__bytes_read.sum')) * scale[units[hdr.index('dram__bytes_read.sum')]]
    wr = float(get('dram__bytes_write.sum')) * scale[units[hdr.index('dram__bytes_write.sum')]]
    tj = os.path.join(os.path.dirname(out), "traffic.json")
    t = json.load(open(tj)) if os.path.exists(tj) else {}
    t[mode] = rd + wr
    t.setdefault("_source", {})[mode] = os.path.basename(out)
    json.dump(t, open(tj, "w"), indent=1)
    print("\n".join(lines[:30]))


if __name__ == "__main__":
    main()
