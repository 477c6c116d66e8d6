"""Summarise an `ncu --page source --csv` dump (SASS view): lane utilisation histogram, stall mix,
and the hottest SASS regions.  python tools/ncu_src_summary.py gpurun_out/prof_X_source.csv [raw.csv]"""
import csv, collections, sys

def main(path, raw=None):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    I = lambda r, k: int(float(r[ci[k]] or 0))
    tot = sum(I(r, 'Instructions Executed') for r in data)
    thr = sum(I(r, 'Thread Instructions Executed') for r in data)
    print("# %s" % rows[0][1])
    print("sass_lines %d warp_instr %.4g thread_instr %.4g avg_active_lanes %.2f" % (len(data), tot, thr, thr / max(tot, 1)))
    b = collections.Counter()
    for r in data:
        n = I(r, 'Instructions Executed')
        if n:
            b[min(31, int(I(r, 'Thread Instructions Executed') / n)) // 4 * 4] += n
    print("warp instructions by active lanes: " + "  ".join("%d-%d:%.1f%%" % (k, k + 3, 100 * b[k] / tot) for k in sorted(b)))
    stalls = [h for h in hdr if h.startswith('stall_')]
    st = {h: sum(I(r, h) for r in data) for h in stalls}
    ts = sum(st.values()) or 1
    print("stall samples: " + "  ".join("%s:%.1f%%" % (k[6:], 100 * v / ts) for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v > ts * 0.01))
    # opcode mix
    ops = collections.Counter()
    for r in data:
        t = r[ci['Source']].split()
        if not t: continue
        op = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
        ops[op.split('.')[0]] += I(r, 'Instructions Executed')
    print("opcode mix: " + "  ".join("%s:%.1f%%" % (k, 100 * v / tot) for k, v in ops.most_common(14)))
    # global memory instructions
    print("memory instructions (executed M, avg lanes, samples, sass):")
    for r in data:
        s = r[ci['Source']]
        if any(x in s for x in ('LDG', 'LDGSTS', 'ATOM', 'RED', 'STG', 'LD.E', 'UBLKCP')) and I(r, 'Instructions Executed') > tot * 0.0005:
            n = I(r, 'Instructions Executed')
            print("  %8.1fM %5.1f %7s  %s" % (n / 1e6, I(r, 'Thread Instructions Executed') / n, r[ci['# Samples']], s[:90]))
    if raw:
        rr = list(csv.reader(open(raw)))
        d = dict(zip(rr[0], zip(rr[1], rr[2])))
        for k in ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
                  'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
                  'lts__t_sector_hit_rate.pct', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
                  'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
                  'smsp__warps_eligible.avg.per_cycle_active', 'launch__registers_per_thread', 'launch__grid_size'):
            if k in d: print("%-70s %-10s %s" % (k, d[k][0], d[k][1]))

if __name__ == "__main__":
    main(*sys.argv[1:3])
