#!/usr/bin/env python
"""Attribute the executed warp instructions / stall samples of one kernel in an ncu report to
source lines and functions (ncu --import-source on; library built with -lineinfo).

usage: ncu_by_source.py <raw.csv> <source.csv> <nvdisasm -g -c output> <mangled-name regex> <src dir>"""
import collections, csv, re, sys


def main(raw, srcsv, sass, pattern, srcdir):
    rows = list(csv.reader(open(raw)))
    hdr, r = rows[0], rows[2]
    for k in ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
              'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
              'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
              'lts__t_sector_hit_rate.pct', 'smsp__warps_eligible.avg.per_cycle_active', 'launch__grid_size',
              'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers']:
        if k in hdr:
            print('   ', k, '=', r[hdr.index(k)])
    st = []
    for i in range(len(hdr)):
        if 'issue_stalled' in hdr[i] and 'average_warp' in hdr[i] and 'not_issued' not in hdr[i]:
            try:
                st.append((float(r[i].replace(',', '')), hdr[i]))
            except ValueError:
                pass
    print('    stalls:', [(n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''),
                          round(v, 2)) for v, n in sorted(st, reverse=True)[:8]])
    txt = open(sass).read()
    m = re.search(r'\.text\.' + pattern + r'[^\n]*\n(.*?)(?=\n\.section|\Z)', txt, re.S)
    cur, mapping = None, []
    for l in m.group(1).split('\n'):
        mm = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if mm:
            cur = (mm.group(1).split('/')[-1], int(mm.group(2)))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s', l):
            mapping.append(cur)
    rows = list(csv.reader(open(srcsv)))
    h = rows[1]
    ii, ni = h.index('Instructions Executed'), h.index('# Samples')
    ex, sm = [], []
    for rr in rows[2:]:
        if len(rr) != len(h):
            continue
        try:
            ex.append(int(rr[ii])); sm.append(int(rr[ni]))
        except ValueError:
            pass
    print('    sass', len(mapping), 'profiled', len(ex), 'warp-instr', sum(ex))
    n = min(len(ex), len(mapping))
    byline, bys = collections.Counter(), collections.Counter()
    for k in range(n):
        byline[mapping[k]] += ex[k]; bys[mapping[k]] += sm[k]
    tot, ts = sum(ex), max(sum(sm), 1)
    src = {}

    def text(loc):
        fn, ln = loc
        if fn not in src:
            try:
                src[fn] = open(srcdir + '/' + fn).read().split('\n')
            except OSError:
                src[fn] = []
        return src[fn][ln - 1].strip()[:84] if 0 < ln <= len(src[fn]) else ''

    def region(loc):
        if loc is None:
            return 'none'
        fn, ln = loc
        text(loc)
        t = src[fn]
        for k in range(min(ln, len(t)) - 1, 0, -1):
            if re.match(r'^(__device__|__global__|static|[a-z_0-9]+_kernel\(|int |void )', t[k]) and '(' in t[k]:
                mm = re.search(r'(\w+)\s*\(', t[k])
                return mm.group(1) if mm else '?'
        return fn
    byreg, sreg = collections.Counter(), collections.Counter()
    for loc, c in byline.items():
        byreg[region(loc)] += c; sreg[region(loc)] += bys[loc]
    print('--- by function: % executed instr, % stall samples')
    for k, c in byreg.most_common(14):
        print('%5.1f%% %5.1f%%  %s' % (100 * c / tot, 100 * sreg[k] / ts, k))
    print('--- top lines')
    for loc, c in byline.most_common(70):
        if loc:
            print('%5.1f%% %5.1f%% %s:%d  %s' % (100 * c / tot, 100 * bys[loc] / ts, loc[0], loc[1], text(loc)))


if __name__ == '__main__':
    main(*sys.argv[1:6])
