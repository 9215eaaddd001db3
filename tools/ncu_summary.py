#!/usr/bin/env python
"""Summarise one kernel of an `ncu --set full --import-source on` report: the headline metrics, the stall reasons, the
dynamic opcode mix and the blocks of SASS that execute together with their share of the issued warp instructions.

    python tools/ncu_summary.py <report.ncu-rep> [warps (x frames) per launch, to print instructions per warp(-frame)] [min share]

(the first kernel of the report, or the first one whose name matches the regex in NCU_KERNEL)"""
import csv, collections, os, re, sys, subprocess
rep=sys.argv[1]
if os.environ.get("NCU_KERNEL"): rep += " -k regex:%s -c 1" % os.environ["NCU_KERNEL"]
warps_per_frame=float(sys.argv[2]) if len(sys.argv)>2 else 1
subprocess.run(f"ncu -i {rep} --page source --csv > /tmp/_src.csv 2>/dev/null; ncu -i {rep} --page raw --csv > /tmp/_raw.csv 2>/dev/null",shell=True)
rows=list(csv.reader(open('/tmp/_raw.csv')))
h=rows[0]; r=rows[2]
for k in ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','smsp__warps_eligible.avg.per_cycle_active','launch__grid_size','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_active']:
    if k in h: print('   ',k,'=',r[h.index(k)], rows[1][h.index(k)])
st=[]
for i in range(len(h)):
    if 'issue_stalled' in h[i] and 'average_warp' in h[i] and 'not_issued' not in h[i]:
        try: st.append((float(r[i].replace(',','')),h[i]))
        except: pass
print('    stalls:',[(n.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),round(v,2)) for v,n in sorted(st,reverse=True)[:9]])
rows=list(csv.reader(open('/tmp/_src.csv')))
h=rows[1]; isrc=h.index('Source'); ie=h.index('Instructions Executed'); ins=h.index('# Samples')
data=[]
for rr in rows[2:]:
    if len(rr)!=len(h): continue
    try: data.append((rr[isrc], int(rr[ie]), int(rr[ins])))
    except: pass
tot=sum(d[1] for d in data); ts=max(1,sum(d[2] for d in data))
print('    total warp instr',tot,'per warp(-frame)',tot/warps_per_frame,'sass n',len(data))
mix=collections.Counter()
for s,e,n in data:
    m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)',s); mix[m.group(2) if m else s[:8]]+=e
print('    mix:',[(o,round(100*c/tot,1)) for o,c in mix.most_common(24)])
blocks=[]; s0=0
for i in range(1,len(data)+1):
    if i==len(data) or abs(data[i][1]-data[s0][1])>0.03*max(data[s0][1],1):
        blocks.append((s0,i-1,data[s0][1],sum(d[1] for d in data[s0:i]),sum(d[2] for d in data[s0:i]))); s0=i
thr=float(sys.argv[3]) if len(sys.argv)>3 else 0.008
for b in blocks:
    if b[3]>thr*tot:
        ops=collections.Counter()
        for s,e,n in data[b[0]:b[1]+1]:
            m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)',s); ops[m.group(2)]+=1
        print('    %5d-%5d n=%4d exec=%8d share=%5.2f%% stall=%5.2f%% %s'%(b[0],b[1],b[1]-b[0]+1,b[2],100*b[3]/tot,100*b[4]/ts,dict(ops.most_common(5))))
