"""Where a kernel spends its warp-time, from an `ncu --set full --import-source on` report: headline metrics, then the
stall-sample mix of (a) everything before the first tensor instruction (staging) and (b) the tile body, and a
250-instruction-wide scan of the body (opcode mix, executed count, top stall reasons per window).

    python tools/ncu_stall_segments.py gpurun_out/attn.ncu-rep

Used for the attention kernel analysis in DESIGN.md section 4 (staging 28 % of warp-time, spill reloads ~10 %)."""
import csv, collections, subprocess, sys
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; vals=rows[2]
want=['gpu__time_duration.sum','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed']
for i,h in enumerate(hdr):
    if h in want: print(h, rows[1][i], vals[i])
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; data=rows[2:]
ia=hdr.index('Instructions Executed'); isrc=hdr.index('Source'); isamp=hdr.index('# Samples')
stall_cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
ops=[r[isrc].split()[0] if not r[isrc].startswith('@') else r[isrc].split()[1] for r in data]
first=next(i for i,o in enumerate(ops) if o.startswith('HMMA'))
last=max(i for i,o in enumerate(ops) if o.startswith('HMMA'))
nw=max(int(r[ia]) for r in data[first:last])
def seg(a,b,name):
    ex=sum(int(r[ia]) for r in data[a:b]); sm=sum(int(r[isamp]) for r in data[a:b])
    st=collections.Counter()
    for r in data[a:b]:
        for c in stall_cols: st[hdr[c]]+=int(r[c] or 0)
    print(name,'static',b-a,'exec/warp',round(ex/nw,1),'samples',sm,dict(st.most_common(7)))
seg(0,first-40,'staging+Q')
# split tile body in phases using HMMA clusters
seg(first-40,last+100,'tile body')
tot=sum(int(r[isamp]) for r in data)
print('total samples',tot)
# finer: chunks of 250 instructions in body
for a in range(first-40,last+100,250):
    b=min(a+250,last+100)
    sm=sum(int(r[isamp]) for r in data[a:b]); ex=sum(int(r[ia]) for r in data[a:b])
    c=collections.Counter(o.split('.')[0] for o in ops[a:b])
    st=collections.Counter()
    for r in data[a:b]:
        for cc in stall_cols: st[hdr[cc][6:]]+=int(r[cc] or 0)
    print(a,'samples',sm,'exec/warp',round(ex/nw), dict(c.most_common(5)), dict(st.most_common(4)))
