import csv, sys, subprocess
rep=sys.argv[1]; B=int(sys.argv[2]) if len(sys.argv)>2 else 200
out=subprocess.run(['ncu','-i',rep,'--page','details'],capture_output=True,text=True).stdout
for l in out.splitlines():
    if any(w in l for w in ['Duration','Executed Ipc Active','Issue Slots Busy','Registers Per','Achieved Occupancy','Avg. Active Threads','Executed Instructions  ','No Eligible','Warp Cycles Per Issued','Grid Size']): print(l[:110])
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); hdr,units,vals=rows[0],rows[1],rows[2]
for h,u,v in zip(hdr,units,vals):
    if 'smsp__average_warp' in h and 'issue_stalled' in h and 'ratio' in h and 'not_issued' not in h:
        try:
            if float(v)>0.1: print(f'  {h[35:-24]:40s} {float(v):8.3f}')
        except: pass
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines())); hdr=rows[1]; data=rows[2:]
ix={h:i for i,h in enumerate(hdr)}
def f(r,k):
    try: return float(r[ix[k]])
    except: return 0.0
ts=sum(f(r,'# Samples') for r in data); te=sum(f(r,'Instructions Executed') for r in data)
print('instrs in kernel', len(data), 'samples',ts,'executed',te)
keys=[k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
for b in range(0,len(data),B):
    seg=data[b:b+B]
    s=sum(f(r,'# Samples') for r in seg); ex=sum(f(r,'Instructions Executed') for r in seg)
    st={k:sum(f(r,k) for r in seg) for k in keys}
    top=sorted(st.items(), key=lambda x:-x[1])[:4]
    ops={}
    for r in seg:
        t=r[ix['Source']].split()
        if not t: continue
        op=t[1] if t[0].startswith('@') else t[0]
        ops[op.split('.')[0]]=ops.get(op.split('.')[0],0)+1
    topo=','.join(k for k,_ in sorted(ops.items(), key=lambda x:-x[1])[:3])
    if s>0.008*ts: print(f'{b:6d} samples {100*s/ts:5.1f}%  exec {100*ex/te:5.1f}%  ' + ' '.join(f'{k[6:]}={100*v/max(s,1):.0f}%' for k,v in top) + '  ['+topo+']')
