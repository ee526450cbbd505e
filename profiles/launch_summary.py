import csv,sys
rows=[r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
seq=[(r[ki].split('(')[0].replace('void rt::','').replace('rt::',''), float(r[vi].replace(',',''))/1e3) for r in rows[1:]]
gens=[i for i,(k,_) in enumerate(seq) if k.startswith('wf_generate')]
g=gens[3]; end=gens[4] if len(gens)>4 else len(seq)
it=[]; cur={}; tot={}
for k,us in seq[g:end]:
    name=k.split('<')[0]; tot[name]=tot.get(name,0)+us
    if name=='wf_cull':
        if cur: it.append(cur)
        cur={}
    cur[name]=cur.get(name,0)+us
it.append(cur)
print('totals (us):',{k:round(v) for k,v in tot.items()}, 'sum', round(sum(tot.values())))
n=int(sys.argv[2]) if len(sys.argv)>2 else 24
for i,c in enumerate(it[:n]): print(i, ' '.join(f'{k[3:]}={v:.0f}' for k,v in c.items() if k.startswith('wf_')))
print('iterations', len(it), 'tail(>=20) sum', round(sum(sum(c.values()) for c in it[20:])))
