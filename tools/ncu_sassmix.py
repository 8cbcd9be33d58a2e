import csv, collections, sys, subprocess
rep, kern, div = sys.argv[1], sys.argv[2], float(sys.argv[3])
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--kernel-name','regex:'+kern],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
# may contain multiple kernels: split on 'Kernel Name' rows; take the one with most instructions
blocks=[]; cur=None
for r in rows:
    if r and r[0]=='Kernel Name': cur={'name':r[1],'rows':[]}; blocks.append(cur); continue
    if cur is not None: cur['rows'].append(r)
best=None
for b in blocks:
    hdr=b['rows'][0]; iS=hdr.index('Source'); iE=hdr.index('Instructions Executed'); iSm=hdr.index('# Samples')
    tot=0; byop=collections.Counter(); samp=collections.Counter()
    for r in b['rows'][1:]:
        if len(r)<=iE: continue
        try: n=int(r[iE]); s=int(r[iSm])
        except: continue
        t=r[iS].split()
        op=t[1] if t[0].startswith('@') else t[0]
        op=op.split('.')[0]
        byop[op]+=n; samp[op]+=s; tot+=n
    b['tot']=tot; b['byop']=byop; b['samp']=samp; b['n']=len(b['rows'])-1
    if best is None or tot>best['tot']: best=b
print(best['name'], 'SASS lines', best['n'], 'total warp-instr', best['tot'], 'per unit', best['tot']/div)
for op,n in best['byop'].most_common(22): print(f"  {op:10s} {n/div:9.1f}  samples {best['samp'][op]}")
