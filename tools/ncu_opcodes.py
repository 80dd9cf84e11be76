"""Executed warp instructions per SASS opcode of one kernel in an ncu report:  python tools/ncu_opcodes.py report.ncu-rep kernel_regex"""
import csv,subprocess,sys,collections
f=sys.argv[1]; kid=sys.argv[2]
out=subprocess.run(['ncu','-i',f,'--page','source','--csv','--kernel-name','regex:'+kid],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hi=[i for i,r in enumerate(rows) if '# Samples' in r][0]
h=rows[hi]; src=h.index('Source'); ie=h.index('Instructions Executed'); ns=h.index('# Samples')
agg=collections.Counter(); tot=0
for r in rows[hi+1:]:
    if len(r)<=ie: continue
    try: e=int(r[ie])
    except: continue
    toks=r[src].strip().split()
    if not toks: continue
    op=toks[1] if toks[0].startswith('@') else toks[0]
    op='.'.join(op.split('.')[:2]) if op.startswith(('LDS','STS','LDG','STG','SHFL','MUFU','RED','VOTE')) else op.split('.')[0]
    agg[op]+=e; tot+=e
print('total',tot)
for o,c in agg.most_common(40): print(o.ljust(14), round(100*c/tot,1), c//1000000)
