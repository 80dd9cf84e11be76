"""Per CUDA source line: share of stall samples and of executed warp instructions of one kernel in an ncu report
captured with --import-source on (build with -lineinfo):  python tools/ncu_source_lines.py report.ncu-rep kernel_regex [min_percent]"""
import csv,subprocess,sys,collections
f=sys.argv[1]; kid=sys.argv[2]; thr=float(sys.argv[3]) if len(sys.argv)>3 else 0.8
out=subprocess.run(['ncu','-i',f,'--page','source','--csv','--print-source','cuda,sass','--kernel-name','regex:'+kid],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
his=[i for i,r in enumerate(rows) if '# Samples' in r]+[len(rows)]
files=[]
for a,b in zip(his[:-1],his[1:]):
    h=rows[a]; ln=h.index('Line No'); ns=h.index('# Samples'); ie=h.index('Instructions Executed'); sc=[i for i,x in enumerate(h) if x=='Source'][0]
    fname=''
    for r in rows[max(0,a-3):a]:
        if r and r[0]=='File Name': fname=r[1]
    agg=collections.OrderedDict(); cur=None
    for r in rows[a+1:b]:
        if len(r)<=ie: continue
        if r[ln].strip():
            cur=(int(r[ln]),r[sc].strip()[:95]); agg.setdefault(cur,[0,0]); continue
        try: s=int(r[ns]); e=int(r[ie])
        except: continue
        if cur: agg[cur][0]+=s; agg[cur][1]+=e
    files.append((fname,agg))
ts=sum(v[0] for _,agg in files for v in agg.values()); te=sum(v[1] for _,agg in files for v in agg.values())
print('samples',ts,'inst',te)
for fname,agg in files:
    fs=sum(v[1] for v in agg.values())
    print('==',fname, round(100*fs/max(te,1),1),'% inst')
    for (l,t),(s,e) in agg.items():
        if 100*s/max(ts,1)>=thr or 100*e/max(te,1)>=thr:
            print(str(l).rjust(4), str(round(100*s/ts,1)).rjust(5), str(round(100*e/te,1)).rjust(5), t)
