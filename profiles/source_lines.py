# profiles/source_lines.py REPORT.ncu-rep NIBBLES [TOP] -- per-source-line stall samples and warp instructions per nibble of one
# ncu --set full --import-source on capture (kernels built with -lineinfo); used for the profiles/r02_*_lines.txt summaries.
import csv, sys, subprocess
rep=sys.argv[1]; nib=float(sys.argv[2]); top=int(sys.argv[3]) if len(sys.argv)>3 else 40
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur=None; L=[]; hdr=None
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if r[0]=='Line No': hdr=r; continue
    if r[0]=='Function Name': fn=r[1]; continue
    if r[0] and r[0].isdigit():
        try: s=int(r[4]); inst=int(r[7])
        except: continue
        L.append((cur,int(r[0]),inst,s,r[1].strip()[:100],r))
print(fn[:120])
tot=sum(x[3] for x in L); ti=sum(x[2] for x in L)
print("total samples",tot,"warp-instr per nibble",round(ti/nib,1))
idx={h:i for i,h in enumerate(hdr)}
names=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg={}
for cur,l,inst,s,src,r in L:
    for n in names:
        v=r[idx[n]]
        if v not in('0','-',''): agg[n[6:]]=agg.get(n[6:],0)+int(v)
print("stall mix:", {k:round(100*v/tot,1) for k,v in sorted(agg.items(),key=lambda x:-x[1])[:9]})
for cur,l,inst,s,src,r in sorted(L,key=lambda x:-x[3])[:top]:
    st={n[6:]:int(r[idx[n]]) for n in names if r[idx[n]] not in('0','-','')}
    t3=sorted(st.items(),key=lambda x:-x[1])[:2]
    print("%5.1f%% %-13s %4d %6.1f i/nib %s | %s"%(100*s/tot,cur[:13],l,inst/nib,src[:78],t3))
