import csv, io, subprocess, sys, collections
path, kernel = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu","-i",path,"--page","source","--print-source","cuda,sass","--csv","--kernel-name",kernel],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
hdr=None; fname=None; func=None
agg=collections.Counter(); inst=collections.Counter(); thr=collections.Counter()
R=[("device.h",130,345,"surfaces/cell_contains/cell_distance"),("device.h",345,430,"lattice"),("device.h",430,612,"find_cell"),("device.h",612,675,"distance_to_boundary"),("device.h",840,1085,"ce xs"),("device.h",1085,1475,"collision physics"),("device.h",1475,2000,"tally"),("engine.cu",100,260,"init/xs/record_flight"),("engine.cu",260,336,"ev_advance"),("engine.cu",336,432,"ev_surface"),("engine.cu",432,480,"ev_collide"),("engine.cu",480,600,"bodies/append")]
for r in rows:
    if not r: continue
    if r[0]=="File Path": fname=r[1].split("/")[-1]
    elif r[0]=="Function Name": func=r[1]
    elif r[0]=="Line No": hdr=r
    elif hdr and r[0].isdigit() and kernel in (func or ""):
        d=dict(zip(hdr,r)); n=int(r[0])
        key=fname
        for f,a,b,name in R:
            if f==fname and a<=n<b: key=name
        try:
            agg[key]+=int(d["# Samples"]); inst[key]+=int(d["Instructions Executed"]); thr[key]+=int(d["Thread Instructions Executed"])
        except: pass
ts=sum(agg.values()); ti=sum(inst.values())
for k,v in agg.most_common():
    print(f"{v/ts:6.3f} smp {inst[k]/ti:6.3f} inst lanes={thr[k]/max(inst[k],1):5.1f} {k}")
