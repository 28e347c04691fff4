import random, subprocess, sys, hashlib
import os; R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
def nals(b):
    idx=[]; i=0
    while True:
        j=b.find(b'\x00\x00\x01',i)
        if j<0: break
        idx.append(j); i=j+3
    return [b[idx[k]:(idx[k+1] if k+1<len(idx) else len(b))] for k in range(len(idx))]
def run(cmd):
    try:
        r=subprocess.run(cmd,capture_output=True,timeout=60); return r.stdout.decode(errors='replace'), r.returncode
    except subprocess.TimeoutExpired: return "timeout", -1
n=int(sys.argv[1]); seed0=int(sys.argv[2]); bad=0; crash=0; idmis=0
for it in range(n):
    r=random.Random(seed0+it)
    gop=r.choice(["IP","IPB"])
    a=["-W",str(r.choice([2,3,5])),"-H",str(r.choice([2,3])),"-n",str(r.choice([12,20,30])),"-s",str(seed0+it),"--gop",gop,"--refs",str(r.randint(1,5)),"--idr",str(r.choice([9,17,40])),"--deblock","0","--wp",str(r.randint(0,2))]
    if r.random()<0.3: a.append("--cavlc")
    if gop=="IP" and r.random()<0.3: a += ["--poc-type",str(r.randint(1,2))]
    subprocess.run([R+"/tools/gen264","-o","/tmp/e264_campaign_dc.264"]+a,capture_output=True)
    ns=nals(open("/tmp/e264_campaign_dc.264","rb").read())
    cand=[k for k,x in enumerate(ns) if (x[3]&31)==1 and ((x[3]>>5)&3)>0]
    if not cand: continue
    drop=set(r.sample(cand, min(len(cand), r.choice([1,1,2,3]))))
    open("/tmp/e264_campaign_dc_drop.264","wb").write(b''.join(x for k,x in enumerate(ns) if k not in drop))
    x,rx=run([R+"/oracle/_ref/ref_decode","/tmp/e264_campaign_dc_drop.264","-c"]); y,ry=run([R+"/oracle/oracle_decode","/tmp/e264_campaign_dc_drop.264","-c"])
    if rx!=0: crash+=1; continue
    if x!=y:
        bad+=1
        ix=[l.split()[3] for l in x.splitlines() if l.startswith("frame ")]; iy=[l.split()[3] for l in y.splitlines() if l.startswith("frame ")]
        if ix!=iy: idmis+=1
        print("DIFF", " ".join(a), "dropped", sorted(drop), "ids equal" if ix==iy else "IDS DIFFER")
print("drop campaign:", n, "streams,", bad, "mismatching (", idmis, "with different FrameIds ),", crash, "reference aborts")
