import random, subprocess, sys, hashlib
import os; R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
def out(cmd):
    try: return subprocess.run(cmd,capture_output=True,timeout=60).stdout
    except subprocess.TimeoutExpired: return b"timeout"
bad=0; n=int(sys.argv[1])
for it in range(n):
    r=random.Random(555+it); parts=[]
    sizes=[(r.choice([2,4,6]), r.choice([2,3,5])) for _ in range(2)]
    for k in range(r.randint(2,5)):
        W,H=r.choice(sizes)
        a=["-W",str(W),"-H",str(H),"-n",str(r.choice([3,6,11])),"-s",str(9000+it*10+k),"--gop",r.choice(["I","IP","IPB"]),"--refs",str(r.randint(1,5)),"--idr",str(r.choice([4,40])),"--deblock","0"]
        if r.random()<0.3: a.append("--cavlc")
        if r.random()<0.3: a += ["--crop-bottom", str(2*r.randint(0,3))]
        subprocess.run([R+"/tools/gen264","-o","/tmp/e264_campaign_part.264"]+a,capture_output=True)
        parts.append(open("/tmp/e264_campaign_part.264","rb").read())
    open("/tmp/e264_campaign_cat.264","wb").write(b"".join(parts))
    x=out([R+"/oracle/_ref/ref_decode","/tmp/e264_campaign_cat.264","-c"]); y=out([R+"/oracle/oracle_decode","/tmp/e264_campaign_cat.264","-c"])
    if x!=y: bad+=1; print("DIFF", it); open("/tmp/e264_campaign_cat_diff_%d.264"%it,"wb").write(b"".join(parts))
print("concatenation campaign:", n, "streams,", bad, "mismatching")
