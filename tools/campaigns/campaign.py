import random, subprocess, sys, os, hashlib
import os; R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
def run(cmd):
    try: return hashlib.md5(subprocess.run(cmd, capture_output=True, timeout=60).stdout).hexdigest()
    except subprocess.TimeoutExpired: return "timeout"
n=int(sys.argv[1]); seed0=int(sys.argv[2]); bad=0
for it in range(n):
    r=random.Random(seed0+it)
    W=r.choice([2,3,4,5,7,11]); H=r.choice([2,3,4,6,9])
    gop=r.choice(["I","IP","IPB","IPB"])
    if W==1 and gop!="I": W=2
    a=["-W",str(W),"-H",str(H),"-n",str(r.choice([4,9,13])),"-s",str(seed0+it),"--gop",gop,"--refs",str(r.randint(1,5)),"--idr",str(r.choice([5,9,17,40])),
       "--t8x8",str(r.choice([0,30,70])),"--scaling",str(r.randint(0,3)),"--wp",str(r.randint(0,2)),"--slices",str(r.randint(1,4)),"--deblock",str(r.choice([0,0,1,2])),
       "--density",str(r.choice([5,25,60])),"--qp",str(r.choice([4,20,28,36,48])),"--pcm",str(r.choice([0,2,40])),"--mvrange",str(r.choice([8,24,90,250])),
       "--intra-pct",str(r.choice([0,10,40])),"--skip-pct",str(r.choice([0,15,50]))]
    if r.random()<0.35: a.append("--cavlc")
    if gop=="IPB":
        if r.random()<0.4: a.append("--temporal")
        if r.random()<0.3: a.append("--bref")
        if r.random()<0.3: a.append("--direct4x4")
        if r.random()<0.3 and "--temporal" not in a: a.append("--dpb")
    if gop=="IP":
        if r.random()<0.4: a.append("--dpb")
        if r.random()<0.3: a += ["--poc-type", str(r.randint(1,2))]
        if r.random()<0.3 and "--dpb" not in a and "--poc-type" not in a: a.append("--nonref-p")
    if gop!="I" and r.random()<0.3: a.append("--mixed-slices")
    if r.random()<0.2: a.append("--ps-update")
    if r.random()<0.2: a.append("--extra-nals")
    if r.random()<0.2: a += ["--crop-left",str(2*r.randint(0,3)),"--crop-top",str(2*r.randint(0,3))]
    p="/tmp/e264_campaign_camp.264"
    g=subprocess.run([R+"/tools/gen264","-o",p]+a, capture_output=True)
    if g.returncode!=0: print("GEN FAIL", " ".join(a), g.stderr[-200:]); continue
    x=run([R+"/oracle/_ref/ref_decode",p,"-c"]); y=run([R+"/oracle/oracle_decode",p,"-c"])
    if x!=y:
        bad+=1; print("DIFF", " ".join(a)); os.replace(p, "/tmp/e264_campaign_camp_diff_%d.264"%(seed0+it))
print("campaign done:", n, "streams,", bad, "mismatching")
