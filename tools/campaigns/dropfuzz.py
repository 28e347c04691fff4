import random, subprocess, os, sys
def nals(b):
    idx=[]; i=0
    while True:
        j=b.find(b'\x00\x00\x01',i)
        if j<0: break
        idx.append(j); i=j+3
    return [b[idx[k]:(idx[k+1] if k+1<len(idx) else len(b))] for k in range(len(idx))]
bad=0
for it in range(int(sys.argv[1])):
    r=random.Random(4242+it)
    a=["-W",str(r.choice([2,3,5])),"-H",str(r.choice([2,3])),"-n","30","-s",str(7000+it),"--gop",r.choice(["IP","IPB"]),"--refs",str(r.randint(1,5)),"--idr",str(r.choice([9,40])),"--deblock","0"]
    if r.random()<0.5: a.append("--dpb")
    subprocess.run([os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tools", "gen264"),"-o","/tmp/e264_campaign_df.264"]+a,capture_output=True)
    ns=nals(open("/tmp/e264_campaign_df.264","rb").read())
    keep=[x for k,x in enumerate(ns) if k<4 or r.random()>0.2]
    r.shuffle(keep) if r.random()<0.1 else None
    open("/tmp/e264_campaign_df2.264","wb").write(b''.join(keep))
    try:
        p=subprocess.run(["/tmp/e264_campaign_dec_asan","/tmp/e264_campaign_df2.264","-q"],capture_output=True,timeout=120,env=dict(os.environ,ASAN_OPTIONS="detect_leaks=0"))
        e=p.stderr.decode(errors="replace")
        if "AddressSanitizer" in e or p.returncode<0 or "runtime error" in e:
            bad+=1; print("BAD", it, p.returncode, [l for l in e.splitlines() if "ERROR" in l or "runtime error" in l][:3]); os.replace("/tmp/e264_campaign_df2.264","/tmp/e264_campaign_dfbad_%d.264"%it)
    except subprocess.TimeoutExpired:
        bad+=1; print("TIMEOUT", it)
print("drop fuzz done, bad =", bad)
