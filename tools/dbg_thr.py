import sys, hashlib, subprocess, os
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from checkers import decode_bytes
import numpy as np
args = sys.argv[1:] or "-W 5 -H 6 -n 24 -s 503 --gop IPB --refs 3 --idr 13 --slices 4 --mixed-slices --deblock 0 --wp 1".split()
subprocess.run(["tools/gen264","-o","/tmp/m.264"]+args,check=True,stderr=subprocess.DEVNULL)
data=open("/tmp/m.264","rb").read()
port=decode_bytes(data,"port",0)[0]
for nt in (0,2):
    got=decode_bytes(data,"gpu",nt)[0]
    for i,(a,b) in enumerate(zip(got,port)):
        if a[3]!=b[3]:
            x=np.frombuffer(a[3],np.uint8); y=np.frombuffer(b[3],np.uint8); d=np.nonzero(x!=y)[0]
            W=a[1]; ys=d[d<W*a[2]]
            print("nt",nt,"frame",i,"id",a[0],"ndiff",len(d),"luma diffs",len(ys),"first luma (x,y):",[(int(p%W),int(p//W)) for p in ys[:6]], "mbs", sorted(set((int(p%W)//16,int(p//W)//16) for p in ys))[:12])
    print("nt",nt,"done", len(got), len(port))
