#!/usr/bin/env python3
"""Summarise an E264B_TRACE timeline (written by e264b_replay): per-kernel durations under contention,
per-stream critical path, and how many launches of each kind overlap on average."""
import csv, sys, collections
rows = [dict(r) for r in csv.DictReader(open(sys.argv[1]))]
for r in rows:
    for k in r: r[k] = int(r[k])
names = {0: "residual", 1: "inter", 2: "intra", 3: "deblock"}
t0 = min(r["start_ns"] for r in rows); t1 = max(r["end_ns"] for r in rows)
span = t1 - t0
print(f"launches {len(rows)}  span {span/1e6:.2f} ms  streams {1+max(r['stream'] for r in rows)}  pictures/stream {1+max(r['pic'] for r in rows)}")
by = collections.defaultdict(list)
for r in rows: by[r["kind"]].append(r["end_ns"] - r["start_ns"])
for k, v in sorted(by.items()):
    v.sort()
    print(f"  {names[k]:9s} n={len(v):5d}  mean {sum(v)/len(v)/1e3:8.1f} us  median {v[len(v)//2]/1e3:8.1f}  p90 {v[int(len(v)*.9)]/1e3:8.1f}  busy-sum/span = {sum(v)/span:6.2f} concurrent")
# per stream: time from first kernel start of picture k to last kernel end, and the gap to the next picture
lat = []; gap = []
pics = collections.defaultdict(list)
for r in rows: pics[(r["stream"], r["rep"], r["pic"])].append(r)
keys = sorted(pics)
prev_end = {}
for key in keys:
    rs = pics[key]; s = min(r["start_ns"] for r in rs); e = max(r["end_ns"] for r in rs)
    lat.append(e - s)
    if key[0] in prev_end: gap.append(s - prev_end[key[0]])
    prev_end[key[0]] = e
lat.sort(); gap.sort()
print(f"  picture latency (first start -> last end): mean {sum(lat)/len(lat)/1e3:.1f} us  median {lat[len(lat)//2]/1e3:.1f}")
if gap: print(f"  gap between a stream's consecutive pictures: mean {sum(gap)/len(gap)/1e3:.1f} us  median {gap[len(gap)//2]/1e3:.1f}  p90 {gap[int(len(gap)*.9)]/1e3:.1f}")
print(f"  pictures/s = {len(keys)/span*1e9:.0f}")
