import sys, subprocess, collections
rows = [l.split() for l in open(sys.argv[1])]
tot = sum(int(r[2]) for r in rows)
bylib = collections.defaultdict(list)
for r in rows: bylib[r[0]].append((r[1], int(r[2])))
fn = collections.Counter(); ln = collections.Counter()
for lib, lst in bylib.items():
    if lib == "?": fn["?"] += sum(c for _, c in lst); continue
    out = subprocess.run(["addr2line", "-f", "-i" if len(sys.argv) > 3 else "-f", "-e", lib] + ["0x" + a for a, _ in lst], capture_output=True, text=True).stdout.splitlines()
    if len(sys.argv) > 3:   # inline mode: variable number of lines; fall back to non-inline
        out = subprocess.run(["addr2line", "-f", "-e", lib] + ["0x" + a for a, _ in lst], capture_output=True, text=True).stdout.splitlines()
    for k, (a, c) in enumerate(lst):
        f = out[2 * k] if 2 * k < len(out) else "?"; l = out[2 * k + 1] if 2 * k + 1 < len(out) else "?"
        fn[f] += c; ln[l.split("/")[-1].split(" ")[0]] += c
print("total samples", tot)
for f, c in fn.most_common(25): print("%6.2f%%  %s" % (100 * c / tot, f))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("--- lines")
for l, c in ln.most_common(n): print("%6.2f%%  %s" % (100 * c / tot, l))
