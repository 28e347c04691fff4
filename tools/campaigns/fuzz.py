import random, subprocess, sys, os
src = open(sys.argv[1], 'rb').read()
n = int(sys.argv[2]); seed0 = int(sys.argv[3]) if len(sys.argv) > 3 else 0
bad = 0
for it in range(n):
    rnd = random.Random(seed0 + it)
    b = bytearray(src)
    k = rnd.choice([1, 2, 4, 16])
    for _ in range(k):
        pos = rnd.randrange(40, len(b))
        mode = rnd.randrange(3)
        if mode == 0: b[pos] ^= 1 << rnd.randrange(8)
        elif mode == 1: b[pos] = rnd.randrange(256)
        else:
            ln = rnd.randrange(1, 64); b[pos:pos+ln] = bytes(rnd.randrange(256) for _ in range(ln))
    if rnd.randrange(4) == 0: b = b[:rnd.randrange(100, len(b))]
    open('/tmp/e264_campaign_f.264', 'wb').write(b)
    try:
        r = subprocess.run(['/tmp/e264_campaign_dec_asan', '/tmp/e264_campaign_f.264', '-q'], capture_output=True, timeout=120, env=dict(os.environ, ASAN_OPTIONS='detect_leaks=0', UBSAN_OPTIONS='print_stacktrace=0'))
        err = r.stderr.decode(errors='replace')
        if 'AddressSanitizer' in err or r.returncode < 0:
            bad += 1
            print('SEED', seed0 + it, 'rc', r.returncode)
            for l in err.splitlines():
                if 'ERROR' in l or '#0' in l or '#1 ' in l or '#2 ' in l or 'SUMMARY' in l: print('   ', l[:160])
            os.rename('/tmp/e264_campaign_f.264', '/tmp/e264_campaign_crash_%d.264' % (seed0 + it))
            if bad >= 3: break
        elif 'runtime error' in err:
            ls = sorted(set(l.split('runtime error:')[0].split('/')[-1] + l.split('runtime error:')[1][:60] for l in err.splitlines() if 'runtime error' in l))
            print('UB', seed0 + it, ls[:4])
    except subprocess.TimeoutExpired:
        print('TIMEOUT seed', seed0 + it); os.rename('/tmp/e264_campaign_f.264', '/tmp/e264_campaign_hang_%d.264' % (seed0 + it)); bad += 1
print('done, bad =', bad)
