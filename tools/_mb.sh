timeout 200 python tools/conv_microbench.py --occ 0.1 0.5 --channels 64 128 2>&1 | grep "conv3\|rror" | grep -v simt | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l[l.index('{'):])
    except Exception: print(l[:200]); continue
    print(d['N'], d['C'], d['precision'], 'fwd %.3f dgrad %.3f wgrad %.3f' % (d['fwd_ms'], d['dgrad_ms'], d['wgrad_ms']))
"
