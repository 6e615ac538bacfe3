import os, sys, subprocess
for dbg in ("0", "3", "31"):
    env = dict(os.environ, PASCO_WGRAD_DEBUG=dbg)
    r = subprocess.run([sys.executable, "tools/conv_microbench.py", "--occ", "0.5", "--channels", "64", "--out", "/tmp/x.jsonl"],
                       env=env, capture_output=True, text=True)
    for l in r.stdout.splitlines():
        if '"conv3"' in l and '"fp32"' in l:
            import json
            d = json.loads(l)
            print("dbg", dbg, "wgrad_ms", round(d["wgrad_ms"], 3), "fwd_ms", round(d["fwd_ms"], 3))
