import json,sys
for f in sys.argv[1:]:
    for l in open(f):
        try: d=json.loads(l)
        except Exception: print(f, l[:200]); continue
        print("%-34s %-8s x%-3d fwd %7.1f (%.2f) bwd %7.1f (%.2f) frac %.3f"%(f.split('/')[-1], d["case"], d["scale"], d["fwd_us"],d["fwd_frac"],d["bwd_us"],d["bwd_frac"],d["frac"]))
