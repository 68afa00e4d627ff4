import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); k=d["roofline"]["kernels"]
        print(f, round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), {a:round(b["avg_us"],1) for a,b in k.items()})
    except Exception as e:
        print(f, 'ERR', e, open(f).read()[-600:])
