import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f)); print(f, round(d["value"]), {k:round(v*1e3,1) for k,v in d["kernel_ms"].items()}, 'e2e',round(d["e2e"]["value"]))
    except Exception as e: print(f, 'ERR', e)
