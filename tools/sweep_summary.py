import json,sys
for f in sys.argv[1:]:
    rows=[json.loads(l) for l in open(f)]
    by={}
    for r in rows: by.setdefault(r["N"],{})[r["op"]]=r["frac"]
    print(f)
    for n,d in by.items(): print("  ",n,d)
