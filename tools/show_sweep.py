#!/usr/bin/env python
import json, sys
for f in sys.argv[1:]:
    for ln in open(f):
        ln = ln.strip()
        if not ln.startswith('{'):
            print(ln); continue
        d = json.loads(ln)
        if 'error' in d:
            print(d); continue
        print(f"{d['tag']:10s} {str(d['opts']):78s} ok={d['ok']} bits={d['bits']} hist={d['hist_ms']:.3f} part={d['part_ms']:.3f} join={d['join_ms']:.3f} tot={d['total_ms']:.3f} frac={d['pass_frac']} joinGB={d['join_GBs']} Gt/s={d['Gtuples_s']}")
