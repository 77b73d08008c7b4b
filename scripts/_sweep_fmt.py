import sys,json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l.strip()); continue
    print(r["len"], r["pairs"], "dist dp %.2f ms %.0f GCUPS dev %.0f | path dp %.2f tb %.2f dev %.0f GCUPS | cpu %.0f GCUPS" % (r["distance"]["dp_ms"], r["distance"]["dp_gcups"], r["distance"]["device_gcups"], r["path"]["dp_ms"], r["path"]["traceback_ms"], r["path"]["device_gcups"], r["myers_cpu"]["gcups"]))
