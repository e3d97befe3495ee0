import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("value %.4g pairs/s  ms/step %.2f"%(d["value"], d["ms_per_step"])); print({k:round(v,2) for k,v in d["phases_ms_rank0"].items()}); print("roofline frac %.4f achieved %.1f GB/s"%(d["roofline"]["frac"], d["roofline"]["achieved"])); print(d["counts_rank0"]); print("spmv", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["spmv"].items() if k!="note"}); print("e2e", d["e2e"])
    else: print(l.rstrip())
