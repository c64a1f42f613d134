import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): 
        print(line[:200]); continue
    d=json.loads(line)
    print("value %.0f ms/step %.2f fp64 %.3f hbm %.4f e2e %.0f solved %.4f kernel_ms %.2f launches %d" % (d["value"], d["ms_per_step"], d["fp64"]["frac"], d["roofline"]["frac"], d["e2e"]["value"], d["solved_frac"], d["roofline"]["kernel_ms_per_step"], d["gpu_launches"]))
