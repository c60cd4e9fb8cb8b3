import json,sys
d=json.load(open(sys.argv[1]))
for r in d["step_host"]:
    t=r["trace_us_last_call"]
    print(r["format"], r["host_threads"], r["chunks"], r.get("actions_direct"), round(r["us_per_slot"],1), round(r["agent_steps_per_s"]/1e6), t[:4], t[-3:])
