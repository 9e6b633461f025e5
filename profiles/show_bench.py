import json, sys
for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:
        print(path, "unreadable", e); continue
    print(path, "value=%.0f %s ms/step=%.3f e2e=%.0f frac=%.3f n=%d check=%s" % (
        d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["n_gpus"], d.get("oracle_check_max_norm_err")))
