import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e_chunked"]["ms_per_step"], d["e2e_ray_buffers"]["ms_per_step"], d["clocks"])
