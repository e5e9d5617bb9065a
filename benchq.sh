timeout -k 10 200 python bench.py --steps 10 --warmup 3 --no-variants --no-cpu-baseline --no-e2e > gpurun_out/q.json 2> gpurun_out/q.err; tail -c 300 gpurun_out/q.err; python - <<PY
import json
d=json.load(open("gpurun_out/q.json"))
print("RESULT", round(d["value"]), round(d["ms_per_step"],4), {k:round(v["ms_per_step"],4) for k,v in d["stages"].items()})
PY
