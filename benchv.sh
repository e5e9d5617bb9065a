# A/B timing of experiment builds (manuscript-ocr_b200/build.py --variant NAME ...): bash benchv.sh NAME [NAME ...]
# ("base" = the product library).  Prints one RESULT line per library; parity is not checked (variants may break it).
for v in "$@"; do
  if [ "$v" = base ]; then unset MS_B200_LIB; else export MS_B200_LIB=$PWD/manuscript-ocr_b200/manuscript_b200/libvariant_$v.so; fi
  timeout -k 10 200 python bench.py --steps 10 --warmup 3 --no-variants --no-cpu-baseline --no-e2e --no-corpus --no-parity > gpurun_out/v_$v.json 2> gpurun_out/v_$v.err || tail -c 300 gpurun_out/v_$v.err
  python - "$v" <<PY
import json, sys
v = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/v_{v}.json"))
    print("RESULT", v, round(d["value"]), round(d["ms_per_step"], 4), {k: round(s["ms_per_step"], 4) for k, s in d["stages"].items()})
except Exception as e:
    print("RESULT", v, "failed", e)
PY
done
