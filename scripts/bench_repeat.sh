# GPU diagnostic: run the default bench a few times in one box and print the headline and the per-step
# ResNet-50 times, to see run-to-run spread
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 120 python bench.py > gpurun_out/bench_rep_$i.json 2> gpurun_out/bench_rep_$i.err || tail -5 gpurun_out/bench_rep_$i.err
  python - gpurun_out/bench_rep_$i.json <<'PY'
import sys, json
t = open(sys.argv[1]).read().strip().splitlines()
if t:
  d = json.loads(t[-1]); r = d["shampoo_step_resnet50"]
  print(round(d["value"]), round(d["e2e"]["value"]), round(r["ms"], 2), r["ms_per_step_list"], round(d["shampoo_step"]["ms"], 3))
PY
done
