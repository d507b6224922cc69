mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_default_s2.json 2> gpurun_out/bench_default_s2.err; echo "bench rc=$?"; cut -c1-2600 gpurun_out/bench_default_s2.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_s2.json 2>/dev/null; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_reference_s2.json
