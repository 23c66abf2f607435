timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_merged.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_merged.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_us'], d['verified']['ok'], d['other_repool_form'], d['gpu_launches'])"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
