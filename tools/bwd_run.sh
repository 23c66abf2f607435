python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/time_op.py align_bwd C2 20; python tools/time_op.py align_bwd C4 20
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_s2b.log 2>&1; echo bench rc=$?; tail -1 gpurun_out/bench_s2b.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac']); print(json.dumps(d.get('ops'))[:3000])"
