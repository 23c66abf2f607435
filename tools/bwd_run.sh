timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_s2c.log 2>&1; echo bench rc=$?; tail -1 gpurun_out/bench_s2c.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac']); r=d['ops']['rows']; print({k:r[k] for k in r if 'nms' in k or 'Proposal' in k})"
