timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_legacy.py -x -q -m gpu -k "align or step or bench or pool" 2>&1 | tail -2
python tools/time_merge.py C4; python tools/time_merge.py C4x3
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-ops 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['verified']['ok'])"
