timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"; done
