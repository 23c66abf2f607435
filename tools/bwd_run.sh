python tools/time_merge.py C4; python tools/time_merge.py C4x3
python tools/time_op.py align_fwd C2 20
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-ops 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['verified']['ok'])"
