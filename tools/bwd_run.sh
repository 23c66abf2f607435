timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_legacy.py -x -q -m gpu -k "align" 2>&1 | tail -2
python tools/time_op.py align_bwd C2 20; python tools/time_op.py align_bwd C4 20
