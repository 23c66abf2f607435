timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_legacy.py -x -q -m gpu -k "nms or proposal or detect" 2>&1 | tail -2
timeout 120 python tools/time_proposal.py C1 30
