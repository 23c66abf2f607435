timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_legacy.py -x -q -m gpu -k "nms" 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:k_align8_bwd_own -s 2 -c 1 -o gpurun_out/prof_bwd_own_C2 -f python tools/prof_op.py align_bwd C2 > gpurun_out/ncu_bwd.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:k_align8_bwd_own -s 2 -c 1 -o gpurun_out/prof_bwd_own_C4 -f python tools/prof_op.py align_bwd C4 > gpurun_out/ncu_bwd4.log 2>&1; echo rc=$?
ncu --set full --clock-control none -k regex:"k_nms_scan3|k_nms_mask_rm" -s 4 -c 2 -o gpurun_out/prof_nms_C1 -f python tools/prof_op.py proposal C1 > gpurun_out/ncu_nms.log 2>&1; echo rc=$?
