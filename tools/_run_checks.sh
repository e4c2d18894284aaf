timeout 600 python -m pytest tests/test_gpu_tapgemm.py tests/test_gpu_resnet_decoder.py -x -q 2>&1 | tail -5
timeout 200 python tools/exp_tapgemm3.py - MV_NO_CONV3=1 2>&1 | tee gpurun_out/exp13.log
