timeout 600 python -m pytest tests/test_gpu_tapgemm.py tests/test_gpu_resnet_decoder.py tests/test_gpu_resnet_encoder.py -x -q 2>&1 | tail -5
timeout 200 python tools/exp_tapgemm3.py - MV_C3_ONE_GROUP=1 2>&1 | tee gpurun_out/exp14.log
