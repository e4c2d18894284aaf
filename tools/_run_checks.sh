timeout 600 python -m pytest tests/test_gpu_tapgemm.py tests/test_gpu_resnet_decoder.py -x -q 2>&1 | tail -4
timeout 200 python tools/exp_tapgemm3.py - MV_C3_SIDE_ONE_GROUP=1 MV_C3_INPLACE=1 2>&1 | grep -v plain | tee gpurun_out/exp15.log
