# end-of-round evidence run (1 GPU): tests, smoke, bench (+cpu baseline), reference arm, ncu launch list, ncu full on the dominant kernels
mkdir -p gpurun_out /tmp/n
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; tail -2 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
MV_BENCH_DUMP=gpurun_out/kernel_times.json timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; head -c 300 gpurun_out/bench.json; echo
timeout 300 python bench.py --impl reference > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; head -c 400 gpurun_out/bench_reference.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv python tools/ncu_step.py 256 2 > gpurun_out/ncu_step.log 2>&1
python tools/ncu_launch_summary.py gpurun_out/launches.csv "# ncu launch list, two eager north-star training steps (B=256/GPU): ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 python tools/ncu_step.py 256 2" > gpurun_out/launch_summary.md; head -14 gpurun_out/launch_summary.md
timeout 600 ncu --set full --clock-control none -k regex:"conv3_kernel|head3_kernel|wgrad_kernel|tapgemm_kernel|lpx" -c 60 -o /tmp/n/dec -f python tools/ncu_decoder.py > gpurun_out/ncu_decoder.log 2>&1
ncu -i /tmp/n/dec.ncu-rep --page raw --csv > gpurun_out/decoder_full_raw.csv
ls -la gpurun_out | tail -12
