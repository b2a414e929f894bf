T=r02i
O=gpurun_out
timeout 500 python -m pytest tests/test_gpu_decode_gemm.py tests/test_gpu_golden.py tests/test_gpu_lockstep.py tests/test_gpu_vecenv.py tests/test_gpu_replay.py -q 2>&1 | tail -4 > $O/${T}_gpu_tests.txt
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1 >> $O/${T}_gpu_tests.txt
timeout 400 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > $O/${T}_bench_driver.json 2> $O/${T}_bench_driver.err
