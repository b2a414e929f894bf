# Round-end evidence run (one B200): full GPU suite, smoke, bench (default + the driver's window + reference arm),
# ncu launch list and one --set full capture per step kernel.  Everything lands in gpurun_out/ under the tag $1.
T=${1:-r02g}
O=gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -6 > $O/${T}_gpu_tests.txt
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2 >> $O/${T}_gpu_tests.txt
timeout 400 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
timeout 200 python bench.py --steps 20 --warmup 5 > $O/${T}_bench_driver.json 2> $O/${T}_bench_driver.err
timeout 300 python bench.py --impl reference > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
FAST="--no-cpu --no-vecenv --transition-envs 0"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/${T}_launches.csv \
  python bench.py --steps 20 --warmup 5 --presteps 8 --no-graph $FAST > $O/${T}_ncu_a.log 2>&1
for k in decode_select_kernel observe_kernel decode_gemm_tc_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 75 -c 1 -f -o $O/${T}_$k \
    python bench.py --steps 20 --warmup 5 --no-graph $FAST > $O/${T}_ncu_$k.log 2>&1
done
