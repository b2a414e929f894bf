# A/B of prebuilt library variants on ONE box: gpurun_variants/*.so are swapped in for libcbsim.so and benched back to back
# (ROUNDS passes over all variants, default 2; run-to-run noise on one box is ~0.1 %)
for round in $(seq 1 ${ROUNDS:-2}); do for v in gpurun_variants/*.so; do cp $v c-cyberbattlesim_b200/libcbsim.so; echo "== $v"; bash tools/quick_bench.sh 2>&1 | head -1; done; done
