# A/B of prebuilt library variants on ONE box: gpurun_variants/*.so are swapped in for libcbsim.so, two quick bench runs each, twice round
for round in 1 2; do for v in gpurun_variants/*.so; do cp $v c-cyberbattlesim_b200/libcbsim.so; echo "== $v"; bash tools/quick_bench.sh 2>&1 | head -1; done; done
