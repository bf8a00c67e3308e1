mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -5 gpurun_out/r2_pytest.log
( timeout 300 python tools/gen_profile.py ) >> gpurun_out/r2_genprof.json 2>> gpurun_out/r2_genprof.err
( timeout 300 python bench.py --no-cpu-baseline --steps 5 ) > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
cat gpurun_out/r2_genprof.json gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_genprof.err gpurun_out/r2_bench.err
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 26 -c 1 -o gpurun_out/r2_prof python tools/profile_step.py --preroll 24 --gens 50 --launches 4 ) > gpurun_out/r2_ncu_full.log 2>&1
tail -3 gpurun_out/r2_ncu_full.log
