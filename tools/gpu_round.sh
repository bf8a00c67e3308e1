mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r8_pytest.log
tail -n 4 gpurun_out/r8_pytest.log
( timeout 300 python tools/gen_profile.py --span 410 ) > gpurun_out/r8_genprof.json 2> gpurun_out/r8_genprof.err
cut -c1-200 gpurun_out/r8_genprof.json; tail -n 3 gpurun_out/r8_genprof.err
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 26 -c 1 -o gpurun_out/r8_prof python tools/profile_step.py --preroll 24 --gens 50 --launches 4 ) > gpurun_out/r8_ncu_full.log 2>&1
tail -n 3 gpurun_out/r8_ncu_full.log
