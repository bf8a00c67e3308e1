# One GPU-box call: parity tests, smoke, bench both arms, ncu launch list, ncu full captures of k_step, k_forest_simulate and k_sp_search.
mkdir -p gpurun_out
R=${R:-r37}
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_gpu.log
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${R}_smoke.log
( timeout 600 python bench.py --impl reference ) > gpurun_out/${R}_bench_reference_arm.json 2> gpurun_out/${R}_bench_reference_arm.err; echo "ref rc=$?"
( timeout 1200 python bench.py ) > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench rc=$?"
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra ) > gpurun_out/${R}_launches.log 2>&1; echo "launches rc=$?"
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 26 -c 1 -f -o gpurun_out/${R}_k_step python tools/profile_step.py --preroll 24 --gens 50 --launches 4 ) > gpurun_out/${R}_ncu_full.log 2>&1; echo "ncu k_step rc=$?"
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_forest_simulate -s 3 -c 1 -f -o gpurun_out/${R}_k_forest_simulate python tools/forest_bench.py --game 0 --trees 16384 --moves 4 --gumbel-m 16 ) > gpurun_out/${R}_ncu_forest.log 2>&1; echo "ncu forest rc=$?"
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sp_search -s 5 -c 1 -f -o gpurun_out/${R}_k_sp_search python tools/tafl_selfplay_bench.py --game 0 --games 8192 --moves 8 --cpu-seconds 1 ) > gpurun_out/${R}_ncu_selfplay.log 2>&1; echo "ncu selfplay rc=$?"
tail -n 3 gpurun_out/${R}_pytest_gpu.log; tail -n 2 gpurun_out/${R}_smoke.log; cat gpurun_out/${R}_bench.json | cut -c1-400
