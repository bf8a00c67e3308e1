# One GPU-box call at HEAD: every GPU parity test, smoke, both bench arms, the ncu launch list of the bench command and one
# full ncu capture of the step kernel.
mkdir -p gpurun_out
R=${R:-r3p}
( timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_gpu.log
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${R}_smoke.log
( timeout 600 python bench.py --impl reference ) > gpurun_out/${R}_bench_reference_arm.json 2> gpurun_out/${R}_bench_reference_arm.err; echo "ref rc=$?"
( timeout 1800 python bench.py ) > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench rc=$?"
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-e2e ) > gpurun_out/${R}_launches.log 2>&1; echo "launches rc=$?"
( timeout 900 ncu --set full --clock-control none -k regex:k_step -s 26 -c 1 -f -o gpurun_out/${R}_k_step python tools/profile_step.py --preroll 24 --gens 50 --launches 4 ) > gpurun_out/${R}_ncu_full.log 2>&1; echo "ncu k_step rc=$?"
tail -n 3 gpurun_out/${R}_pytest_gpu.log; tail -n 2 gpurun_out/${R}_smoke.log; cut -c1-300 gpurun_out/${R}_bench.json
