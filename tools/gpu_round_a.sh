mkdir -p gpurun_out
R=r3a
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_gpu.log
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${R}_smoke.log
( timeout 600 python bench.py --impl reference ) > gpurun_out/${R}_bench_reference_arm.json 2> gpurun_out/${R}_bench_reference_arm.err; echo "ref rc=$?"
( timeout 1200 python bench.py ) > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench rc=$?"
tail -n 3 gpurun_out/${R}_pytest_gpu.log; tail -n 2 gpurun_out/${R}_smoke.log; cat gpurun_out/${R}_bench.json | cut -c1-400
