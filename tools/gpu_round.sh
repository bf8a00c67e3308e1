mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r9_pytest.log
tail -n 4 gpurun_out/r9_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r9_smoke.log 2>&1; tail -n 2 gpurun_out/r9_smoke.log
( timeout 400 python bench.py --impl reference --steps 4 --warmup 3 ) > gpurun_out/r9_bench_ref.json 2> gpurun_out/r9_bench_ref.err
( timeout 700 python bench.py ) > gpurun_out/r9_bench.json 2> gpurun_out/r9_bench.err; echo "bench rc=$?" >> gpurun_out/r9_bench.err
cat gpurun_out/r9_bench_ref.json gpurun_out/r9_bench.json; tail -n 3 gpurun_out/r9_bench.err
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r9_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --preroll 6 ) > gpurun_out/r9_ncu_bench.log 2>&1
tail -n 2 gpurun_out/r9_ncu_bench.log | cut -c1-300
