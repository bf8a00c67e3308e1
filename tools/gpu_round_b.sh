mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_stargambit.py -m gpu -x -q ) > gpurun_out/r3b_pytest_sg.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r3b_pytest_sg.log
( timeout 600 python tools/sg_bench.py --game 23 --games 4096 --moves 64 ) > gpurun_out/r3b_sg_bench.jsonl 2> gpurun_out/r3b_sg_bench.err; echo "bench rc=$?"
( timeout 600 python tools/sg_bench.py --game 23 --games 4096 --moves 64 --no-valid ) >> gpurun_out/r3b_sg_bench.jsonl 2>> gpurun_out/r3b_sg_bench.err
( timeout 600 python tools/sg_bench.py --game 20 --games 8192 --moves 64 ) >> gpurun_out/r3b_sg_bench.jsonl 2>> gpurun_out/r3b_sg_bench.err
cat gpurun_out/r3b_sg_bench.jsonl | cut -c1-600; tail -3 gpurun_out/r3b_sg_bench.err
