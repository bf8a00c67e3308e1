mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r10_pytest.log
tail -n 6 gpurun_out/r10_pytest.log
( timeout 700 python bench.py --no-cpu-baseline --steps 5 ) > gpurun_out/r10_bench.json 2> gpurun_out/r10_bench.err; echo "bench rc=$?" >> gpurun_out/r10_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r10_bench.json')); print(d['value'], d['e2e'], d['e2e_nn_host'], d['e2e_nn_device'])"; tail -n 5 gpurun_out/r10_bench.err
