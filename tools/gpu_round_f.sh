mkdir -p gpurun_out
( timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra ) > gpurun_out/r3f_bench_short.json 2> gpurun_out/r3f_bench_short.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3f_bench_short.json').read().strip().splitlines()[-1])
print('value',d['value']); print('pybind',d['e2e_nn_pybind_dlpack']); print('reference',d['e2e_nn_reference']); print('nn_device',d['e2e_nn_device']['value'])
PY
tail -3 gpurun_out/r3f_bench_short.err
( timeout 600 python tools/tafl_selfplay_bench.py --game 23 --games 1024 --moves 16 --cpu-seconds 5 ) > gpurun_out/r3f_sg_selfplay.jsonl 2> gpurun_out/r3f_sg_selfplay.err; echo "sg selfplay rc=$?"
( timeout 600 python tools/tafl_selfplay_bench.py --game 20 --games 1024 --moves 16 --cpu-seconds 5 ) >> gpurun_out/r3f_sg_selfplay.jsonl 2>> gpurun_out/r3f_sg_selfplay.err
cut -c1-900 gpurun_out/r3f_sg_selfplay.jsonl; tail -3 gpurun_out/r3f_sg_selfplay.err
