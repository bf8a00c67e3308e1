mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_stargambit_search.py tests/test_tafl_selfplay.py tests/test_zz_pybind_tafl_pm.py -m gpu -x -q ) > gpurun_out/r3g_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r3g_pytest.log
for n in 1024 4096 8192; do
( timeout 600 python tools/tafl_selfplay_bench.py --game 23 --games $n --moves 16 --cpu-seconds 1 ) >> gpurun_out/r3g_sg_selfplay.jsonl 2>> gpurun_out/r3g_sg_selfplay.err; echo "sg selfplay $n rc=$?"
done
( timeout 600 python tools/tafl_selfplay_bench.py --game 20 --games 8192 --moves 16 --cpu-seconds 1 ) >> gpurun_out/r3g_sg_selfplay.jsonl 2>> gpurun_out/r3g_sg_selfplay.err
python - <<'PY'
import json
for l in open('gpurun_out/r3g_sg_selfplay.jsonl'):
    d=json.loads(l); print(d['workload'][:80], round(d['simulations_per_second']/1e6,2), 'M sims/s', round(d['moves_per_second']), 'moves/s')
PY
tail -3 gpurun_out/r3g_sg_selfplay.err
