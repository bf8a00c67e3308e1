mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_stargambit_search.py tests/test_forest.py tests/test_tafl_selfplay.py -m gpu -x -q ) > gpurun_out/r3c_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r3c_pytest.log
