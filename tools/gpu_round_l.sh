mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_stargambit_search.py tests/test_zz_pybind_tafl_pm.py tests/test_tafl_selfplay.py -m gpu -x -q ) > gpurun_out/r3l_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r3l_pytest.log
