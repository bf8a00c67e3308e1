mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_pybind_dlpack.py tests/test_pybind_module.py tests/test_cabi.py -m gpu -x -q ) > gpurun_out/r3e_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r3e_pytest.log
