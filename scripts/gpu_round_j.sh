timeout 900 python -m pytest tests -m gpu -q -k "edt or sharded_driver" > gpurun_out/pytest_r1j.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r1j.log
timeout 600 python scripts/configs_bench.py --tag r1j --only-edt > gpurun_out/configs_r1j.log 2>&1; echo "configs rc=$?"; tail -c 1500 gpurun_out/configs_r1j.log
bash scripts/ncu_flood.sh r1j_flood
