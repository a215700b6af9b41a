mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/pytest_full5.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/pytest_full5.log
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/smoke_r02.log
