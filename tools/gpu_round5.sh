#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest5.log 2>&1; echo "pytest_rc=$?" | tee -a gpurun_out/pytest5.log
tail -30 gpurun_out/pytest5.log
