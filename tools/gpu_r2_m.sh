#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -k 10 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2m_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r2m_tests.log
timeout -k 10 900 python bench.py > gpurun_out/r2m_bench.log 2> gpurun_out/r2m_bench.err; echo "bench exit $?"; tail -c 300 gpurun_out/r2m_bench.err
