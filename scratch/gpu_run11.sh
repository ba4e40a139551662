#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q -k "boost or shim or tdt or ctc" > gpurun_out/r02_pytest_boost.log 2>&1
tail -30 gpurun_out/r02_pytest_boost.log
