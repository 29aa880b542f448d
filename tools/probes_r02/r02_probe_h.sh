#!/bin/bash
set -u
OUT=gpurun_out; TAG=r02h2; mkdir -p $OUT; export PYTHONUNBUFFERED=1
UFM_POW_EXACT=0 timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_suite_pow0.log 2>&1; tail -15 $OUT/${TAG}_suite_pow0.log
