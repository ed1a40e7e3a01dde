#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/loop_check.py 50 2>&1 | tail -3
timeout 600 python tools/layered_demo.py 50 4 128 2>&1 | tail -8
