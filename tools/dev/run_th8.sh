#!/bin/bash
timeout 300 python tools/dev/th8_check.py 2>&1 | grep "H=8.*s16=1"
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -k "talking or cfg2 or cfg1" -x -q -p no:cacheprovider 2>&1 | tail -2
