#!/bin/bash
# 1-GPU visit: dictionary surface + adapter + the datum test with its message, then the whole gpu suite
timeout 600 python -m pytest tests/test_gpu_dict_surface.py tests/test_gpu_block_api.py -x -q 2>&1 | tail -40
timeout 300 tests/cpp/test_host_adapter | tail -12
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
