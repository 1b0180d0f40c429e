#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_macro_blocks.py tests/test_golden_vectors.py -x -q 2>&1 | tail -25
