#!/bin/bash
timeout 300 python tools/debug_mixed.py 2>&1 | grep -v " ok " | head -20
echo "---- done"
timeout 600 python -m pytest tests/test_gpu_string_codecs.py tests/test_gpu_cs_codecs.py -q 2>&1 | tail -4
