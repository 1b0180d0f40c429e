#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_span_columns.py tests/test_gpu_string_codecs.py -q -x 2>&1 | tail -25
