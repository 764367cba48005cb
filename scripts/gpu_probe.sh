#!/bin/bash
timeout 60 python -m pytest tests/test_gpu_index.py -m gpu -q -k "empty_arrays" 2>&1 | grep -vE "^$|warnings|Docs" | tail -40 | cut -c1-300
