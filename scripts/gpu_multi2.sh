#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 500 2>&1 | tail -12
