#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_reduce.py -m gpu -q -x --timeout 600 2>&1 | tail -8
