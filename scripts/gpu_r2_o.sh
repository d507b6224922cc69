#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_tearfree.py -q -m gpu -x 2>&1 | tail -30
