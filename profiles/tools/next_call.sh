#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden_track or warmup" 2>&1 | tail -5
