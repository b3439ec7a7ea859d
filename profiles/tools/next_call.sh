#!/bin/bash
timeout 200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_pytest.log; tail -3 gpurun_out/r2_pytest.log
