#!/bin/bash
# The command of the next gpurun call (edited between calls; the snapshot is taken when the call gets its box).
bash profiles/tools/run_n1.sh
bash profiles/tools/run_n1_extra.sh
