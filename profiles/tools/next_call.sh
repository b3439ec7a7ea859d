#!/bin/bash
bash profiles/tools/final_n1.sh
