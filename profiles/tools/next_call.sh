#!/bin/bash
bash profiles/tools/c3_call.sh
