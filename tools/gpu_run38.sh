#!/bin/bash
cd "$(dirname "$0")/.."
for i in 1 2 3 4; do python tools/prof_acq_e2e.py 5000 2>&1 | grep "^host" | awk '{print $1, $2, $3, $4, $5, $6, $7}' | tr '\n' ' '; echo; done
