#!/bin/bash
cd $GRAFT_REPO_ROOT
for path in three_pass single_pass; do
python tests/tools/diag_reflect.py reflect nodust strict $path rk3 2>&1 | tail -12
done
python tests/tools/diag_reflect.py outflow nodust strict three_pass rk3 2>&1 | tail -6
