#!/bin/bash
cd $GRAFT_REPO_ROOT
python scripts/diag_reflect.py reflect dust 2>&1 | tail -40
python scripts/diag_reflect.py reflect nodust 2>&1 | tail -20
AB200_NO_SWEEP=1 python scripts/diag_reflect.py reflect dust 2>&1 | tail -20
