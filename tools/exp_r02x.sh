#!/bin/bash
# round-2 run X: smoke() as the driver runs it
set -u
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
