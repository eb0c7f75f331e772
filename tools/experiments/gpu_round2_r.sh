#!/bin/bash
mkdir -p gpurun_out/r2r
timeout 300 python tools/chain_diag.py 64 > gpurun_out/r2r/chain_diag_64.log 2>&1; cat gpurun_out/r2r/chain_diag_64.log
