#!/bin/bash
for d in 0 32; do HVIT_DBG=$d timeout 120 python tests/gemm_probe.py 2>&1 | grep "proj\|fc2"; done
