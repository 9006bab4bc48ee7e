#!/bin/bash
timeout 300 python tests/attn_probe.py 2>&1 | grep "^B=" 
HVIT_PROF=1 timeout 300 python tests/attn_probe.py 2>&1 | grep "prof" | head -12
