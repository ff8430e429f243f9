#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" FB_MANIAC_DEBUG=1 timeout -s KILL 200 python tools/decode_once.py cfg2 > gpurun_out/dbg_cfg2.log 2>&1; grep -E "ch (48|54|59|60) .*cycles/symbol" gpurun_out/dbg_cfg2.log; grep wall gpurun_out/dbg_cfg2.log; }
run FB_X=1

