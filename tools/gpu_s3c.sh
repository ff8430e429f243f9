#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" FB_MANIAC_DEBUG=2 timeout -s KILL 120 python tools/decode_once.py mid > gpurun_out/dbg_mid_prof.log 2>&1; echo "rc=$?"; grep -E -A1 "ch (30|36|42|45|48|53|54) .*cycles/symbol" gpurun_out/dbg_mid_prof.log | grep -v "^--"; grep wall gpurun_out/dbg_mid_prof.log; }
run FB_MANIAC_WSLEEP=100
run FB_MANIAC_WSLEEP=0
run FB_MANIAC_WSLEEP=400
run FB_MANIAC_WSLEEP=100 FB_MANIAC_WUSED=4
