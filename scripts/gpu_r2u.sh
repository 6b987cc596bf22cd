#!/bin/bash
LIB=visual-odometry-rs_b200/lib/libvors_b200.so
echo "== current"; python scripts/diag_determinism.py 2>&1 | tail -5
cp $LIB /tmp/stock.so; cp visual-odometry-rs_b200/lib_variants/oldserial.so $LIB
echo "== old serial"; python scripts/diag_determinism.py 2>&1 | tail -5
cp /tmp/stock.so $LIB
