#!/usr/bin/env bash
for cfg in "0 0" "135000 0" "135000 32" "135000 48" "135000 64" "135000 100"; do set -- $cfg
  echo "--- smem $1 grid $2"; NNGP_PANEL_SMEM=$1 NNGP_PANEL_GRID=$2 timeout 120 python tools/fit_once.py 8192 128 2 5
done
for n in 4096 16384 32768; do echo "--- N=$n smem 135000"; NNGP_PANEL_SMEM=135000 timeout 120 python tools/fit_once.py $n 128 2 3; done
