#!/bin/bash
set -u
bash tools/gpu_r02a.sh
MASKS="0 1 2 4 6 8" bash tools/gpu_decb_dbg.sh
