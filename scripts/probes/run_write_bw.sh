#!/bin/bash
# Build and run the HBM bandwidth probe (stores / bulk stores / loads); prints one JSON line per grid size.
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_bw write_bw.cu
timeout 120 ./write_bw
