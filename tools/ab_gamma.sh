#!/bin/bash
# development aid (under gpurun): first- vs second-generation Gamma stage, kernel-set time
for v in 1 0; do
  echo "== LWB200_GAMMA_V1=$v"
  LWB200_GAMMA_V1=$v python tools/prof_c3.py 512 4 c3
  LWB200_GAMMA_V1=$v LWB200_GAMMA_DIRECT=0 python tools/prof_c3.py 1 12 c2
done
