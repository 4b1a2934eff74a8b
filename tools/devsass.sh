#!/bin/bash
# Development helper: compile k_pdm.cu with only the order-2 instances and print the
# opcode histogram of k_pdm_v2_ws2<2,3,FORM,P>.  usage: tools/devsass.sh FORM P [full]
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DPDM_DEV_FAST -I include -c synth_tools_b200/csrc/k_pdm.cu -o /tmp/k_pdm_dev.o || exit 1
tools/sass.sh /tmp/k_pdm_dev.o "_Z12k_pdm_v2_ws2ILi2ELi3ELi${1}ELi${2}ELi${NS:-2}EEv11PdmV2Params13PdmV2Ws2Extra" > /tmp/dev_sass.txt
if [ "$3" = full ]; then cat /tmp/dev_sass.txt; else awk '{print $1}' /tmp/dev_sass.txt | sed 's/;//' | sort | uniq -c | sort -rn | head -${3:-10}; fi
