#!/bin/bash
O=gpurun_out/${1:-lat}
mkdir -p $O
run() { echo "# $*" | tee -a $O/latency.jsonl; env "$@" timeout 300 python scripts/latency_probe.py 50 2>&1 | tail -1 | tee -a $O/latency.jsonl; }
run OPP_PDL=0
run OPP_PDL=1
run OPP_PDL=1 OPP_B200_LIB=$PWD/variants/libopp_pdl0.so
run OPP_PDL=1 OPP_B200_LIB=$PWD/variants/libopp_pdl2.so
run OPP_PDL=0 OPP_B200_TWO_STREAMS=0
run OPP_PDL=1 OPP_B200_TWO_STREAMS=0
run OPP_PDL=1 OPP_B200_TWO_STREAMS=0 OPP_B200_LIB=$PWD/variants/libopp_pdl2.so
