#!/bin/bash
mkdir -p gpurun_out
for P in nan big none; do for T in 99 27; do POISON=$P PT=$T python scripts/uninit_probe.py 2>/dev/null | cut -c1-700; done; done | tee gpurun_out/r02_uninit_probe.log
