#!/bin/bash
bash scripts/r02_ncu.sh gpurun_out/c5 c3_full quad_sweep_kernelILi0ELb1ELb1 c3 full 18944
bash scripts/r02_ncu.sh gpurun_out/c5 c3_red quad_sweep_kernelILi2ELb1ELb1 c3 reduced 18944
bash scripts/r02_ncu.sh gpurun_out/c5 c3_mt quad_sweep_kernelILi0ELb1ELb1 c3 full 18944 --mt
bash scripts/r02_ncu.sh gpurun_out/c5 c2_full quad_sweep_kernelILi0ELb1ELb1 c2 full 262144
bash scripts/r02_ncu.sh gpurun_out/c5 c5_red quad_sweep_kernelILi2ELb1ELb0 c5 reduced 37888 --spl 4
ls -la gpurun_out/c5
