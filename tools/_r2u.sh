mkdir -p gpurun_out/r2u; O=gpurun_out/r2u
FLATDD_B200_BLOCK_SKIP=2 ncu --set full --import-source on --clock-control none -k regex:dmavm_block_ws -s 2 -c 1 -o $O/one_block_skip2_w8 python tools/block_ablate.py 26 "3,7,12,20" > $O/ncu1.log 2>&1
FLATDD_B200_LIB=build/variants/w12.so FLATDD_B200_BLOCK_SKIP=2 ncu --set full --import-source on --clock-control none -k regex:dmavm_block_ws -s 2 -c 1 -o $O/one_block_skip2_w12 python tools/block_ablate.py 26 "3,7,12,20" > $O/ncu2.log 2>&1
tail -2 $O/ncu1.log $O/ncu2.log
