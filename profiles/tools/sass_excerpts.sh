#!/bin/bash
# SASS evidence of the hot kernels (no GPU needed):  bash profiles/tools/sass_excerpts.sh > profiles/r2_sass_excerpts.txt
# Per kernel: registers / shared memory, instruction count, and the counts of the mnemonics that
# characterise it (LDG.E.NA.128.CONSTANT = ld.global.nc.L1::no_allocate.v4, warp reductions / votes /
# shuffles / MATCH, LDGSTS = cp.async, atomics; no UTMALDG / UTCMMA: nothing here is a tile copy or a contraction); then the first lines of the streaming loop of K1.
lib=offsetguided_b200/libogdecoder.so
echo "# $(cuobjdump -lelf $lib | head -3 | tr '\n' ' ')"
for k in nms_candidates_kernelILb1 select_topk_kernel amax_scan_kernelIfLb1ELb1 block_list_kernel \
         'fused_block_kernelIfLi4ELb1ELb1' 'limb_score_kernelILb0ELi32' group_warp_kernel group_kernel; do
    sym=$(cuobjdump -res-usage $lib 2>/dev/null | grep -o "Function [^:]*$k[^:]*" | head -1 | cut -d' ' -f2)
    [ -z "$sym" ] && continue
    echo "== $sym"
    cuobjdump -res-usage $lib 2>/dev/null | grep -A1 "Function $sym:" | tail -1
    cuobjdump -sass -fun "$sym" $lib > /tmp/k.sass 2>/dev/null
    echo "   instructions: $(grep -c '^\s*/\*[0-9a-f]\{4\}\*/' /tmp/k.sass)"
    for m in 'LDG.E.NA.128.CONSTANT' 'LDG.E.128' 'LDG.E.CONSTANT' 'LDG.E.64' 'LDGSTS' 'LDS.128' 'STS.128' 'REDUX' 'VOTE' \
             'SHFL' 'MATCH' 'ATOMG' 'ATOMS' 'RED.E' 'BAR.SYNC' 'WARPSYNC' 'FFMA' 'FMNMX' 'DFMA' 'UTMALDG' 'UTCMMA'; do
        c=$(grep -c " $m" /tmp/k.sass)
        [ "$c" != "0" ] && echo "   $m: $c"
    done
done
echo
echo "== K1 streaming loads (nms_candidates_kernel<true>): the eight 128-bit loads of a chunk and the hot-row test"
sym=$(cuobjdump -res-usage $lib 2>/dev/null | grep -o "Function [^:]*nms_candidates_kernelILb1[^:]*" | head -1 | cut -d' ' -f2)
cuobjdump -sass -fun "$sym" $lib 2>/dev/null | grep -E "LDG.E.NA.128|FMNMX3|REDUX|FSETP.GE" | head -28 | sed 's#/\* 0x[0-9a-f]* \*/##'
