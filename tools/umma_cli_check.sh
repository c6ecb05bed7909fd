#!/bin/bash
# dev tool: whole runs of the drop-in CLI with the tcgen05 dense kernels (BLK_DENSE=umma) and with the IMMA
# kernels (BLK_DENSE=mma); both kernel files must be byte-identical to the one the unmodified sequential
# reference wrote for the same command line (hashes in tools/data/cases.txt, made by tools/make_cli_cases.py).
# No Python on the box.
cd "$(dirname "$0")/.."
DRV=block-lanczos-algorithm-parallelization_b200/driver/lanczos_modp
out=$(mktemp -d)
fail=0
while read m p n side want iters; do
        tag="${m}_p${p}_n${n}_${side#--}"
        for mode in umma mma; do
                BLK_DENSE=$mode timeout 120 $DRV --matrix tools/data/$m.mtx --prime $p --n $n $side --output-file $out/$mode.mtx > $out/$mode.log 2>&1 \
                        || { echo "$tag: $mode run failed: $(tail -2 $out/$mode.log)"; fail=1; }
        done
        su=$(sha256sum < $out/umma.mtx | cut -d' ' -f1); sm=$(sha256sum < $out/mma.mtx | cut -d' ' -f1)
        it=$(grep -o "after [0-9]* iterations" $out/umma.log | head -1)
        if [ "$su" == "$want" ] && [ "$sm" == "$want" ]; then echo "$tag: umma == mma == reference ($it, reference $iters)"
        else echo "$tag: MISMATCH umma=${su:0:12} mma=${sm:0:12} ref=${want:0:12} ($it, reference $iters)"; fail=1; fi
done < tools/data/cases.txt
rm -rf $out
[ $fail == 0 ] && echo "umma_cli_check: PASS" || echo "umma_cli_check: FAILED"
exit $fail
