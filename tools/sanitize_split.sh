#!/bin/bash
# compute-sanitizer on the split launch (cluster kernel + fine-level kernel of one frame) and on the frame preparation after the
# border-pixel lists; logs -> <outdir>/r02_sanitizer_*.log.  usage: tools/sanitize_split.sh <outdir>
out=${1:-gpurun_out}
mkdir -p $out
cs="compute-sanitizer --print-limit 20 --error-exitcode 0"
run() { # name tool command...
    name=$1; tool=$2; shift 2
    echo "== $name / $tool" | tee $out/r02_sanitizer_${name}_${tool}.log
    timeout 900 $cs --tool $tool "$@" 2>&1 | grep -v "performance database" | tail -25 >> $out/r02_sanitizer_${name}_${tool}.log
    tail -4 $out/r02_sanitizer_${name}_${tool}.log
}
for tool in memcheck racecheck synccheck; do
    run gn_split_pair_batch1 $tool python tools/sanitize_driver.py gn1
    run gn_split_pair_frame_call $tool --kernel-name kns=k_gn_persistent python tools/sanitize_driver.py frame1
    # eight sequences through ONE pair, one after the other (per-sequence hand-over blocks, restaged lists)
    run gn_split_pair_8seq $tool --kernel-name kns=k_gn_persistent python tools/sanitize_driver.py frame8
    run prepare_frame_border_lists_1seq $tool --kernel-name kns=k_prepare_frame python tools/sanitize_driver.py frame1
    run prepare_frame_border_lists_8seq $tool --kernel-name kns=k_prepare_frame python tools/sanitize_driver.py frame8
done
