#!/bin/bash
# compute-sanitizer runs kept as evidence under profiles/r02_sanitizer_*.log (racecheck / synccheck / memcheck on
# the GPU box).  usage: tools/sanitize.sh <outdir>
out=${1:-gpurun_out}
cs="compute-sanitizer --print-limit 20 --error-exitcode 0"
run() { # name tool command...
    name=$1; tool=$2; shift 2
    echo "== $name / $tool" | tee $out/r02_sanitizer_${name}_${tool}.log
    timeout 900 $cs --tool $tool "$@" 2>&1 | grep -v "performance database" | tail -25 >> $out/r02_sanitizer_${name}_${tool}.log
    tail -4 $out/r02_sanitizer_${name}_${tool}.log
}
for tool in memcheck racecheck synccheck; do
    run gn_persistent_batch1 $tool python tools/sanitize_driver.py gn1
    run gn_persistent_batch3 $tool python tools/sanitize_driver.py gn3
    run batched_engine_8seq $tool python tools/sanitize_driver.py batch8
    run predict_splat_resolve $tool python -m pytest tests/test_predict.py -q -m gpu -x -k "combined_predict_is_bit_exact"
    run fern_search $tool python -m pytest tests/test_ferns.py -q -m gpu -x -k "encode_search_and_database"
    # the one-launch frame preparation (shared-memory tiles, cp.async staging), alone: the tracker behind it is covered above
    run prepare_frame_1seq $tool --kernel-name kns=k_prepare_frame python tools/sanitize_driver.py frame1
    run prepare_frame_8seq $tool --kernel-name kns=k_prepare_frame python tools/sanitize_driver.py frame8
done
