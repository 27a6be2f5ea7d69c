#!/bin/bash
# compute-sanitizer over the small parity cases (golden vectors of the reference, every boundary of the clustered
# stream, long-read records, long reads through every kernel family, the one-pass report).  Run on the GPU box:
#     bash tools/profile/sanitize.sh            -> gpurun_out/r2_sanitizer_{memcheck,racecheck,synccheck}.log
# The logs are copied to profiles/ by hand once they are clean.
set -u
out=${1:-gpurun_out}
mkdir -p "$out"
small="tests/test_gpu_golden.py::test_gpu_matches_reference_golden tests/test_gpu_parity.py::test_ell_stream_boundaries \
tests/test_gpu_parity.py::test_ell_long_read_records tests/test_gpu_parity.py::test_long_reads_every_path \
tests/test_gpu_parity.py::test_one_pass_report_equals_seven_reassign_calls tests/test_gpu_parity.py::test_edge_shapes \
tests/test_gpu_parity.py::test_identical_loci_keep_exact_ties_under_renumbering"
race="tests/test_gpu_golden.py::test_gpu_matches_reference_golden tests/test_gpu_parity.py::test_ell_stream_boundaries \
tests/test_gpu_parity.py::test_ell_long_read_records tests/test_gpu_parity.py::test_edge_shapes"
for tool in memcheck racecheck synccheck; do
    sel=$small
    [ "$tool" = racecheck ] && sel=$race
    log="$out/r2_sanitizer_$tool.log"
    echo "== compute-sanitizer --tool $tool  ($(date -u +%FT%TZ))" > "$log"
    echo "== python -m pytest -x -q -p no:cacheprovider $sel" >> "$log"
    t0=$(date +%s)
    timeout ${SANITIZE_TIMEOUT:-420} compute-sanitizer --tool $tool --error-exitcode 86 --print-limit 60 \
        python -m pytest -x -q -p no:cacheprovider $sel >> "$log" 2>&1
    echo "== exit code $? after $(( $(date +%s) - t0 )) s" >> "$log"
    tail -4 "$log"
done
