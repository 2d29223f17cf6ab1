#!/usr/bin/env bash
# One gpurun call that produces everything a round needs from ONE B200 (saves the ~1 min of box time every call costs):
#   gpurun --timeout 1800 -- 'bash tools/gpu_round.sh r02'
# Outputs land in gpurun_out/<tag>_*; summarise here afterwards: tools/ncu_summary.py (one kernel -> JSON), tools/ncu_durations.py
# (launch list -> one row per kernel), tools/sass_grep.py (no GPU needed).
set -u
tag="${1:-rXX}"
out=gpurun_out
mkdir -p "$out"
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 | tee "$out/${tag}_pytest_gpu.txt"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1 | tee "$out/${tag}_smoke.txt"
python bench.py > "$out/${tag}_bench.json" 2> "$out/${tag}_bench.err"
python bench.py --impl reference > "$out/${tag}_bench_reference.json" 2>/dev/null
python tools/bench_kernels.py > "$out/${tag}_bench_kernels.jsonl" 2>&1
RIO_BENCH_NO_SMI=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file "$out/${tag}_kernel_launches.csv" python tools/bench_kernels.py > /dev/null 2>&1
RIO_BENCH_NO_SMI=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 \
    --csv --log-file "$out/${tag}_bench_launches.csv" python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_assign_trie --launch-skip 3 -c 1 -f \
    -o "$out/${tag}_ncu_trie" python tools/profile_trie.py > /dev/null 2>&1
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > "$out/${tag}_sanitizer_memcheck.log" 2>&1
tail -n 2 "$out/${tag}_sanitizer_memcheck.log"
timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py --no-umma > "$out/${tag}_sanitizer_racecheck.log" 2>&1
tail -n 2 "$out/${tag}_sanitizer_racecheck.log"
python - <<PY
import json
d = json.load(open("$out/${tag}_bench.json"))
print("value %.4g  ms/step %.4f  kernel %.4f  hbm frac %.3f  e2e %.4g  clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["clocks"]))
PY
ls -la "$out" | grep "${tag}_" | awk '{print $5, $9}'
