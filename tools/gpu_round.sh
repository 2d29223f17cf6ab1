#!/usr/bin/env bash
# One gpurun call that produces everything a round needs from ONE B200 (saves the ~1 min of box time every call costs):
#   gpurun --timeout 1200 -- 'bash tools/gpu_round.sh r02'
# Outputs land in gpurun_out/<tag>_*; summarise the .ncu-rep files here afterwards with tools/ncu_summary.py.
set -u
tag="${1:-rXX}"
out=gpurun_out
mkdir -p "$out"
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 | tee "$out/${tag}_pytest_gpu.txt"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1 | tee "$out/${tag}_smoke.txt"
python tools/tune_assign.py > "$out/${tag}_tune_assign.txt" 2>&1
python bench.py > "$out/${tag}_bench.json" 2> "$out/${tag}_bench.err"
python bench.py --impl reference > "$out/${tag}_bench_reference.json" 2>/dev/null
RIO_BENCH_NO_SMI=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 \
    --csv --log-file "$out/${tag}_launches.csv" python bench.py --steps 5 --warmup 3 > /dev/null 2>&1
RIO_BENCH_NO_SMI=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_assign_hrw_v2 --launch-skip 6 -c 1 -f \
    -o "$out/${tag}_ncu_assign" python bench.py --steps 3 --warmup 3 --no-extra > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_affinity_umma --launch-skip 1 -c 1 -f \
    -o "$out/${tag}_ncu_umma" python tools/profile_umma.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_affinity_resolve --launch-skip 1 -c 1 -f \
    -o "$out/${tag}_ncu_resolve" python tools/profile_umma.py > /dev/null 2>&1
timeout 200 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > "$out/${tag}_sanitizer_memcheck.log" 2>&1
tail -n 2 "$out/${tag}_sanitizer_memcheck.log"
python - <<PY
import json
d = json.load(open("$out/${tag}_bench.json"))
print("value %.4g  ms/step %.4f  kernel %.4f  alu frac %.3f  e2e %.4g  clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["alu_roofline"]["frac"], d["e2e"]["value"], d["clocks"]))
PY
ls -la "$out" | grep "${tag}_" | awk '{print $5, $9}'
