#!/bin/bash
# One parameterised GPU-box job runner (see tools/gpu_runs/README.md).  Run from the repository root on the GPU box:
#   bash tools/gpu_job.sh <mode> [args...] [-- <mode> [args...]]...
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
PORT=29600

run_mode() {
  local mode="$1"; shift
  case "$mode" in
    tests)          # [pytest args]  (default: the whole -m gpu suite)
      local args=("$@"); [ ${#args[@]} -eq 0 ] && args=(tests)
      timeout "${KT_TEST_TIMEOUT:-1500}" python -m pytest "${args[@]}" -m gpu -x -q -p no:cacheprovider 2>&1 | tail -n "${KT_TAIL:-40}" | tee -a gpurun_out/tests.log
      ;;
    tests-all)      # like tests, but without -x (collect every failure)
      local args=("$@"); [ ${#args[@]} -eq 0 ] && args=(tests)
      timeout "${KT_TEST_TIMEOUT:-1500}" python -m pytest "${args[@]}" -m gpu -q -p no:cacheprovider 2>&1 | tail -n "${KT_TAIL:-60}" | tee -a gpurun_out/tests.log
      ;;
    bench)          # <tag> [bench.py args]
      local tag="$1"; shift
      timeout "${KT_BENCH_TIMEOUT:-900}" python bench.py "$@" > "gpurun_out/bench_${tag}.json" 2> "gpurun_out/bench_${tag}.err"
      echo "bench ${tag} rc=$?"; tail -c 3000 "gpurun_out/bench_${tag}.json"; tail -n 5 "gpurun_out/bench_${tag}.err"
      ;;
    multi)          # <N> <tag> [bench.py args]   (gpurun --gpus N)
      local n="$1" tag="$2"; shift 2
      PORT=$((PORT + 1))
      timeout "${KT_BENCH_TIMEOUT:-900}" $TR --nproc-per-node "$n" --master-port "$PORT" bench.py --gpus "$n" "$@" > "gpurun_out/bench_${tag}.json" 2> "gpurun_out/bench_${tag}.err"
      echo "multi ${tag} rc=$?"; tail -c 3000 "gpurun_out/bench_${tag}.json"; tail -n 8 "gpurun_out/bench_${tag}.err"
      ;;
    launches)       # <tag> [bench.py args]: ncu launch list of the bench command (per-launch times are cold and serialised)
      # env: KT_NCU_SKIP / KT_NCU_COUNT (window), KT_WORDS (plane words -> traffic per word), KT_COMMIT (recorded in the json)
      local tag="$1"; shift
      timeout 1800 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -s "${KT_NCU_SKIP:-0}" -c "${KT_NCU_COUNT:-600}" --csv --log-file "gpurun_out/launches_${tag}.csv" \
        python bench.py --no-e2e --no-cpu-baseline --no-check --no-secondary "$@" > "gpurun_out/launches_${tag}.log" 2>&1
      echo "launches ${tag} rc=$?"
      python tools/ncu_summary.py "gpurun_out/launches_${tag}.csv" --words "${KT_WORDS:-0}" --json "gpurun_out/traffic_${tag}.json" \
        --commit "${KT_COMMIT:-unknown}" --command "python bench.py $*" | tee "gpurun_out/launches_${tag}.md"
      ;;
    ncufull)        # <tag> <kernel regex> [bench.py args]: one full capture of a kernel
      local tag="$1" kern="$2"; shift 2
      timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:${kern}" -s "${KT_NCU_SKIP:-8}" -c "${KT_NCU_COUNT:-2}" \
        -o "gpurun_out/ncu_${tag}" -f python bench.py --no-e2e --no-cpu-baseline --no-check --no-secondary "$@" > "gpurun_out/ncu_${tag}.log" 2>&1
      echo "ncufull ${tag} rc=$?"; ls -la gpurun_out/ncu_${tag}.ncu-rep
      ;;
    ncumulti)       # <N> <tag> <kernel regex> [bench.py args]: full capture of one launch of a kernel on every rank of a torchrun job
      local n="$1" tag="$2" kern="$3"; shift 3
      PORT=$((PORT + 1))
      timeout 1500 ncu --target-processes all --set full --clock-control none --import-source on -k "regex:${kern}" -s "${KT_NCU_SKIP:-8}" -c 1 \
        -o "gpurun_out/ncu_${tag}_%p" -f $TR --nproc-per-node "$n" --master-port "$PORT" bench.py --gpus "$n" --no-e2e --no-cpu-baseline --no-check --no-secondary "$@" \
        > "gpurun_out/ncu_${tag}.log" 2>&1
      echo "ncumulti ${tag} rc=$?"; ls -la gpurun_out/ncu_${tag}_*.ncu-rep; tail -n 5 "gpurun_out/ncu_${tag}.log"
      ;;
    py)             # <script> [args]
      timeout "${KT_PY_TIMEOUT:-1200}" python "$@" 2>&1 | tail -n "${KT_TAIL:-60}"
      ;;
    smoke)
      timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -n 5
      ;;
    *) echo "unknown mode: $mode"; return 2;;
  esac
}

cur=()
for a in "$@"; do
  if [ "$a" == "--" ]; then
    [ ${#cur[@]} -gt 0 ] && run_mode "${cur[@]}"
    cur=()
  else
    cur+=("$a")
  fi
done
[ ${#cur[@]} -gt 0 ] && run_mode "${cur[@]}"
