set -x
tools/_bin/probe_cub_sort 16777216 557056 > gpurun_out/r2v_cub_sort.jsonl
tools/_bin/probe_cub_sort 4980736 155648 >> gpurun_out/r2v_cub_sort.jsonl
tools/_bin/probe_cub_sort 67108864 1966080 >> gpurun_out/r2v_cub_sort.jsonl
cat gpurun_out/r2v_cub_sort.jsonl
