#!/bin/bash
# quick single-GPU sweep of GEMM scheduling switches for the bf16x3 step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
run() { tag=$1; shift; env "$@" timeout 100 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $O/s19_$tag.json 2> $O/s19_$tag.err; python - <<PY
import json
for l in open("$O/s19_$tag.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$tag", round(d["ms_per_step"], 3), round(d["value"]), round(d["roofline"]["gemm_ms_per_step"], 3))
PY
}
run base X=1
run pairk1024 GRAPPA_B200_GEMM_PAIR_MINK_X3=1024
run max100 GRAPPA_B200_GEMM_MAX_SMS=100
run max132 GRAPPA_B200_GEMM_MAX_SMS=132
run grp80 GRAPPA_B200_GEMM_GROUP_SMS=80
run grp116 GRAPPA_B200_GEMM_GROUP_SMS=116
run bn128 GRAPPA_B200_GEMM_BN=128
