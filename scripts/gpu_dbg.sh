mkdir -p gpurun_out
export HSTU_DEBUG_SPIN=1
python -m generative_recommenders_b200.build --force > gpurun_out/dbg_build.log 2>&1
for cfg in "1 4096 2" "1 8192 1" "2 8192 8"; do echo "== $cfg"; timeout 200 python scripts/dbg_bwd.py $cfg 2>&1 | sort | uniq -c | sort -rn | head -14; done > gpurun_out/dbg_bwd.log 2>&1
cat gpurun_out/dbg_bwd.log | head -80
