mkdir -p gpurun_out; : > gpurun_out/sweep.log
run() { echo "== $*" >> gpurun_out/sweep.log; env "$@" python tools/counters.py ${SCENE:-million} 0 0 ${SPP:-16} 2>&1 | tail -${TAILN:-2} >> gpurun_out/sweep.log; }
run KFRT_LIB_DIR=kuafu_b200/lib
for v in $VARIANTS; do TAILN=5 run KFRT_LIB_DIR=kuafu_b200/lib_$v; done
for e in $ENVS; do run KFRT_LIB_DIR=kuafu_b200/lib $e; done
for e in $ENVS2; do run KFRT_LIB_DIR=kuafu_b200/lib_$V2 $e; done
run KFRT_LIB_DIR=kuafu_b200/lib
cat gpurun_out/sweep.log
