set -x
cd $GRAFT_REPO_ROOT
JMPC_LIB=$PWD/build/variants/lib_exp.so python tests/tools/lat_threshold.py 2>&1 | grep "warps per SM" | tee gpurun_out/r2_lat_threshold.log
