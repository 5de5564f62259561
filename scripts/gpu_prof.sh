# ncu --set full capture of our kernels at config-3 shape.  Usage: bash scripts/gpu_prof.sh <tag> <kernel regex> [count]
mkdir -p gpurun_out
TAG=$1; PAT=$2; CNT=${3:-6}
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:$PAT" -s ${SKIP:-0} -c $CNT -f -o gpurun_out/prof_$TAG \
   python scripts/prof_kernels.py --iters 1 --what ${WHAT:-triplet,egt,ffn} > gpurun_out/prof_$TAG.log 2>&1
tail -3 gpurun_out/prof_$TAG.log
ls -la gpurun_out/prof_$TAG.ncu-rep
