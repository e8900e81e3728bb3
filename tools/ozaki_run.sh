#!/bin/bash
# Build and run the INT8-tcgen05 Ozaki prototype on the GPU box; results -> gpurun_out/ozaki_<tag>.json
tag=${1:-r02}
mkdir -p gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o /tmp/ozaki_proto tools/ozaki_proto.cu -lcublas || exit 1
python tools/ozaki_gen.py /tmp/factor.bin 2048 2048 2048 > gpurun_out/ozaki_gen_$tag.log 2>&1
OZAKI_RATE=1 timeout 300 /tmp/ozaki_proto 2048 2048 /tmp/factor.bin > gpurun_out/ozaki_2048_$tag.json 2> gpurun_out/ozaki_2048_$tag.err; echo "2048 rc=$?"
cat gpurun_out/ozaki_2048_$tag.json | head -40
if [ "$2" != "small" ]; then
timeout 600 /tmp/ozaki_proto 8192 8192 > gpurun_out/ozaki_8192_$tag.json 2> gpurun_out/ozaki_8192_$tag.err; echo "8192 rc=$?"
cat gpurun_out/ozaki_8192_$tag.json | head -40
fi
