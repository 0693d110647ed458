#!/bin/bash
# what the box has: GPUs, host cores / NUMA, PCIe, NVENC / NVDEC libraries and engine counters
OUT=gpurun_out/${1:-r2}_box.txt
{
nvidia-smi -L
nproc; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA node\(s\)"
nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current,clocks.max.sm,memory.total --format=csv
echo "--- encoder / decoder libraries"
ls /usr/lib/x86_64-linux-gnu 2>/dev/null | grep -E "nvidia-encode|nvcuvid|nvjpeg" 
ls /usr/local/cuda/lib64 2>/dev/null | grep -E "nvjpeg" | head -3
echo "--- nvidia-smi -q encoder/decoder sections"
nvidia-smi -q -i 0 | grep -iE -A4 "encoder|decoder|jpeg|ofa" | head -60
echo "--- ottgpu_nvenc_probe"
python - <<'PY'
import ctypes as C, os, subprocess
subprocess.check_call(["make","-s","-C","ott-packager_b200/host"])
C.CDLL(os.path.abspath("ott-packager_b200/libottgpu.so"), mode=C.RTLD_GLOBAL)
C.CDLL(os.path.abspath("oracle/_ref/libfillet_rt.so"), mode=C.RTLD_GLOBAL)
H=C.CDLL(os.path.abspath("ott-packager_b200/libottgpu_host.so"))
b=C.create_string_buffer(512); H.ottgpu_nvenc_probe.argtypes=[C.c_char_p,C.c_size_t]
print(H.ottgpu_nvenc_probe(b,512), b.value.decode())
PY
} > $OUT 2>&1
cat $OUT
