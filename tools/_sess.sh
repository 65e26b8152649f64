set -u
mkdir -p gpurun_out
build/f32x2_probe > gpurun_out/f32x2_probe.log 2>&1; cat gpurun_out/f32x2_probe.log
