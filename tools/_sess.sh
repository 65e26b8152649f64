set -u
mkdir -p gpurun_out
MESHODE_B200_LIB=build/variants/libmeshode_cluster_notmem.so timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_mini.py cluster > gpurun_out/sanitize_racecheck_cluster_notmem.log 2>&1; echo "exit $?"; grep -v "Host Frame" gpurun_out/sanitize_racecheck_cluster_notmem.log | tail -8
MESHODE_B200_LIB=build/variants/libmeshode_cluster_notmem.so timeout 300 compute-sanitizer --tool synccheck python tools/sanitize_mini.py cluster > gpurun_out/sanitize_synccheck_cluster_notmem.log 2>&1; echo "exit $?"; grep -v "Host Frame" gpurun_out/sanitize_synccheck_cluster_notmem.log | tail -4
MESHODE_B200_LIB=build/variants/libmeshode_cluster_notmem.so timeout 300 python -m pytest tests/test_gpu_deform.py -q -k "cluster or partial" 2>&1 | tail -3
