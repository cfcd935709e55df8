python - <<'PY'
import ctypes
cuda = ctypes.CDLL("libcudart.so")
v = ctypes.c_int()
for name, attr in (("MaxPersistingL2CacheSize", 108), ("L2CacheSize", 38), ("MaxAccessPolicyWindowSize", 109)):
    cuda.cudaDeviceGetAttribute(ctypes.byref(v), attr, 0)
    print(name, v.value)
PY
VQWN_TC_FLAGS=16 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:wavenet_tc_cluster -c 2 python tools/tc_time.py 64 64 tc 2>&1 | grep -v "^==PROF==" | tail -25
