"""Which NVLink byte counters does NVML expose on this box?  (bench.nvlink_counters picks the first that works.)"""
import pynvml as N

N.nvmlInit()
h = N.nvmlDeviceGetHandleByIndex(0)
print("driver", N.nvmlSystemGetDriverVersion(), "name", N.nvmlDeviceGetName(h))
for name in ["NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX", "NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX",
             "NVML_FI_DEV_NVLINK_THROUGHPUT_RAW_RX", "NVML_FI_DEV_NVLINK_THROUGHPUT_RAW_TX"]:
    fid = getattr(N, name)
    for scope in (0xFFFFFFFF, 0, 1, 17):
        try:
            v = N.nvmlDeviceGetFieldValues(h, [(fid, scope)])[0]
            print(name, "scope", hex(scope), "ret", v.nvmlReturn, "type", v.valueType, "ull", v.value.ullVal, "ul", v.value.ulVal)
        except Exception as e:
            print(name, "scope", hex(scope), "EXC", repr(e))
for link in range(0, 18):
    try:
        st = N.nvmlDeviceGetNvLinkState(h, link)
    except Exception as e:
        print("link", link, "state EXC", repr(e))
        break
    print("link", link, "state", st)
