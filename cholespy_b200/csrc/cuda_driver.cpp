// CUDA driver shim: see cuda_driver.h.  New code; replaces reference src/cuda_driver.cpp.
#include "cuda_driver.h"

#include <dlfcn.h>

#include <cstring>
#include <map>
#include <mutex>
#include <string>

// The sm_100a cubin produced from kernels.cu, embedded by the build (kernels_image.cpp).
extern "C" const unsigned char chb_kernels_image[];
extern "C" const unsigned long long chb_kernels_image_size;

CUresult (*cuInit)(unsigned int) = nullptr;
CUresult (*cuDeviceGet)(CUdevice *, int) = nullptr;
CUresult (*cuDeviceGetCount)(int *) = nullptr;
CUresult (*cuDeviceGetAttribute)(int *, int, CUdevice) = nullptr;
CUresult (*cuDevicePrimaryCtxRetain)(CUcontext *, CUdevice) = nullptr;
CUresult (*cuDevicePrimaryCtxRelease)(CUdevice) = nullptr;
CUresult (*cuCtxPushCurrent)(CUcontext) = nullptr;
CUresult (*cuCtxPopCurrent)(CUcontext *) = nullptr;
CUresult (*cuCtxSynchronize)(void) = nullptr;
CUresult (*cuGetErrorName)(CUresult, const char **) = nullptr;
CUresult (*cuGetErrorString)(CUresult, const char **) = nullptr;
CUresult (*cuMemAlloc)(void **, size_t) = nullptr;
CUresult (*cuMemFree)(void *) = nullptr;
CUresult (*cuMemHostAlloc)(void **, size_t, unsigned int) = nullptr;
CUresult (*cuMemFreeHost)(void *) = nullptr;
CUresult (*cuMemcpyAsync)(void *, const void *, size_t, CUstream) = nullptr;
CUresult (*cuMemsetD8Async)(void *, unsigned char, size_t, CUstream) = nullptr;
CUresult (*cuMemsetD32Async)(void *, unsigned int, size_t, CUstream) = nullptr;
CUresult (*cuModuleLoadData)(CUmodule *, const void *) = nullptr;
CUresult (*cuModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
CUresult (*cuModuleUnload)(CUmodule) = nullptr;
CUresult (*cuFuncSetAttribute)(CUfunction, int, int) = nullptr;
CUresult (*cuFuncGetAttribute)(int *, int, CUfunction) = nullptr;
CUresult (*cuOccupancyMaxActiveBlocksPerMultiprocessor)(int *, CUfunction, int, size_t) = nullptr;
CUresult (*cuLaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                           void **, void **) = nullptr;
CUresult (*cuLaunchCooperativeKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                                      CUstream, void **) = nullptr;
CUresult (*cuStreamSynchronize)(CUstream) = nullptr;
CUresult (*cuStreamCreate)(CUstream *, unsigned int) = nullptr;
CUresult (*cuStreamDestroy)(CUstream) = nullptr;
CUresult (*cuStreamWaitEvent)(CUstream, CUevent, unsigned int) = nullptr;
CUresult (*cuEventCreate)(CUevent *, unsigned int) = nullptr;
CUresult (*cuEventDestroy)(CUevent) = nullptr;
CUresult (*cuEventRecord)(CUevent, CUstream) = nullptr;
CUresult (*cuMemcpy2DAsync)(const CUDA_MEMCPY2D *, CUstream) = nullptr;
CUresult (*cuPointerGetAttribute)(void *, int, CUdeviceptr) = nullptr;
CUresult (*cuDeviceCanAccessPeer)(int *, CUdevice, CUdevice) = nullptr;
CUresult (*cuCtxEnablePeerAccess)(CUcontext, unsigned int) = nullptr;
CUresult (*cuMemcpyPeerAsync)(CUdeviceptr, CUcontext, CUdeviceptr, CUcontext, size_t, CUstream) = nullptr;

CUdevice cu_device = 0;
CUcontext cu_context = nullptr;

namespace {
void *g_libcuda = nullptr;
bool g_symbols_ok = false;
std::mutex g_mutex;
std::string g_error;                                  // initialisation failures (set under g_mutex)
thread_local CUresult g_first_error = CUDA_SUCCESS;   // per calling thread: concurrent C-ABI calls do not lose or
thread_local std::string g_cuda_error;                // misattribute each other's CUDA errors
std::map<int, chb::DeviceState *> g_devices;
std::map<std::pair<int, std::string>, CUfunction> g_functions;

constexpr int kAttrMultiprocessorCount = 16;       // CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT
constexpr int kAttrComputeCapabilityMajor = 75;    // CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR
constexpr int kAttrComputeCapabilityMinor = 76;    // CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR
constexpr int kPointerAttrDeviceOrdinal = 9;       // CU_POINTER_ATTRIBUTE_DEVICE_ORDINAL

template <typename Fn>
bool resolve(Fn &slot, const char *name) {
    slot = reinterpret_cast<Fn>(dlsym(g_libcuda, name));
    if (!slot) g_error = std::string("libcuda: missing symbol ") + name;
    return slot != nullptr;
}

bool load_symbols() {
    if (g_symbols_ok) return true;
    if (!g_libcuda) g_libcuda = dlopen("libcuda.so.1", RTLD_LAZY);
    if (!g_libcuda) g_libcuda = dlopen("libcuda.so", RTLD_LAZY);
    if (!g_libcuda) {
        g_error = "CUDA driver (libcuda.so) not found";
        return false;
    }
    // Async entry points are the legacy-default-stream variants, like the reference
    // (src/cuda_driver.cpp:96-109), so frameworks' blocking streams order against them.
    bool ok = resolve(cuInit, "cuInit") && resolve(cuDeviceGet, "cuDeviceGet") &&
              resolve(cuDeviceGetCount, "cuDeviceGetCount") && resolve(cuDeviceGetAttribute, "cuDeviceGetAttribute") &&
              resolve(cuDevicePrimaryCtxRetain, "cuDevicePrimaryCtxRetain") &&
              resolve(cuDevicePrimaryCtxRelease, "cuDevicePrimaryCtxRelease_v2") &&
              resolve(cuCtxPushCurrent, "cuCtxPushCurrent_v2") && resolve(cuCtxPopCurrent, "cuCtxPopCurrent_v2") &&
              resolve(cuCtxSynchronize, "cuCtxSynchronize") && resolve(cuGetErrorName, "cuGetErrorName") &&
              resolve(cuGetErrorString, "cuGetErrorString") && resolve(cuMemAlloc, "cuMemAlloc_v2") &&
              resolve(cuMemFree, "cuMemFree_v2") && resolve(cuMemHostAlloc, "cuMemHostAlloc") &&
              resolve(cuMemFreeHost, "cuMemFreeHost") && resolve(cuMemcpyAsync, "cuMemcpyAsync") &&
              resolve(cuMemsetD8Async, "cuMemsetD8Async") && resolve(cuMemsetD32Async, "cuMemsetD32Async") &&
              resolve(cuModuleLoadData, "cuModuleLoadData") && resolve(cuModuleGetFunction, "cuModuleGetFunction") &&
              resolve(cuModuleUnload, "cuModuleUnload") && resolve(cuFuncSetAttribute, "cuFuncSetAttribute") &&
              resolve(cuFuncGetAttribute, "cuFuncGetAttribute") &&
              resolve(cuOccupancyMaxActiveBlocksPerMultiprocessor, "cuOccupancyMaxActiveBlocksPerMultiprocessor") &&
              resolve(cuLaunchKernel, "cuLaunchKernel") &&
              resolve(cuLaunchCooperativeKernel, "cuLaunchCooperativeKernel") &&
              resolve(cuStreamSynchronize, "cuStreamSynchronize") && resolve(cuStreamCreate, "cuStreamCreate") &&
              resolve(cuStreamDestroy, "cuStreamDestroy_v2") && resolve(cuStreamWaitEvent, "cuStreamWaitEvent") &&
              resolve(cuEventCreate, "cuEventCreate") && resolve(cuEventDestroy, "cuEventDestroy_v2") &&
              resolve(cuEventRecord, "cuEventRecord") && resolve(cuMemcpy2DAsync, "cuMemcpy2DAsync_v2") &&
              resolve(cuPointerGetAttribute, "cuPointerGetAttribute") &&
              resolve(cuDeviceCanAccessPeer, "cuDeviceCanAccessPeer") && resolve(cuCtxEnablePeerAccess, "cuCtxEnablePeerAccess") &&
              resolve(cuMemcpyPeerAsync, "cuMemcpyPeerAsync");
    if (!ok) return false;
    CUresult rc = cuInit(0);
    if (rc != CUDA_SUCCESS) {
        g_error = "cuInit failed with code " + std::to_string(rc);
        return false;
    }
    g_symbols_ok = true;
    return true;
}
}  // namespace

void cuda_check_impl(CUresult errval, const char *file, const int line) {
    if (errval == CUDA_SUCCESS || errval == CUDA_ERROR_DEINITIALIZED) return;
    const char *name = nullptr, *msg = nullptr;
    if (cuGetErrorName) cuGetErrorName(errval, &name);
    if (cuGetErrorString) cuGetErrorString(errval, &msg);
    fprintf(stderr, "cuda_check(): API error = %04d (%s): \"%s\" in %s:%i.\n", (int) errval, name ? name : "?",
            msg ? msg : "?", file, line);
    if (g_first_error == CUDA_SUCCESS) {
        g_first_error = errval;
        g_cuda_error = std::string("CUDA error ") + std::to_string(errval) + " (" + (name ? name : "?") + ") at " + file + ":" +
                  std::to_string(line);
    }
}
CUresult cuda_first_error() { return g_first_error; }
void cuda_clear_error() {
    g_first_error = CUDA_SUCCESS;
    g_cuda_error.clear();
}

namespace chb {

const char *driver_error_text() { return g_cuda_error.empty() ? g_error.c_str() : g_cuda_error.c_str(); }

DeviceState *device_state(int ordinal) {
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_devices.find(ordinal);
    if (it != g_devices.end()) return it->second;
    if (!load_symbols()) return nullptr;
    int count = 0;
    if (cuDeviceGetCount(&count) != CUDA_SUCCESS || ordinal < 0 || ordinal >= count) {
        g_error = "CUDA device " + std::to_string(ordinal) + " not available (" + std::to_string(count) + " visible)";
        return nullptr;
    }
    DeviceState *ds = new DeviceState();
    ds->ordinal = ordinal;
    CUresult rc = cuDeviceGet(&ds->dev, ordinal);
    if (rc == CUDA_SUCCESS) rc = cuDeviceGetAttribute(&ds->sm_count, kAttrMultiprocessorCount, ds->dev);
    if (rc == CUDA_SUCCESS) rc = cuDeviceGetAttribute(&ds->cc_major, kAttrComputeCapabilityMajor, ds->dev);
    if (rc == CUDA_SUCCESS) rc = cuDeviceGetAttribute(&ds->cc_minor, kAttrComputeCapabilityMinor, ds->dev);
    if (rc == CUDA_SUCCESS && ds->cc_major != 10) {
        g_error = "device " + std::to_string(ordinal) + " is sm_" + std::to_string(ds->cc_major * 10 + ds->cc_minor) +
                  "; this build carries sm_100a code only (no fallback path)";
        delete ds;
        return nullptr;
    }
    if (rc == CUDA_SUCCESS) rc = cuDevicePrimaryCtxRetain(&ds->ctx, ds->dev);
    if (rc == CUDA_SUCCESS) {
        rc = cuCtxPushCurrent(ds->ctx);
        if (rc == CUDA_SUCCESS) {
            rc = cuModuleLoadData(&ds->module, chb_kernels_image);
            cuCtxPopCurrent(nullptr);
        }
    }
    if (rc != CUDA_SUCCESS) {
        const char *name = nullptr;
        cuGetErrorName(rc, &name);
        g_error = "CUDA device " + std::to_string(ordinal) + " initialisation failed: " + (name ? name : "?");
        delete ds;
        return nullptr;
    }
    g_devices[ordinal] = ds;
    return ds;
}

CUfunction kernel(DeviceState *ds, const char *name) {
    std::lock_guard<std::mutex> lock(g_mutex);
    auto key = std::make_pair(ds->ordinal, std::string(name));
    auto it = g_functions.find(key);
    if (it != g_functions.end()) return it->second;
    CUfunction fn = nullptr;
    CUresult rc = cuModuleGetFunction(&fn, ds->module, name);
    if (rc != CUDA_SUCCESS) {
        g_error = std::string("kernel not found in embedded cubin: ") + name;
        return nullptr;
    }
    g_functions[key] = fn;
    return fn;
}

int max_resident_blocks(DeviceState *ds, CUfunction fn, int block, size_t smem) {
    int per_sm = 0;
    cuda_check(cuOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, block, smem));
    return per_sm * ds->sm_count;
}

int device_of_pointer(const void *dptr) {
    if (!load_symbols()) return -1;
    int ordinal = -1;
    if (cuPointerGetAttribute(&ordinal, kPointerAttrDeviceOrdinal, (CUdeviceptr) dptr) != CUDA_SUCCESS) return -1;
    return ordinal;
}

}  // namespace chb

bool init_cuda() {
    chb::DeviceState *ds = chb::device_state(0);
    if (!ds) {
        fprintf(stderr, "init_cuda(): %s -- disabling CUDA backend!\n", chb::driver_error_text());
        return false;
    }
    cu_device = ds->dev;
    cu_context = ds->ctx;
    return true;
}

void shutdown_cuda() {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (!g_libcuda) return;
    for (auto &kv : g_devices) {
        cuda_check(cuDevicePrimaryCtxRelease(kv.second->dev));
        delete kv.second;
    }
    g_devices.clear();
    g_functions.clear();
    cu_context = nullptr;
}
